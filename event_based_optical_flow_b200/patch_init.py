"""The per-patch initialiser of the pyramid: batched 2-dof candidate costs.

The reference, at every pyramid level above the coarsest, runs one Optuna TPE study PER PATCH and evaluates every trial with
`calculate_cost_for_small_patch` -- crop, origin shift, numpy warp, numpy vote, scipy Gaussian, cv2 Sobel -- on the CPU
(src/solver/patch_contrast_pyramid.py:320-415).  `PatchCandidateEvaluator` prepares the patches' events once per frame and
level (one grouped copy on the GPU, 12 bytes per event, and the un-warped images' energies) and then answers "K candidates for
each of P patches" with one C-ABI call (cmax_patch_candidates: ONE kernel for the whole batch when the patch image fits in
shared memory, which it does at every level of a 260x346 sensor).  `run_patch_studies` is the reference's loop with the
patch and trial loops swapped: trial t of ALL patches is asked, evaluated together, and told -- every study still sees exactly
its own history, so each patch's TPE sampling is the sequence the reference would have run.

No CPU fallback: the evaluator needs the CUDA library and a CUDA device.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _rect(p) -> Tuple[int, int, int, int]:
    if hasattr(p, "x_min"):
        return int(p.x_min), int(p.x_max), int(p.y_min), int(p.y_max)
    x0, x1, y0, y1 = p
    return int(x0), int(x1), int(y0), int(y1)


def group_patch_events(ev: torch.Tensor, rects: torch.Tensor, normalize_t: bool = True):
    """-> (patch_events [m,3] f32 = (x - x_min, y - y_min, dt), offsets [P+1] i64, counts [P], theta_scale [P] f64) on ev's
    device, in ev's dtype until the final cast: every patch's events in event order (crop_event + set_event_origin_to_zero,
    src/utils/event_utils.py:50-88), dt = (t - t_middle) / (max dt - min dt) of the PATCH's events (src/warp.py:217-225, :250-258),
    theta_scale = the patch's time span the reference multiplies a candidate by (pyramid.py:366-371; 1 without normalisation)."""
    P = int(rects.shape[0])
    x, y, t = ev[:, 0], ev[:, 1], ev[:, 2]
    # the crop masks of all patches at once, a bounded number of patches per pass; nonzero() lists them patch by patch
    per_pass = max(1, (1 << 26) // max(1, len(ev)))
    pid, eid = [], []
    for lo in range(0, P, per_pass):
        r = rects[lo:lo + per_pass]
        inside = (r[:, 0:1] <= x) & (x < r[:, 1:2]) & (r[:, 2:3] <= y) & (y < r[:, 3:4])
        pp, ee = torch.nonzero(inside, as_tuple=True)
        pid.append(pp + lo)
        eid.append(ee)
    pid, eid = torch.cat(pid), torch.cat(eid)
    counts = torch.bincount(pid, minlength=P)
    offsets = torch.zeros(P + 1, dtype=torch.int64, device=ev.device)
    offsets[1:] = torch.cumsum(counts, 0)
    tp = t[eid]

    def seg(values, how):
        init = torch.full((P,), float("inf") if how == "amin" else float("-inf"), dtype=ev.dtype, device=ev.device)
        return init.scatter_reduce(0, pid, values, how)

    t_min, t_max = seg(tp, "amin"), seg(tp, "amax")
    span = t_max - t_min
    dt = tp - (t_min + span * 0.5)[pid]
    if normalize_t:
        dt = dt / (seg(dt, "amax") - seg(dt, "amin"))[pid]
    patch_events = torch.stack([x[eid] - rects[pid, 0], y[eid] - rects[pid, 2], dt], 1).to(torch.float32).contiguous()
    scale = (span if normalize_t else torch.ones_like(span)).double()
    scale = torch.where(counts > 0, scale, torch.zeros_like(scale))
    return patch_events, offsets, counts, scale


class PatchCandidateEvaluator:
    """loss[p, k] of candidate (trans_x, trans_y)[p, k] on patch p == `objective_initial(trial, cropped events of p, .)` of the
    reference (src/solver/patch_contrast_pyramid.py:364-415).

    events: [n, 4] (x, y, t, p) tensor or array, any float dtype (the time normalisation is done in THAT dtype, the kernels take
    the fp32 result); patches: rectangles (x_min, x_max, y_min, y_max) or the reference's Patch objects (x = row axis);
    patch_image_size: the level's `scaled_patch_size` (the image every patch is voted into); outer_padding, sigma,
    normalize_t: `solver.padding`, `iwe.blur_sigma`, `normalize_t_in_batch`.
    `valid[p]`: the patch has more than `min_events` events (the reference leaves the others at their initial guess, :338).
    `global_images=True` forces the multi-kernel path with images in global memory (the default picks the one-launch
    shared-memory kernel whenever two patch images fit, cmax_patch_candidates in include/cmax_b200.h)."""

    def __init__(self, events, patches: Sequence, patch_image_size: Tuple[int, int], *, outer_padding=0, sigma: float = 1.0,
                 normalize_t: bool = True, min_events: int = 10, device: Optional[torch.device] = None,
                 global_images: Optional[bool] = None):
        ev = torch.as_tensor(events)
        if device is None:
            device = ev.device if ev.is_cuda else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
        if device is None:
            raise RuntimeError("PatchCandidateEvaluator needs a CUDA device (no CPU fallback)")
        _lib.load()
        self.device = torch.device(device)
        ev = ev.detach().to(self.device)
        if not ev.dtype.is_floating_point:
            ev = ev.double()
        if ev.ndim != 2 or ev.shape[1] < 3:
            raise ValueError(f"events must be [n, >=3], got {tuple(ev.shape)}")
        self.image_size = (int(patch_image_size[0]), int(patch_image_size[1]))
        pad = (int(outer_padding), int(outer_padding)) if isinstance(outer_padding, (int, float)) else tuple(int(p) for p in outer_padding)
        self.pad = pad
        self.padded_size = (self.image_size[0] + 2 * pad[0], self.image_size[1] + 2 * pad[1])
        self.sigma = float(sigma)
        self.normalize_t = bool(normalize_t)
        rects = torch.tensor([_rect(p) for p in patches], dtype=ev.dtype, device=self.device)  # [P, 4]
        self.n_patches = P = int(rects.shape[0])
        if P == 0:
            raise ValueError("no patches")
        with torch.cuda.device(self.device):
            self.patch_events, self.offsets, counts, self.theta_scale = group_patch_events(ev, rects, self.normalize_t)
            counts_host = counts.cpu()
        self.counts = counts_host.numpy()
        self.valid = self.counts > int(min_events)
        self.max_patch_events = int(counts_host.max()) if P else 0
        self.flags = 0 if global_images is None else (_lib.PATCH_GLOBAL_IMAGES if global_images else 0)
        self._workspace = None
        self._workspace_k = 0
        # the un-warped image's energy (orig_iwe, pyramid.py:391-396) does not depend on the candidate: once per frame and level
        self.orig_energy = self._run(torch.zeros(P, 1, 2, dtype=torch.float64, device=self.device), None, None, keep_images=False)[:, 0].contiguous()

    def _run(self, cand: torch.Tensor, scale, orig, keep_images: bool) -> torch.Tensor:
        if cand.ndim != 3 or cand.shape[0] != self.n_patches or cand.shape[1] < 1 or cand.shape[2] != 2:
            raise ValueError(f"candidates must be [{self.n_patches}, K >= 1, 2], got {tuple(cand.shape)}")
        P, K = int(cand.shape[0]), int(cand.shape[1])
        cand = cand.detach().to(device=self.device, dtype=torch.float64).contiguous()
        flags = self.flags | (_lib.PATCH_KEEP_IMAGES if keep_images else 0)
        with torch.cuda.device(self.device):
            if self._workspace is None or self._workspace_k < K:
                nbytes = _lib.load().cmax_patch_candidates_workspace_bytes(P, K, *self.image_size, *self.pad)
                self._workspace = torch.empty(max(int(nbytes), 4), dtype=torch.uint8, device=self.device)
                self._workspace_k = K
            out = torch.empty(P, K, dtype=torch.float64, device=self.device)
            _lib.call("cmax_patch_candidates", self.patch_events.data_ptr(), self.offsets.data_ptr(), self.max_patch_events, P,
                      cand.data_ptr(), scale.data_ptr() if scale is not None else None, K, self.image_size[0], self.image_size[1],
                      self.pad[0], self.pad[1], self.sigma, orig.data_ptr() if orig is not None else None, flags,
                      self._workspace.data_ptr(), self._workspace.numel(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out

    def energies(self, theta: torch.Tensor, keep_images: bool = False) -> torch.Tensor:
        """mean squared gradient magnitude [P, K] (float64, device) of the blurred images of events warped with theta [P, K, 2]
        (pixels per patch time span, i.e. AFTER the reference's t_scale multiplication)."""
        return self._run(theta, None, None, keep_images)

    def images(self, n_candidates: int) -> torch.Tensor:
        """The blurred images [P, K, Hp, Wp] the last call with K = n_candidates and `keep_images=True` left in the workspace."""
        n = self.n_patches * n_candidates * self.padded_size[0] * self.padded_size[1]
        return self._workspace[: 4 * n].view(torch.float32).view(self.n_patches, n_candidates, *self.padded_size)

    def losses(self, candidates: torch.Tensor, keep_images: bool = False) -> torch.Tensor:
        """loss [P, K] (float64, DEVICE) of candidates [P, K, 2] as the sampler suggests them: one launch."""
        return self._run(candidates, self.theta_scale, self.orig_energy, keep_images)

    def evaluate(self, candidates, keep_images: bool = False) -> np.ndarray:
        """loss [P, K] (float64, host) of candidates [P, K, 2] = (trans_x, trans_y) as the sampler suggests them.  NaN losses are
        reported as 0.0 like the reference does (pyramid.py:374-375); rows of patches that are not `valid` are NaN.
        One H2D copy, one kernel, one D2H copy."""
        out = self.losses(torch.as_tensor(candidates, dtype=torch.float64), keep_images).cpu().numpy()
        out[~self.valid] = np.nan
        return out


def sampling_range(motion0_i: np.ndarray, abs_range: float = 10.0) -> np.ndarray:
    """[[low_x, high_x], [low_y, high_y]] of `sampling_initial` (pyramid.py:417-430)."""
    m = np.asarray(motion0_i, dtype=np.float64)
    c = np.stack([0.8 * m, m - abs_range, 1.2 * m, m + abs_range])
    return np.stack([c.min(0), c.max(0)], 1)


def run_patch_studies(evaluator: PatchCandidateEvaluator, motion0: np.ndarray, n_trials: int, n_startup_trials: int, *, optuna=None,
                      suggest=None) -> np.ndarray:
    """`initialize_guess_from_optuna_sampling` (pyramid.py:320-362) with the patch loop inside the trial loop.

    motion0: [2, P] initial guess; returns motion1 [2, P] (the best candidate of every valid patch, motion0 elsewhere).
    `suggest(trial, key, motion0_i)` defaults to the reference's `sampling_initial` rule."""
    if optuna is None:
        import optuna  # the reference's own dependency (src/solver/patch_contrast_pyramid.py:8)
    motion0 = np.asarray(motion0, dtype=np.float64).reshape(2, -1)
    P = evaluator.n_patches
    if motion0.shape[1] != P:
        raise ValueError(f"motion0 has {motion0.shape[1]} patches, the evaluator {P}")
    if suggest is None:
        def suggest(trial, key, m0):
            lo, hi = sampling_range(m0)[0 if key == "trans_x" else 1]
            return trial.suggest_uniform(key, lo, hi)
    live = [i for i in range(P) if evaluator.valid[i]]
    if not live:
        return motion0.copy()
    studies = {i: optuna.create_study(direction="minimize", sampler=optuna.samplers.TPESampler(n_startup_trials=n_startup_trials)) for i in live}
    cand = np.zeros((P, 1, 2))
    for _ in range(int(n_trials)):
        trials = {i: studies[i].ask() for i in live}
        for i, tr in trials.items():
            cand[i, 0, 0] = suggest(tr, "trans_x", motion0[:, i])
            cand[i, 0, 1] = suggest(tr, "trans_y", motion0[:, i])
        loss = evaluator.evaluate(cand)
        for i, tr in trials.items():
            studies[i].tell(tr, float(loss[i, 0]))
    motion1 = motion0.copy()
    for i in live:
        best = studies[i].best_params
        motion1[:, i] = (best["trans_x"], best["trans_y"])
    return motion1


def n_trials_at(n_iter: int, current_scale: int, coarsest_scale: int) -> int:
    """Optuna runs trials while `i_trial < n_trials`; the reference passes the float n_iter / (scale - coarsest) (pyramid.py:346-349)."""
    return int(math.ceil(n_iter / (current_scale - coarsest_scale)))
