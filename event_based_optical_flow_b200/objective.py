"""Fused contrast-maximization objective: one CM iteration = warp + IWE + cost + gradient on the GPU.

This is the fast path behind the reference's per-iteration seam
`PatchContrastMaximization.calculate_cost` / `get_arg_for_cost` (src/solver/patch_contrast_base.py:273-352):
events are made resident once per `optimize()` (`EventPlan`), and every objective evaluation is ONE C-ABI call
(`cmax_objective`: K1 -> image kernel -> K3; sharded over several GPUs `cmax_objective_sharded`, which does its two sums
over NVLink peer memory inside those kernels) with no host synchronisation.  The gradient w.r.t. the motion is analytic; events never
receive a gradient (the reference only ever asks for d cost / d motion, scipy_autograd/torch_wrapper.py:38-40).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

Direction = Union[str, float]

# cost name -> (statistic, form, (arg-dict key, direction) per fused reference time, multi-focal weights)
# src/costs/*.py and src/solver/patch_contrast_base.py:303-347: "iwe"/"backward_iwe" warp to the FIRST event,
# "forward_iwe" to the LAST, "middle_iwe" to the middle; multi-focal = N(fwd) + N(bwd) + 2 N(mid).
COST_TABLE = {
    "image_variance": ("variance", "plain", (("iwe", "first"),), (1.0,)),
    "gradient_magnitude": ("gradmag", "plain", (("iwe", "first"),), (1.0,)),
    "normalized_image_variance": ("variance", "normalized", (("iwe", "first"),), (1.0,)),
    "normalized_gradient_magnitude": ("gradmag", "normalized", (("iwe", "first"),), (1.0,)),
    "multi_focal_normalized_image_variance": (
        "variance", "multifocal", (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")), (1.0, 1.0, 2.0)),
    "multi_focal_normalized_gradient_magnitude": (
        "gradmag", "multifocal", (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")), (1.0, 1.0, 2.0)),
}


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: the B200 contrast-maximization path has no CPU fallback")


def _pad2(outer_padding) -> Tuple[int, int]:
    if isinstance(outer_padding, (int, float)):
        return int(outer_padding), int(outer_padding)
    return int(outer_padding[0]), int(outer_padding[1])


class EventPlan:
    """Resident, validated (and optionally source-pixel-ordered) float4 copy of one event batch.

    Events are constant during one `solver.optimize()` (src/solver/patch_contrast_pyramid.py:186), so the work the
    reference redoes on every call -- dtype conversion, `clone`, four min/max reductions over t (src/warp.py:201-259)
    -- is done once here.  `t_range` must be the GLOBAL (t_min, t_max) when the batch is one shard of a larger one.
    """

    def __init__(self, events: torch.Tensor, image_size: Tuple[int, int], outer_padding=0, order: str = "pixel",
                 t_range: Optional[Tuple[float, float]] = None):
        _require_cuda(events, "events")
        if events.dim() != 2 or events.shape[1] < 3:
            raise ValueError(f"events must be [n, >=3] (x=row, y=col, t, p); got {tuple(events.shape)}")
        if order not in _lib.ORDER:
            raise ValueError(f"order must be one of {list(_lib.ORDER)}, got {order}")
        self.lib = _lib.load()
        self.image_size = (int(image_size[0]), int(image_size[1]))
        self.pad = _pad2(outer_padding)
        self.padded_size = (self.image_size[0] + 2 * self.pad[0], self.image_size[1] + 2 * self.pad[1])
        self.device = events.device
        ev = events.detach()
        if ev.shape[1] == 3:
            ev = torch.cat([ev, ev.new_zeros(len(ev), 1)], dim=1)
        # keep the fp32 [n,4] array alive: an un-sorted plan borrows it
        self._events_f32 = ev[:, :4].to(torch.float32).contiguous()
        self.n = int(self._events_f32.shape[0])
        H, W = self.image_size
        with torch.cuda.device(self.device):
            nbytes = self.lib.cmax_plan_workspace_bytes(self.n, H, W, _lib.ORDER[order])
            if nbytes == 0:
                _lib.check("cmax_plan_workspace_bytes", _lib.ERR_ARG if self.n >= 0 else _lib.ERR_CUDA)
            self._workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            ws_ptr = (self._workspace.data_ptr() + 255) // 256 * 256
            tmin, tmax = (float("nan"), float("nan")) if t_range is None else (float(t_range[0]), float(t_range[1]))
            handle = C.c_void_p()
            _lib.call("cmax_plan_create", C.byref(handle), self._events_f32.data_ptr(), self.n, 4, H, W, self.pad[0], self.pad[1],
                      tmin, tmax, _lib.ORDER[order], ws_ptr, nbytes, _stream_ptr())
        self._handle = handle
        a, b, n, o = C.c_float(), C.c_float(), C.c_int64(), C.c_int32()
        _lib.call("cmax_plan_info", self._handle, C.byref(a), C.byref(b), C.byref(n), C.byref(o))
        self.t_min, self.t_max = a.value, b.value
        self.order = {v: k for k, v in _lib.ORDER.items()}[o.value]
        self.refs: Tuple[Direction, ...] = ("first",)
        self.n_bins = 0
        if self.order != "asis":
            self._events_f32 = None  # the plan owns a re-ordered copy inside its workspace

    @property
    def handle(self):
        if self._handle is None:
            raise RuntimeError("EventPlan already closed")
        return self._handle

    @property
    def n_strips(self) -> int:
        """Strips the plan cut the batch into; 0 = the strip kernels are not available for this batch (see cmax_plan_strips)."""
        n = C.c_int64(0)
        _lib.call("cmax_plan_strips", self.handle, C.byref(n))
        return int(n.value)

    def set_refs(self, directions: Sequence[Direction], n_bins: int = 0) -> None:
        directions = tuple(directions)
        if (directions, n_bins) == (self.refs, self.n_bins):
            return
        arr = _lib.refs_array(directions)
        with torch.cuda.device(self.device):
            _lib.call("cmax_plan_set_refs", self.handle, arr, len(directions), int(n_bins), _stream_ptr())
        self.refs, self.n_bins = directions, int(n_bins)

    def set_variant(self, vote_variant: int = 0, grad_variant: int = 0) -> None:
        _lib.call("cmax_plan_set_variant", self.handle, int(vote_variant), int(grad_variant))

    def set_compact(self, enable: bool = True) -> bool:
        """Choose the packed-event format (8-byte compact when the batch allows it, else / or forced 16-byte); returns
        whether the compact format is in use.  See cmax_plan_set_compact in include/cmax_b200.h."""
        out = C.c_int32(0)
        with torch.cuda.device(self.device):
            _lib.call("cmax_plan_set_compact", self.handle, 1 if enable else 0, C.byref(out), _stream_ptr())
        return bool(out.value)

    def set_stage_mask(self, mask: int = 7) -> None:
        """Measurement builds only (-DCMAX_MEASURE): the release library accepts no mask but 7."""
        _lib.call("cmax_plan_set_stage_mask", self.handle, int(mask))

    def close(self) -> None:
        if getattr(self, "_handle", None) is not None:
            self.lib.cmax_plan_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _motion_shape(model: str, image_size, n_bins: int) -> Tuple[int, ...]:
    H, W = image_size
    if model == "dense-flow":
        return (2, H, W)
    if model == "dense-flow-voxel":
        return (n_bins, 2, H, W)
    return (2,)


class ContrastObjective:
    """cost(motion) and d cost / d motion for one resident event batch.

    Args mirror the reference's configuration vocabulary: `motion_model` as in `Warp.warp_event`
    (src/warp.py:156-199), `cost` a key of `costs.functions` (src/costs/__init__.py:35), `sigma` = `iwe.blur_sigma`,
    `direction` = the cost direction (src/costs/base.py:20-25), `outer_padding` as in `EventImageConverter`.
    `process_group`: events are sharded over its ranks; the partial IWE and the partial gradient are summed once each per
    evaluation, either with an NCCL all-reduce (`exchange="nccl"`) or over NVLink peer memory
    (`exchange="peer"`: the workspaces live in torch symmetric memory and the sums happen INSIDE the image kernel and a
    gradient-exchange kernel, behind flags raised on the peers -- no collective launch, no barrier kernel; see
    cmax_objective_sharded in include/cmax_b200.h).
    """

    def __init__(self, events: Union[torch.Tensor, EventPlan], image_size: Tuple[int, int], *, cost: str = "image_variance",
                 motion_model: str = "dense-flow", sigma: float = 0.0, omit_boundary: bool = True, direction: str = "minimize",
                 outer_padding=0, n_bins: Optional[int] = None, order: str = "pixel", process_group=None,
                 t_range: Optional[Tuple[float, float]] = None, orig_events: Optional[torch.Tensor] = None,
                 exchange: str = "nccl", cuda_graph: bool = False):
        if cost not in COST_TABLE:
            raise KeyError(f"cost {cost!r} has no fused CUDA form; available: {sorted(COST_TABLE)}")
        if motion_model not in _lib.MOTION or motion_model == "tile-flow":  # (the tile-flow model is reached through TileFlowObjective)
            from .warp import MotionModelKeyError
            raise MotionModelKeyError(motion_model)
        if direction not in ("minimize", "maximize", "natural"):
            raise ValueError(f"direction should be minimize, maximize, and natural. Got {direction}.")
        self.cost = cost
        self.motion_model = motion_model
        stat, form, refs, weights = COST_TABLE[cost]
        self.stat, self.form = stat, form
        self.ref_keys = tuple(k for k, _ in refs)
        self.group = process_group
        if exchange not in ("nccl", "peer"):
            raise ValueError(f"exchange must be 'nccl' or 'peer', got {exchange}")
        self.exchange = exchange if process_group is not None else "nccl"
        if orig_events is None and isinstance(events, torch.Tensor):
            orig_events = events
        # second order (Hessian-vector products) runs on the modular operators, which take the caller's event array
        # (an objective built from an EventPlan keeps the caller's tensor through `orig_events`)
        self._events_ref = events.detach() if isinstance(events, torch.Tensor) else (orig_events.detach() if orig_events is not None else None)
        self._modular_tp = None
        self.plan = events if isinstance(events, EventPlan) else EventPlan(events, image_size, outer_padding, order, t_range)
        self.device = self.plan.device
        self.image_size = self.plan.image_size
        self.padded_size = self.plan.padded_size
        self.n_bins = int(n_bins) if motion_model == "dense-flow-voxel" else 0
        if motion_model == "dense-flow-voxel" and not (1 <= self.n_bins <= _lib.MAX_BINS):
            raise ValueError(f"dense-flow-voxel needs 1 <= n_bins <= {_lib.MAX_BINS}")
        self.directions = tuple(d for _, d in refs)
        self.plan.set_refs(self.directions, self.n_bins)
        # Sign conventions of src/costs/*.py.  plain: minimize -> -stat, maximize/natural -> +stat.
        # normalised: minimize -> orig/warped, maximize/natural -> warped/orig.  multi-focal: minimize -> sum of
        # orig/warped, maximize -> -(sum of warped/orig), natural -> +(sum of warped/orig) (the inner normalised cost
        # is built with the same direction, multi_focal_normalized_*.py:36-38, :96-101).
        sign = 1 if direction == "minimize" else -1
        self._post_sign = -1.0 if (direction == "natural" and form == "multifocal") else 1.0
        self.spec = _lib.CostSpec(_lib.STAT[stat], _lib.FORM[form], sign, 1 if omit_boundary else 0, float(sigma),
                                  (C.c_float * _lib.MAX_REFS)(*(list(weights) + [0.0] * (_lib.MAX_REFS - len(weights)))))
        self.sigma = float(sigma)
        self.omit_boundary = bool(omit_boundary)
        self.motion_shape = _motion_shape(motion_model, self.image_size, self.n_bins)
        self.lib = _lib.load()
        self._symm = None
        with torch.cuda.device(self.device):
            nbytes = self.lib.cmax_objective_workspace_bytes(self.plan.handle, C.byref(self.spec))
            if self.exchange == "peer":
                self._setup_peer_exchange(nbytes)
            else:
                self._ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._ws_ptr = (self._ws.data_ptr() + 255) // 256 * 256
        with torch.cuda.device(self.device):
            _lib.call("cmax_objective_workspace_init", self.plan.handle, self._ws_ptr, _stream_ptr())
        self._cost = torch.zeros(1, dtype=torch.float64, device=self.device)
        # small batches (the shipped YAMLs' 30 k events) are launch / host bound: with `cuda_graph` every evaluation replays a
        # graph captured once per (value | value + gradient) into static buffers -- one launch, no per-call ctypes traffic
        self.cuda_graph = bool(cuda_graph) and process_group is None
        self._graphs: dict = {}
        self._orig_stat = None
        if form != "plain":
            self._orig_stat = self._orig_statistic(orig_events)
        self._iwe_view = None

    # -- NVLink peer-memory exchange: workspace, partial gradient and flag block are ONE symmetric allocation per rank
    def _setup_peer_exchange(self, nbytes: int) -> None:
        """[objective workspace][partial gradient][flag block]; `cmax_peers` holds every rank's addresses of the three."""
        import torch.distributed._symmetric_memory as symm_mem
        world = torch.distributed.get_world_size(self.group)
        rank = torch.distributed.get_rank(self.group)
        if world > _lib.MAX_PEERS:
            raise ValueError(f"exchange='peer' supports up to {_lib.MAX_PEERS} ranks (one NVLink domain), got {world}")
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                symm_mem.enable_symm_mem_for_group(self.group.group_name)
            except Exception:
                pass  # newer torch enables it implicitly
        n_motion = int(np.prod(self.motion_shape))
        al = lambda v: (v + 255) // 256 * 256  # noqa: E731
        grad_off = al(nbytes)
        flags_off = al(grad_off + 4 * n_motion)
        total = flags_off + 256
        self._ws = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        if self._ws.data_ptr() % 256 != 0:
            raise RuntimeError("symmetric allocation is not 256-byte aligned")
        self._ws.zero_()
        self._symm = symm_mem.rendezvous(self._ws, self.group)
        bases = [int(p) for p in self._symm.buffer_ptrs]
        iwe_off = int(self.lib.cmax_objective_iwe_offset(self.plan.handle))
        self._grad_part = self._ws[grad_off:grad_off + 4 * n_motion].view(torch.float32).view(self.motion_shape)
        peers = _lib.Peers()
        peers.n_peers, peers.rank = world, rank
        for r, b in enumerate(bases):
            peers.iwe[r], peers.grad[r], peers.flags[r] = b + iwe_off, b + grad_off, b + flags_off
        self._peers = peers
        torch.cuda.synchronize(self.device)
        torch.distributed.barrier(group=self.group)  # every rank's flags are zero before anyone raises one

    # -- the statistic of the un-warped IWE (normalised costs); constant per optimize(), so computed once
    #    (the reference recomputes it on every call, src/solver/patch_contrast_base.py:295-301)
    def _orig_statistic(self, orig_events: Optional[torch.Tensor]) -> torch.Tensor:
        from . import ops
        Hp, Wp = self.padded_size
        if orig_events is None:
            raise ValueError("normalised / multi-focal costs need `orig_events` (the un-warped events of this rank)")
        img = ops.vote(orig_events.detach().to(torch.float32), self.padded_size, self.plan.pad, None, "bilinear_vote")
        if self.group is not None:
            torch.distributed.all_reduce(img, group=self.group)
        if self.sigma > 0:
            img = ops.blur3(img[None], self.sigma)[0]
        # NormalizedImageVariance crops iwe but NOT orig_iwe (src/costs/normalized_image_variance.py:38-41)
        omit = self.omit_boundary and self.stat == "gradmag"
        stats, _ = ops.image_stats(img[None], self.stat, omit, want_grad=False)
        return stats[0, :1].clone()

    # -- one evaluation
    def _check_motion(self, motion: torch.Tensor) -> torch.Tensor:
        _require_cuda(motion, "motion")
        if tuple(motion.shape) != self.motion_shape:
            raise ValueError(f"motion for {self.motion_model} must have shape {self.motion_shape}, got {tuple(motion.shape)}")
        return motion.detach().to(torch.float32).contiguous()

    def _vote_and_fold(self, m: torch.Tensor, stream: int, model: Optional[int] = None) -> None:
        """K1 + fold: this rank's (partial) IWE stack is in `self._iwe_view` afterwards."""
        iwe_ptr = C.c_void_p()
        _lib.call("cmax_objective_vote", self.plan.handle, _lib.MOTION[self.motion_model] if model is None else model, m.data_ptr(),
                  self._ws_ptr, stream)
        _lib.call("cmax_objective_fold", self.plan.handle, self._ws_ptr, C.byref(iwe_ptr), stream)
        if self._iwe_view is None:
            off = iwe_ptr.value - self._ws.data_ptr()
            Hp, Wp = self.padded_size
            k = len(self.directions)
            self._iwe_view = self._ws[off:off + 4 * k * Hp * Wp].view(torch.float32).view(k, Hp, Wp)

    def _evaluate(self, m: torch.Tensor, cost: torch.Tensor, grad: Optional[torch.Tensor], stream: int, motion_model: Optional[str] = None) -> None:
        """One evaluation: writes cost[0] and (if given) grad.  No host sync, CUDA-graph capturable.  `motion_model` overrides
        the objective's own (TileFlowObjective evaluates a dense-flow objective with the fused "tile-flow" model)."""
        model = _lib.MOTION[motion_model or self.motion_model]
        orig = self._orig_stat.data_ptr() if self._orig_stat is not None else None
        gptr = grad.data_ptr() if grad is not None else None
        self.plan.set_refs(self.directions, self.n_bins)  # no-op unless another objective re-packed the shared plan
        if self.group is None:  # single GPU: K1 -> image kernel -> K3
            _lib.call("cmax_objective", self.plan.handle, model, m.data_ptr(), C.byref(self.spec), orig, self._ws_ptr, cost.data_ptr(), gptr, stream)
        elif self.exchange == "peer":  # sharded, both sums inside the kernels over NVLink peer memory
            _lib.call("cmax_objective_sharded", self.plan.handle, model, m.data_ptr(), C.byref(self.spec), orig, self._ws_ptr,
                      C.byref(self._peers), cost.data_ptr(), gptr, stream)
        else:  # sharded, NCCL all-reduces between the stages (the baseline exchange)
            self._vote_and_fold(m, stream, model)
            torch.distributed.all_reduce(self._iwe_view, group=self.group)
            _lib.call("cmax_objective_cost", self.plan.handle, C.byref(self.spec), orig, self._ws_ptr, 1 if grad is not None else 0,
                      cost.data_ptr(), gptr, grad.numel() if grad is not None else 0, stream)
            if grad is not None:
                _lib.call("cmax_objective_grad", self.plan.handle, model, m.data_ptr(), self._ws_ptr, gptr, 1, stream)
                torch.distributed.all_reduce(grad, group=self.group)

    def value_and_grad(self, motion: torch.Tensor, want_grad: bool = True):
        """-> (cost: 0-dim float64 CUDA tensor, grad: fp32 tensor shaped like motion or None).  No host sync."""
        m = self._check_motion(motion)
        with torch.cuda.device(self.device):
            if self.cuda_graph:
                cost, grad = self._replay(m, want_grad)
            else:
                cost = torch.empty(1, dtype=torch.float64, device=self.device)
                grad = torch.empty(self.motion_shape, dtype=torch.float32, device=self.device) if want_grad else None
                self._evaluate(m, cost, grad, _stream_ptr())
        if self._post_sign < 0:
            cost = -cost
            grad = -grad if grad is not None else None
        return cost[0], grad

    def _replay(self, m: torch.Tensor, want_grad: bool):
        """Graph-cached evaluation: capture `_evaluate` once per variant into static buffers, then copy-in / replay / copy-out."""
        entry = self._graphs.get(want_grad)
        if entry is None:
            sm = torch.empty_like(m)
            sc = torch.empty(1, dtype=torch.float64, device=self.device)
            sg = torch.empty(self.motion_shape, dtype=torch.float32, device=self.device) if want_grad else None
            sm.copy_(m)
            self.plan.set_refs(self.directions, self.n_bins)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside the capture (first-use initialisation of the kernels)
                for _ in range(2):
                    self._evaluate(sm, sc, sg, side.cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._evaluate(sm, sc, sg, _stream_ptr())
            entry = (graph, sm, sc, sg)
            self._graphs[want_grad] = entry
        graph, sm, sc, sg = entry
        self.plan.set_refs(self.directions, self.n_bins)  # (no-op unless another objective re-packed the shared plan)
        sm.copy_(m)
        graph.replay()
        return sc.clone(), (sg.clone() if want_grad else None)

    def value(self, motion: torch.Tensor) -> torch.Tensor:
        return self.value_and_grad(motion, want_grad=False)[0]

    def iwe(self, motion: torch.Tensor) -> torch.Tensor:
        """The (un-blurred) IWE stack [n_ref, Hp, Wp] for `motion` (a copy; summed over all ranks when sharded)."""
        m = self._check_motion(motion)
        with torch.cuda.device(self.device):
            self.plan.set_refs(self.directions, self.n_bins)
            self._vote_and_fold(m, _stream_ptr())
            out = self._iwe_view.clone()
            if self.group is not None:
                torch.distributed.all_reduce(out, group=self.group)  # off the hot path: plain NCCL
        return out

    def step_into(self, motion_f32: torch.Tensor, cost_out: torch.Tensor, grad_out: torch.Tensor) -> None:
        """Allocation-free evaluation into caller buffers (fp32 motion, float64[1] cost, fp32 gradient), CUDA-graph
        capturable (also sharded: the flag exchange lives inside the kernels, and NCCL collectives are capturable)."""
        if self._post_sign < 0:
            raise RuntimeError("step_into supports the minimize / maximize directions")
        self._evaluate(motion_f32, cost_out, grad_out, _stream_ptr())

    # -- second order: the same cost composed from the modular operators (warp -> vote -> blur -> statistic, one C-ABI
    #    call each, every backward differentiable once more), so that torch can differentiate the gradient itself
    def modular_cost(self, motion: torch.Tensor) -> torch.Tensor:
        """The cost as a twice-differentiable function of `motion` (same value as `value(motion)` up to fp32 summation
        order).  Composition and sign conventions of src/solver/patch_contrast_base.py:289-352 and src/costs/*.py.
        Sharded objectives: a collective call (every rank, same motion)."""
        from . import ops
        if self._events_ref is None:
            raise NotImplementedError("Hessian-vector products need the event tensor: pass `orig_events=` when the objective is built from an EventPlan")
        ev = self._events_ref
        if ev.shape[1] == 3:
            ev = torch.cat([ev, ev.new_zeros(len(ev), 1)], dim=1)
        ev = ev[:, :4].to(torch.float32)
        if self._modular_tp is None:  # (sharded: reference time / period / bin edges from the GLOBAL time range, like the plan)
            self._modular_tp = ops.time_params(ev, self.directions, self.n_bins, t_range=(self.plan.t_min, self.plan.t_max))
        # sharded: every rank warps and votes ITS events; the partial IWEs are summed over the ranks (ops.SumPartials) and the
        # replicated motion is put to local use (ops.UseReplicated) -- the two are each other's adjoint, so the recorded backward
        # and its own backward (the Hessian-vector product) come out right on every rank
        m_local = ops.UseReplicated.apply(motion, self.group) if self.group is not None else motion
        stats = []
        for r in range(len(self.directions)):
            warped = ops.WarpFunction.apply(ev, m_local, self.motion_model, self.image_size, self._modular_tp, r)
            iwe = ops.VoteFunction.apply(warped, None, self.padded_size, self.plan.pad, "bilinear_vote")
            if self.group is not None:
                iwe = ops.SumPartials.apply(iwe, self.group)
            if self.sigma > 0:
                iwe = ops.BlurFunction.apply(iwe, self.sigma)
            stats.append(ops.ImageStatFunction.apply(iwe, self.stat, self.omit_boundary).double())
        sign = int(self.spec.direction_sign)
        if self.form == "plain":
            total = -sign * stats[0]
        else:
            orig = self._orig_stat[0]
            weights = [float(self.spec.weights[r]) if self.form == "multifocal" else 1.0 for r in range(len(stats))]
            if sign > 0:
                total = sum(w * orig / c for w, c in zip(weights, stats))
            elif self.form == "normalized":
                total = stats[0] / orig
            else:
                total = -sum(w * c / orig for w, c in zip(weights, stats))
        return self._post_sign * total

    def differentiable_grad(self, motion: torch.Tensor) -> torch.Tensor:
        """d cost / d motion as a differentiable function of `motion` (for double backward)."""
        with torch.enable_grad():
            cost = self.modular_cost(motion)
            (grad,) = torch.autograd.grad(cost, motion, create_graph=True)
        return grad.to(motion.dtype)

    def hvp(self, motion: torch.Tensor, vector: torch.Tensor) -> torch.Tensor:
        """Hessian-vector product d/d eps grad cost(motion + eps * vector) at eps = 0 (what scipy's Newton-CG asks for)."""
        _require_cuda(motion, "motion")
        m = motion.detach().clone().requires_grad_(True)
        grad = self.differentiable_grad(m)
        (hv,) = torch.autograd.grad(grad, m, vector.to(grad.dtype))
        return hv

    def __call__(self, motion: torch.Tensor) -> torch.Tensor:
        """Autograd-aware scalar in motion's dtype: drop-in for the reference's `calculate_cost` result."""
        return _ObjectiveFunction.apply(motion, self)


class TileFlowObjective:
    """cost(patch motion [2,hp,wp]) with the tile-flow -> dense-flow map of the reference in front of a dense-flow
    `ContrastObjective` (what `objective_scipy` evaluates, src/solver/patch_contrast_pyramid.py:430-462: dense =
    interpolate(motion) * t_scale -> calculate_cost).  The gradient comes back on the patch grid (hp*wp*2 numbers), so
    the host round trip per optimiser step is a few KB.

    `fused=True`: the event kernels evaluate the map at every source pixel themselves (motion model "tile-flow",
    cmax_plan_set_tile_flow) -- no dense [2,H,W] flow or gradient exists, a CM iteration stays at three launches, and a sharded
    objective exchanges 2*hp*wp floats instead of a dense gradient.  `fused=False`: the up-sampling kernel, the dense objective
    and the adjoint kernel are composed (same numbers: both evaluate the same expression).  Default: fused for SHARDED
    objectives whose plan has strips (the 2 KB gradient exchange is what pays), composed otherwise -- on one GPU the per-strip
    evaluation of the grid and the node reductions cost K1 / K3 slightly more (49.1 us) than the two small extra kernels of the
    composition (46.4 us at config 2; bench.py `tile_flow`)."""

    def __init__(self, objective: ContrastObjective, patch_size, sliding_window, patch_shift=(0, 0), t_scale: float = 1.0,
                 fused: Optional[bool] = None):
        from . import ops
        if objective.motion_model != "dense-flow":
            raise ValueError("TileFlowObjective wraps a dense-flow ContrastObjective")
        self.objective = objective
        self.image_shape = objective.image_size
        self.window = (int(sliding_window[0]), int(sliding_window[1]))
        self.pad = ops.tile_flow_geometry(self.image_shape, patch_size, sliding_window, patch_shift)
        self.t_scale = float(t_scale)
        can_fuse = objective.plan.n_strips > 0 and objective.plan.n > 0
        if fused and not can_fuse:
            raise ValueError("the fused tile-flow model needs a plan with strips (a pixel-ordered batch dense enough to be cut into strips)")
        self.fused = (can_fuse and objective.group is not None) if fused is None else bool(fused)

    def _set_geometry(self, grid) -> None:
        """The geometry lives in the PLAN (host-side state of the C library), and several TileFlowObjectives may share one
        plan: compare with what the plan currently holds, not with what this object set last."""
        geom = (int(grid[0]), int(grid[1]), int(self.pad[0]), int(self.pad[1]), self.window[0], self.window[1], self.t_scale)
        plan = self.objective.plan
        if geom != getattr(plan, "_tile_geom", None):
            if geom[0] * geom[1] > 1024:
                raise ValueError(f"the fused tile-flow model supports patch grids of up to 1024 nodes, got {geom[0]}x{geom[1]}")
            _lib.call("cmax_plan_set_tile_flow", plan.handle, *geom)
            plan._tile_geom = geom

    def _check(self, motion: torch.Tensor) -> None:
        _require_cuda(motion, "motion")
        if motion.dim() != 3 or motion.shape[0] != 2:
            raise ValueError(f"tile-flow motion must be [2,hp,wp], got {tuple(motion.shape)}")

    def value_and_grad(self, motion: torch.Tensor, want_grad: bool = True):
        from . import ops
        self._check(motion)
        obj = self.objective
        if self.fused:
            m = motion.detach().to(torch.float32).contiguous()
            with torch.cuda.device(obj.device):
                cost = torch.empty(1, dtype=torch.float64, device=obj.device)
                grad = torch.empty_like(m) if want_grad else None
                self._set_geometry(m.shape[-2:])
                obj._evaluate(m, cost, grad, _stream_ptr(), motion_model="tile-flow")
            if obj._post_sign < 0:
                cost = -cost
                grad = -grad if grad is not None else None
            return cost[0], grad
        dense = ops.tile_flow_upsample(motion, self.image_shape, self.pad, self.window)
        if self.t_scale != 1.0:
            dense = dense * self.t_scale
        cost, gdense = obj.value_and_grad(dense, want_grad)
        if not want_grad:
            return cost, None
        gm = ops.tile_flow_upsample_backward(gdense, motion.shape[-2:], self.pad, self.window)
        if self.t_scale != 1.0:
            gm = gm * self.t_scale
        return cost, gm

    def value(self, motion: torch.Tensor) -> torch.Tensor:
        return self.value_and_grad(motion, want_grad=False)[0]

    def step_into(self, motion_f32: torch.Tensor, cost_out: torch.Tensor, grad_out: torch.Tensor) -> None:
        """Evaluation into caller buffers (fp32 motion [2,hp,wp], float64[1] cost, fp32 gradient); CUDA-graph capturable.
        Fused: allocation-free, three launches."""
        if self.fused and self.objective._post_sign > 0:
            self._set_geometry(motion_f32.shape[-2:])
            self.objective._evaluate(motion_f32, cost_out, grad_out, _stream_ptr(), motion_model="tile-flow")
            return
        cost, gm = self.value_and_grad(motion_f32)
        cost_out.copy_(cost.reshape(1))
        grad_out.copy_(gm)

    # -- second order: the differentiable composition of the modular operators (tile-flow up-sampling is linear)
    def modular_cost(self, motion: torch.Tensor) -> torch.Tensor:
        from . import ops
        dense = ops.TileFlowFunction.apply(motion, tuple(self.image_shape), tuple(self.pad), tuple(self.window))
        if self.t_scale != 1.0:
            dense = dense * self.t_scale
        return self.objective.modular_cost(dense)

    def differentiable_grad(self, motion: torch.Tensor) -> torch.Tensor:
        with torch.enable_grad():
            cost = self.modular_cost(motion)
            (grad,) = torch.autograd.grad(cost, motion, create_graph=True)
        return grad.to(motion.dtype)

    def __call__(self, motion: torch.Tensor) -> torch.Tensor:
        """Autograd-aware scalar in motion's dtype (first order: the fused kernels' analytic gradient; under
        `create_graph=True` the gradient is rebuilt from the twice-differentiable operators, as for ContrastObjective)."""
        return _ObjectiveFunction.apply(motion, self)


class TimeAwareObjective:
    """cost(motion) for the time-aware configuration: motion -> [tile-flow upsample] -> dense flow at t0 -> flow voxel
    (upwind / Burgers propagation, src/utils/flow_utils.py:99-161) -> voxel warp + IWE + cost, and the gradient back
    through every stage -- what `TimeAwarePatchContrastMaximization.objective_scipy` evaluates
    (src/solver/time_aware_patch_contrast.py:42-80, src/solver/patch_contrast_pyramid.py:430-462), every stage a CUDA
    kernel of this library.  `objective` is a `ContrastObjective(motion_model="dense-flow-voxel", n_bins=time_bin)`;
    `tile` = dict(patch_size, sliding_window, patch_shift) makes the motion a [2,hp,wp] patch grid, None a dense [2,H,W]
    flow.  `scale_later` as in the reference: the voxel is built from motion / max(motion) and scaled back."""

    def __init__(self, objective: ContrastObjective, scheme: str = "burgers", t0_location: str = "middle", tile: Optional[dict] = None,
                 scale_later: bool = False):
        from . import ops
        if objective.motion_model != "dense-flow-voxel":
            raise ValueError("TimeAwareObjective wraps a dense-flow-voxel ContrastObjective")
        ops._voxel_args(scheme, t0_location)
        self.objective = objective
        self.time_bin = objective.n_bins
        self.scheme, self.t0_location = scheme, t0_location
        self.image_shape = objective.image_size
        self.scale_later = bool(scale_later)
        self.tile = None
        if tile is not None:
            window = (int(tile["sliding_window"][0]), int(tile["sliding_window"][1]))
            self.tile = (ops.tile_flow_geometry(self.image_shape, tile["patch_size"], window, tile.get("patch_shift", (0, 0))), window)

    def value_and_grad(self, motion: torch.Tensor, want_grad: bool = True):
        from . import ops
        _require_cuda(motion, "motion")
        m = motion.detach().to(torch.float32)
        dense = ops.tile_flow_upsample(m, self.image_shape, *self.tile) if self.tile is not None else m.contiguous()
        scale = m.max() if self.scale_later else None
        d_in = dense / scale if scale is not None else dense
        voxel = ops.flow_voxel(d_in, self.time_bin, self.scheme, self.t0_location)
        v_in = voxel * scale if scale is not None else voxel
        cost, gvox = self.objective.value_and_grad(v_in, want_grad)
        if not want_grad:
            return cost, None
        g_d_in = ops.flow_voxel_backward(d_in, voxel, gvox * scale if scale is not None else gvox, self.scheme, self.t0_location)
        gdense = g_d_in / scale if scale is not None else g_d_in
        gm = ops.tile_flow_upsample_backward(gdense, m.shape[-2:], *self.tile) if self.tile is not None else gdense
        if scale is not None:
            # d/d scale of (voxel(dense/scale) * scale), routed to the arg-max element of the motion (torch's max backward)
            gs = (gvox * voxel).sum() - (g_d_in * d_in).sum() / scale
            flat = gm.reshape(-1)
            flat[m.reshape(-1).argmax()] += gs
        return cost, gm

    def value(self, motion: torch.Tensor) -> torch.Tensor:
        return self.value_and_grad(motion, want_grad=False)[0]

    def step_into(self, motion_f32: torch.Tensor, cost_out: torch.Tensor, grad_out: torch.Tensor) -> None:
        """Evaluation into caller buffers (float64[1] cost, fp32 gradient shaped like the motion); CUDA-graph capturable
        (the intermediate flow / voxel tensors come from the graph's private pool)."""
        cost, gm = self.value_and_grad(motion_f32)
        cost_out.copy_(cost.reshape(1))
        grad_out.copy_(gm)


class _ObjectiveFunction(torch.autograd.Function):
    """cost(motion) on the fused kernels.  First order (every scipy / torch optimiser except the Newton family): the
    backward hands out the analytic gradient the forward already computed.  Second order: when the backward itself is
    being recorded (`create_graph=True`, which is how `torch.autograd.functional.vhp` obtains the Hessian-vector product
    for Newton-CG / trust-*, scipy_autograd/torch_wrapper.py:51-73), the gradient is rebuilt as a differentiable function of
    `motion` from the modular CUDA operators, every one of which has a second-order kernel (ops.py)."""

    @staticmethod
    def forward(ctx, motion: torch.Tensor, obj: ContrastObjective):
        need = ctx.needs_input_grad[0]
        cost, grad = obj.value_and_grad(motion, want_grad=need)
        ctx.grad = grad
        ctx.obj = obj
        ctx.dtype = motion.dtype
        ctx.save_for_backward(motion)
        return cost.to(motion.dtype if motion.dtype.is_floating_point else torch.float32)

    @staticmethod
    def backward(ctx, g):
        if ctx.grad is None:
            return None, None
        if torch.is_grad_enabled():
            (motion,) = ctx.saved_tensors
            return ctx.obj.differentiable_grad(motion) * g, None
        return (ctx.grad.to(ctx.dtype) * g.to(ctx.dtype)), None


def cm_objective(events: torch.Tensor, motion: torch.Tensor, image_size: Tuple[int, int], **kw) -> torch.Tensor:
    """One-shot functional form (builds a plan every call; prefer a cached `ContrastObjective` inside a solver)."""
    kw.setdefault("orig_events", events)
    obj = ContrastObjective(events, image_size, **kw)
    return obj(motion)
