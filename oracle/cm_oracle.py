"""CPU oracle for the contrast-maximization hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU ops, any float dtype) of the
reference's *torch branch* for: reference-time / dt normalisation, the three event
warps, the bilinear-vote / count IWE, the 3x3 Gaussian blur, and the contrast costs,
plus closed-form analytic gradients.  It exists only so that `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs can
check and time the CUDA path against it.  Nothing under `event_based_optical_flow_b200/`
may import it; the product path fails loudly when the CUDA library is missing.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the unmodified reference from
/root/reference in the build container, runs it on seeded inputs and commits the outputs
under `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function here
against those vectors and against the hand-computed vectors of the reference's own tests
(tests/test_warp.py:96-195, tests/test_event_image_converter.py:17-110).

Conventions (reference: src/utils/event_utils.py:38, src/event_image_converter.py:344-345):
an event is (x, y, t, p) with x = ROW (height) and y = COLUMN (width); flow is [2, H, W]
with channel 0 the row component; flat pixel index = x * W + y.

Each function cites the reference lines it follows.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

Direction = Union[str, float]

_DIRECTION_FRACTION = {"middle": 0.5, "before": -1.0, "after": 2.0}


# --------------------------------------------------------------------------------------
# reference time and dt   (src/warp.py:201-259)
# --------------------------------------------------------------------------------------
def reference_time(t: torch.Tensor, direction: Direction) -> torch.Tensor:
    """0-dim tensor in t.dtype.  src/warp.py:201-233."""
    t_lo, t_hi = t.min(), t.max()
    if isinstance(direction, float):
        return t_lo + (t_hi - t_lo) * direction
    if direction == "first":
        return t_lo
    if direction == "last":
        return t_hi
    if direction in _DIRECTION_FRACTION:
        return t_lo + (t_hi - t_lo) * _DIRECTION_FRACTION[direction]
    raise ValueError(f"direction must be first/middle/last/before/after or float, got {direction}")


def normalised_dt(t: torch.Tensor, ref: torch.Tensor, normalize_t: bool = True,
                  period: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dt = t - ref, divided by (max dt - min dt) of the batch.  src/warp.py:254-259."""
    dt = t - ref
    if normalize_t:
        if period is None:
            period = dt.max() - dt.min()
        dt = dt / period
    return dt


def ref_and_period(t: torch.Tensor, direction: Direction) -> Tuple[float, float]:
    """(ref, period) as python floats holding values representable in t.dtype."""
    ref = reference_time(t, direction)
    dt = t - ref
    return float(ref), float(dt.max() - dt.min())


# --------------------------------------------------------------------------------------
# warps   (src/warp.py:301-313, 339-365, 506-520)
# --------------------------------------------------------------------------------------
def _source_index(events: torch.Tensor, W: int) -> torch.Tensor:
    # truncation toward zero of the UN-warped coordinates, src/warp.py:305
    return events[:, 0].long() * W + events[:, 1].long()


def warp_dense(events: torch.Tensor, flow: torch.Tensor, direction: Direction = "first",
               normalize_t: bool = True) -> torch.Tensor:
    """events [n,>=3], flow [2,H,W] -> warped [n,C] = (x - dt*f0[src], y - dt*f1[src], dt, p).

    The product dt*f is rounded before the subtraction (two roundings, no fma).  src/warp.py:301-313.
    """
    W = flow.shape[-1]
    dt = normalised_dt(events[:, 2], reference_time(events[:, 2], direction), normalize_t)
    src = _source_index(events, W)
    f = flow.reshape(2, -1)
    out = events.clone()
    out[:, 0] = events[:, 0] - dt * f[0][src]
    out[:, 1] = events[:, 1] - dt * f[1][src]
    out[:, 2] = dt
    return out


def voxel_bin_edges(dt_min: float, dt_max: float, n_bins: int) -> np.ndarray:
    """float64 edges, last edge = dt_max + 1000.  src/warp.py:342-345."""
    edges = np.arange(0, n_bins) / n_bins * (dt_max - dt_min) + dt_min
    return np.append(edges, dt_max + 1e3)


def voxel_bin_of(dt: torch.Tensor, n_bins: int) -> torch.Tensor:
    """Bin index per event (or -1 if in no bin): edges[b] <= dt < edges[b+1], the scalar edge
    being rounded to dt.dtype before the comparison (torch scalar semantics).  src/warp.py:346-352."""
    edges = voxel_bin_edges(dt.min().item(), dt.max().item(), n_bins)
    edges_t = torch.tensor(edges, dtype=torch.float64).to(dt.dtype)
    b = torch.full(dt.shape, -1, dtype=torch.long)
    for k in range(n_bins):
        b[(edges_t[k] <= dt) & (dt < edges_t[k + 1])] = k
    return b


def warp_voxel(events: torch.Tensor, voxel: torch.Tensor, direction: Direction = "first",
               normalize_t: bool = True) -> torch.Tensor:
    """Time-aware warp: voxel [T,2,H,W]; an event in time bin b uses voxel[b].  src/warp.py:339-365."""
    T, _, H, W = voxel.shape
    dt = normalised_dt(events[:, 2], reference_time(events[:, 2], direction), normalize_t)
    b = voxel_bin_of(dt, T)
    src = _source_index(events, W)
    f = voxel.reshape(T, 2, -1)
    hit = b >= 0
    bb = b.clamp(min=0)
    out = events.clone()
    out[:, 0] = torch.where(hit, events[:, 0] - dt * f[bb, 0, src], events[:, 0])
    out[:, 1] = torch.where(hit, events[:, 1] - dt * f[bb, 1, src], events[:, 1])
    out[:, 2] = dt
    return out


def warp_2dof(events: torch.Tensor, theta: torch.Tensor, direction: Direction = "first",
              normalize_t: bool = True) -> torch.Tensor:
    """x' = x + dt*theta0, y' = y + dt*theta1 (translation sign).  src/warp.py:506-520."""
    dt = normalised_dt(events[:, 2], reference_time(events[:, 2], direction), normalize_t)
    return torch.stack([events[:, 0] + dt * theta[0], events[:, 1] + dt * theta[1], dt, events[:, 3]], dim=1)


# --------------------------------------------------------------------------------------
# events -> image   (src/event_image_converter.py:209-255, 316-374, 153-158)
# --------------------------------------------------------------------------------------
def vote_geometry(xy: torch.Tensor, image_size: Tuple[int, int], pad: Tuple[int, int] = (0, 0)):
    """floor indices, fractions, 4 flat target indices and 4 in-bounds masks.

    image_size is the PADDED size.  Corner order: (r,c), (r+1,c), (r,c+1), (r+1,c+1).
    src/event_image_converter.py:340-363.
    """
    h, w = image_size
    fl = torch.floor(xy[:, :2] + 1e-6)
    frac = xy[:, :2] - fl
    fl = fl.long()
    col = fl[:, 1] + pad[1]
    row = fl[:, 0] + pad[0]
    idx = torch.stack([col + row * w, col + (row + 1) * w, (col + 1) + row * w, (col + 1) + (row + 1) * w])
    c0, c1 = (0 <= col) & (col < w), (0 <= col + 1) & (col + 1 < w)
    r0, r1 = (0 <= row) & (row < h), (0 <= row + 1) & (row + 1 < h)
    mask = torch.stack([c0 & r0, c0 & r1, c1 & r0, c1 & r1])
    return row, col, frac, idx, mask


def bilinear_vote(xy: torch.Tensor, image_size: Tuple[int, int], pad: Tuple[int, int] = (0, 0),
                  weight: Union[float, torch.Tensor] = 1.0) -> torch.Tensor:
    """[n,>=2] -> [h,w] (padded size).  Accumulation order is the reference's: all corner-0 terms in event
    order, then corner 1, 2, 3 (one sequential scatter_add_ over the concatenated list).
    src/event_image_converter.py:316-374."""
    h, w = image_size
    _, _, frac, idx, mask = vote_geometry(xy, image_size, pad)
    fx, fy = frac[:, 0], frac[:, 1]
    vals = torch.stack([(1 - fx) * (1 - fy) * weight, fx * (1 - fy) * weight,
                        (1 - fx) * fy * weight, fx * fy * weight])
    idx = (idx * mask).reshape(-1)
    vals = (vals * mask).reshape(-1)
    image = xy.new_zeros(h * w)
    image.scatter_add_(0, idx, vals)
    return image.reshape(h, w)


def count_vote(xy: torch.Tensor, image_size: Tuple[int, int], pad: Tuple[int, int] = (0, 0)) -> torch.Tensor:
    """Unweighted 4-corner count.  src/event_image_converter.py:209-255."""
    h, w = image_size
    _, _, _, idx, mask = vote_geometry(xy, image_size, pad)
    image = xy.new_zeros(h * w)
    image.scatter_add_(0, (idx * mask).reshape(-1), mask.reshape(-1).to(xy.dtype))
    return image.reshape(h, w)


def gaussian_kernel3(sigma: float, dtype=torch.float32) -> torch.Tensor:
    """3 taps exp(-x^2/2s^2)/sum at x in {-1,0,1}.  torchvision 0.26 _get_gaussian_kernel1d,
    called from src/event_image_converter.py:158 with kernel_size=3."""
    x = torch.linspace(-1.0, 1.0, steps=3, dtype=dtype)
    pdf = torch.exp(-0.5 * (x / sigma) ** 2)
    return pdf / pdf.sum()


def gaussian_blur3(image: torch.Tensor, sigma: float) -> torch.Tensor:
    """3x3 Gaussian, reflect padding, one 2-D correlation with the outer-product kernel.
    (torchvision gaussian_blur(kernel_size=3, sigma) as used at src/event_image_converter.py:153-158.)"""
    k1 = gaussian_kernel3(sigma, image.dtype)
    k2 = torch.mm(k1[:, None], k1[None, :])
    img = torch.nn.functional.pad(image[None, None], (1, 1, 1, 1), mode="reflect")
    return torch.nn.functional.conv2d(img, k2[None, None])[0, 0]


def create_iwe(xy: torch.Tensor, image_size: Tuple[int, int], pad: Tuple[int, int] = (0, 0),
               method: str = "bilinear_vote", sigma: float = 0.0,
               weight: Union[float, torch.Tensor] = 1.0) -> torch.Tensor:
    """src/event_image_converter.py:126-159.  image_size here is the UNPADDED (H, W)."""
    full = (image_size[0] + 2 * pad[0], image_size[1] + 2 * pad[1])
    if method == "bilinear_vote":
        img = bilinear_vote(xy, full, pad, weight)
    elif method == "count":
        img = count_vote(xy, full, pad)
    else:
        raise NotImplementedError(method)
    if sigma > 0:
        img = gaussian_blur3(img, sigma)
    return img


# --------------------------------------------------------------------------------------
# costs   (src/costs/*.py, src/utils/stat_utils.py:51-83)
# --------------------------------------------------------------------------------------
_SOBEL_X = [[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]]   # derivative along rows (height)
_SOBEL_Y = [[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]   # derivative along columns


def image_variance(iwe: torch.Tensor, omit_boundary: bool = True) -> torch.Tensor:
    """Unbiased variance of the (cropped) image; 'natural' sign.  src/costs/image_variance.py:37-58."""
    if omit_boundary:
        iwe = iwe[1:-1, 1:-1]
    return torch.var(iwe)


def sobel_pair(iwe: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Zero-padded 3x3 cross-correlations /8.  src/utils/stat_utils.py:51-83, gradient_magnitude.py:67."""
    kx = torch.tensor(_SOBEL_X, dtype=iwe.dtype)[None, None]
    ky = torch.tensor(_SOBEL_Y, dtype=iwe.dtype)[None, None]
    x = iwe[None, None]
    gx = torch.nn.functional.conv2d(x, kx, padding=1)[0, 0] / 8.0
    gy = torch.nn.functional.conv2d(x, ky, padding=1)[0, 0] / 8.0
    return gx, gy


def gradient_magnitude(iwe: torch.Tensor, omit_boundary: bool = True) -> torch.Tensor:
    """mean(gx^2 + gy^2) over the crop; 'natural' sign.  src/costs/gradient_magnitude.py:60-76."""
    gx, gy = sobel_pair(iwe)
    if omit_boundary:
        gx, gy = gx[1:-1, 1:-1], gy[1:-1, 1:-1]
    return torch.mean(gx * gx + gy * gy)


def normalized_image_variance(iwe, orig_iwe, omit_boundary=True) -> torch.Tensor:
    """var(orig, NOT cropped) / var(iwe cropped), the 'minimize' form.
    src/costs/normalized_image_variance.py:38-64."""
    return image_variance(orig_iwe, False) / image_variance(iwe, omit_boundary)


def normalized_gradient_magnitude(iwe, orig_iwe, omit_boundary=True) -> torch.Tensor:
    """gm(orig)/gm(iwe), both cropped.  src/costs/normalized_gradient_magnitude.py:63-79."""
    return gradient_magnitude(orig_iwe, omit_boundary) / gradient_magnitude(iwe, omit_boundary)


def multi_focal(normalised_fn, forward_iwe, backward_iwe, middle_iwe, orig_iwe, omit_boundary=True):
    """N(fwd) + N(bwd) + 2 N(mid).  src/costs/multi_focal_normalized_*.py:73-101 / 64-91."""
    loss = normalised_fn(forward_iwe, orig_iwe, omit_boundary) + normalised_fn(backward_iwe, orig_iwe, omit_boundary)
    if middle_iwe is not None:
        loss = loss + normalised_fn(middle_iwe, orig_iwe, omit_boundary) * 2
    return loss


COSTS = ("image_variance", "gradient_magnitude", "normalized_image_variance",
         "normalized_gradient_magnitude", "multi_focal_normalized_image_variance",
         "multi_focal_normalized_gradient_magnitude")


def cost_value(name: str, iwes: dict, omit_boundary: bool = True) -> torch.Tensor:
    """'minimize'-direction value of a named cost from a dict of images
    (keys as in src/solver/patch_contrast_base.py:290-350)."""
    if name == "image_variance":
        return -image_variance(iwes["iwe"], omit_boundary)
    if name == "gradient_magnitude":
        return -gradient_magnitude(iwes["iwe"], omit_boundary)
    if name == "normalized_image_variance":
        return normalized_image_variance(iwes["iwe"], iwes["orig_iwe"], omit_boundary)
    if name == "normalized_gradient_magnitude":
        return normalized_gradient_magnitude(iwes["iwe"], iwes["orig_iwe"], omit_boundary)
    if name == "multi_focal_normalized_image_variance":
        return multi_focal(normalized_image_variance, iwes["forward_iwe"], iwes["backward_iwe"],
                           iwes.get("middle_iwe"), iwes["orig_iwe"], omit_boundary)
    if name == "multi_focal_normalized_gradient_magnitude":
        return multi_focal(normalized_gradient_magnitude, iwes["forward_iwe"], iwes["backward_iwe"],
                           iwes.get("middle_iwe"), iwes["orig_iwe"], omit_boundary)
    raise KeyError(name)


# --------------------------------------------------------------------------------------
# the composed objective, the way the solver seam composes it
# (src/solver/patch_contrast_base.py:289-352)
# --------------------------------------------------------------------------------------
_REFS_OF_COST = {
    "image_variance": (("iwe", "first"),),
    "gradient_magnitude": (("iwe", "first"),),
    "normalized_image_variance": (("iwe", "first"),),
    "normalized_gradient_magnitude": (("iwe", "first"),),
    "multi_focal_normalized_image_variance": (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")),
    "multi_focal_normalized_gradient_magnitude": (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")),
}


def objective(events: torch.Tensor, motion: torch.Tensor, image_size: Tuple[int, int], *,
              motion_model: str = "dense-flow", cost: str = "image_variance", sigma: float = 0.0,
              omit_boundary: bool = True, pad: Tuple[int, int] = (0, 0), method: str = "bilinear_vote",
              return_images: bool = False):
    """One CM evaluation: warp(s) -> IWE(s) -> cost, differentiable by torch autograd w.r.t. `motion`."""
    warp_fn = {"dense-flow": warp_dense, "dense-flow-voxel": warp_voxel,
               "2d-translation": warp_2dof, "rigid-optical-flow": warp_2dof}[motion_model]
    images = {}
    if "normalized" in cost:
        images["orig_iwe"] = create_iwe(events.detach(), image_size, pad, method, sigma)
    for key, direction in _REFS_OF_COST[cost]:
        warped = warp_fn(events, motion, direction)
        images[key] = create_iwe(warped, image_size, pad, method, sigma)
    value = cost_value(cost, images, omit_boundary)
    if return_images:
        return value, images
    return value


def objective_value_and_grad(events, motion, image_size, **kw):
    """(cost, dcost/dmotion) through torch autograd -- exactly what
    src/solver/scipy_autograd/torch_wrapper.py:30-49 asks of the reference."""
    motion = motion.detach().clone().requires_grad_(True)
    value = objective(events, motion, image_size, **kw)
    (grad,) = torch.autograd.grad(value, motion)
    return value.detach(), grad


# --------------------------------------------------------------------------------------
# closed forms (no autograd): dL/dIWE images and the per-event chain  
# --------------------------------------------------------------------------------------
def dvariance_dimage(iwe: torch.Tensor, omit_boundary: bool = True) -> torch.Tensor:
    """d var(crop) / d iwe  = 2/(M-1) (I - mean) inside the crop, 0 on the border."""
    g = torch.zeros_like(iwe)
    crop = iwe[1:-1, 1:-1] if omit_boundary else iwe
    m = crop.numel()
    d = 2.0 / (m - 1) * (crop - crop.mean())
    if omit_boundary:
        g[1:-1, 1:-1] = d
    else:
        g = d
    return g


def dgradmag_dimage(iwe: torch.Tensor, omit_boundary: bool = True) -> torch.Tensor:
    """Adjoint of the zero-padded Sobel pair applied to (2/M) * masked (gx, gy) / 8."""
    gx, gy = sobel_pair(iwe)
    mask = torch.zeros_like(iwe)
    if omit_boundary:
        mask[1:-1, 1:-1] = 1.0
    else:
        mask[:] = 1.0
    m = mask.sum()
    ax = (2.0 / m) * gx * mask / 8.0
    ay = (2.0 / m) * gy * mask / 8.0
    kx = torch.tensor(_SOBEL_X, dtype=iwe.dtype).flip(0, 1)[None, None]
    ky = torch.tensor(_SOBEL_Y, dtype=iwe.dtype).flip(0, 1)[None, None]
    return (torch.nn.functional.conv2d(ax[None, None], kx, padding=1)
            + torch.nn.functional.conv2d(ay[None, None], ky, padding=1))[0, 0]


def blur3_adjoint(g: torch.Tensor, sigma: float) -> torch.Tensor:
    """Transpose of gaussian_blur3 (reflect padding folds the border taps back inside)."""
    h, w = g.shape
    k1 = gaussian_kernel3(sigma, g.dtype)
    k2 = torch.mm(k1[:, None], k1[None, :])
    full = torch.nn.functional.conv_transpose2d(g[None, None], k2[None, None])[0, 0]  # (h+2, w+2)
    out = full[1:-1, 1:-1].clone()
    out[1, :] += full[0, 1:-1]
    out[h - 2, :] += full[h + 1, 1:-1]
    out[:, 1] += full[1:-1, 0]
    out[:, w - 2] += full[1:-1, w + 1]
    out[1, 1] += full[0, 0]
    out[1, w - 2] += full[0, w + 1]
    out[h - 2, 1] += full[h + 1, 0]
    out[h - 2, w - 2] += full[h + 1, w + 1]
    return out


def event_gradient(warped_xy: torch.Tensor, G: torch.Tensor, pad: Tuple[int, int] = (0, 0)):
    """(dL/dx', dL/dy') per event given G = dL/dIWE (padded size)."""
    h, w = G.shape
    _, _, frac, idx, mask = vote_geometry(warped_xy, (h, w), pad)
    fx, fy = frac[:, 0], frac[:, 1]
    g = G.reshape(-1)[idx * mask] * mask
    g00, g10, g01, g11 = g[0], g[1], g[2], g[3]
    gx = (1 - fy) * (g10 - g00) + fy * (g11 - g01)
    gy = (1 - fx) * (g01 - g00) + fx * (g11 - g10)
    return gx, gy


def flow_gradient_dense(events: torch.Tensor, dt: torch.Tensor, gx: torch.Tensor, gy: torch.Tensor,
                        image_size: Tuple[int, int]) -> torch.Tensor:
    """dL/dflow [2,H,W]: x' = x - dt f  =>  scatter -dt * dL/dx' at the source pixel."""
    H, W = image_size
    src = _source_index(events, W)
    out = events.new_zeros(2, H * W)
    out[0].scatter_add_(0, src, -dt * gx)
    out[1].scatter_add_(0, src, -dt * gy)
    return out.reshape(2, H, W)


# --------------------------------------------------------------------------------------
# tile (patch-grid) flow -> dense flow   (src/solver/patch_contrast_base.py:462-506)
# --------------------------------------------------------------------------------------
def tile_flow_geometry(image_shape, patch_size, sliding_window, patch_shift, grid):
    """(pad_h, pad_w, full_h, full_w, h1, w1): replicate padding of the patch grid, size after the integer-factor
    resize, and the offsets of the central crop.  src/solver/patch_contrast_base.py:470-479, 494-505."""
    pad_h = int(patch_size[0] / 2 // sliding_window[0]) + patch_shift[0] // sliding_window[0] + 1
    pad_w = int(patch_size[1] / 2 // sliding_window[1]) + patch_shift[1] // sliding_window[1] + 1
    full_h = (grid[0] + 2 * pad_h) * sliding_window[0]
    full_w = (grid[1] + 2 * pad_w) * sliding_window[1]
    h1 = full_h // 2 - image_shape[0] // 2
    w1 = full_w // 2 - image_shape[1] // 2
    return pad_h, pad_w, full_h, full_w, h1, w1


def upsample_tile_flow(motion: torch.Tensor, image_shape, patch_size, sliding_window, patch_shift) -> torch.Tensor:
    """[2,hp,wp] patch motion -> [2,H,W] dense flow: NEGATE, replicate-pad, bilinear resize (align_corners=False) by
    the sliding window, central crop.  Differentiable by torch autograd.  src/solver/patch_contrast_base.py:462-506
    (torchvision `resize` on a tensor = F.interpolate(mode="bilinear", align_corners=False); its antialias flag has no
    effect when up-sampling)."""
    grid = tuple(motion.shape[-2:])
    pad_h, pad_w, full_h, full_w, h1, w1 = tile_flow_geometry(image_shape, patch_size, sliding_window, patch_shift, grid)
    padded = torch.nn.functional.pad(-motion[None], (pad_w, pad_w, pad_h, pad_h), mode="replicate")
    dense = torch.nn.functional.interpolate(padded, size=(full_h, full_w), mode="bilinear", align_corners=False)[0]
    return dense[..., h1:h1 + image_shape[0], w1:w1 + image_shape[1]]


# --------------------------------------------------------------------------------------
# time-aware flow voxel   (src/utils/flow_utils.py:99-161, 439-493, 567-639)
# --------------------------------------------------------------------------------------
def _shift_rows(a: torch.Tensor, k: int) -> torch.Tensor:
    """a[i + k] with the border row replicated (k = +1 / -1)."""
    H = a.shape[-2]
    idx = torch.clamp(torch.arange(H) + k, 0, H - 1)
    return a.index_select(-2, idx)


def _shift_cols(a: torch.Tensor, k: int) -> torch.Tensor:
    W = a.shape[-1]
    idx = torch.clamp(torch.arange(W) + k, 0, W - 1)
    return a.index_select(-1, idx)


def _one_sided(a: torch.Tensor):
    """(row-backward, row-forward, col-backward, col-forward) differences of [H,W]; a replicated neighbour makes the
    difference across the image border exactly zero (the reference pads `torch.diff` with zeros)."""
    return (a - _shift_rows(a, -1), _shift_rows(a, 1) - a, a - _shift_cols(a, -1), _shift_cols(a, 1) - a)


def flow_voxel_step(flow: torch.Tensor, dt: float, scheme: str) -> torch.Tensor:
    """One explicit time step of a [2,H,W] flow (channel 0 = row component).  Negative dt = backward in time: the
    reference flips the sign of the flow, steps by |dt| and flips back (flow_utils.py:459-462, 587-590).
    upwind: flow_utils.py:464-493;  burgers: flow_utils.py:592-639.  The order of the additions follows the reference's
    expression, so fp32 results are bit-identical to it."""
    if dt == 0:
        return flow
    sgn = 1.0 if dt > 0 else -1.0
    h = abs(dt)
    f = flow * sgn
    u, v = f[0], f[1]
    zero = torch.zeros_like(u)
    up, um = torch.maximum(u, zero), torch.minimum(u, zero)
    vp, vm = torch.maximum(v, zero), torch.minimum(v, zero)
    if scheme == "upwind":
        out = []
        for c in (u, v):
            rb, rf, cb, cf = _one_sided(c)
            out.append(c - h * (((up * rb + um * rf) + vp * cb) + vm * cf))
        return torch.stack(out) * sgn
    if scheme == "burgers":
        u_b, u_f = _shift_rows(u, -1), _shift_rows(u, 1)
        v_b, v_f = _shift_cols(v, -1), _shift_cols(v, 1)
        bu = ((u * u) * torch.sign(u) + torch.maximum(torch.sign(u_b), zero) * ((-u_b) * u_b)
              - torch.minimum(torch.sign(u_f), zero) * (u_f * u_f)) / 2.0
        bv = ((v * v) * torch.sign(v) + torch.maximum(torch.sign(v_b), zero) * ((-v_b) * v_b)
              - torch.minimum(torch.sign(v_f), zero) * (v_f * v_f)) / 2.0
        _, _, u_cb, u_cf = _one_sided(u)
        v_rb, v_rf, _, _ = _one_sided(v)
        nu = u - h * ((vp * u_cb + vm * u_cf) + bu)
        nv = v - h * ((up * v_rb + um * v_rf) + bv)
        return torch.stack([nu, nv]) * sgn
    raise NotImplementedError(f"scheme {scheme!r}")


def flow_voxel(dense: torch.Tensor, time_bin: int, scheme: str = "upwind", t0_location: str = "middle") -> torch.Tensor:
    """[2,H,W] flow at t0 -> [time_bin,2,H,W].  Differentiable by torch autograd.  flow_utils.py:99-161: level t0
    (0 or time_bin // 2) holds the input; levels above are stepped forward by dt = 1/time_bin, levels below backward.
    The reference's Burgers backward loop also runs for i = 0 and so writes level -1 (= time_bin - 1) one more backward
    step away from level 0 (flow_utils.py:140-141); the forward loop then overwrites that level -- unless t0 is the
    last level (time_bin 1, or time_bin 2 with t0 in the middle), where the write sticks."""
    if t0_location not in ("first", "middle"):
        raise NotImplementedError(f"t0_location {t0_location!r} not supported")
    if scheme not in ("upwind", "burgers"):
        raise NotImplementedError(f"scheme {scheme!r}")
    T = int(time_bin)
    h = 1.0 / T
    t0 = 0 if t0_location == "first" else T // 2
    levels = [None] * T
    levels[t0] = dense
    for i in range(t0, 0, -1):
        levels[i - 1] = flow_voxel_step(levels[i], -h, scheme)
    if scheme == "burgers":
        levels[T - 1] = flow_voxel_step(levels[0], -h, scheme)
    for i in range(t0, T - 1):
        levels[i + 1] = flow_voxel_step(levels[i], h, scheme)
    return torch.stack(levels)
