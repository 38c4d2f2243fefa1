"""Modular CUDA operators (one C-ABI call each) + their autograd wrappers.

These sit behind the drop-in `Warp` / `EventImageConverter` / cost classes so that the reference's unchanged solver
code (which composes warp -> IWE -> cost itself, src/solver/patch_contrast_base.py:289-352, and differentiates with
torch autograd, scipy_autograd/torch_wrapper.py:30-73) runs on the hand-written kernels.  Everything computes in fp32
on the current CUDA stream; callers convert dtypes.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
from torch.autograd.function import once_differentiable

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("the B200 contrast-maximization operators need CUDA tensors (no CPU fallback)")
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------ time
def time_params(events: torch.Tensor, directions: Sequence, n_bins: int = 0, normalize_t: bool = True,
                t_range: Optional[Tuple[float, float]] = None) -> torch.Tensor:
    """Device-resident cmax_time_params_t for these events (min/max of t taken from THIS batch, as the reference
    does on every warp call, src/warp.py:201-259; `t_range` = the GLOBAL range when the batch is one shard of a larger one)."""
    ev = _f32c(events)
    with torch.cuda.device(ev.device):
        if t_range is not None:
            mm = torch.tensor([float(t_range[0]), float(t_range[1])], dtype=torch.float32, device=ev.device)
        else:
            mm = torch.empty(2, dtype=torch.float32, device=ev.device)
            _lib.call("cmax_time_range", ev.data_ptr(), ev.shape[0], ev.shape[1], mm.data_ptr(), _stream())
        tp = torch.empty(_lib.TIME_PARAMS_BYTES, dtype=torch.uint8, device=ev.device)
        _lib.call("cmax_time_params", mm.data_ptr(), _lib.refs_array(tuple(directions)), len(directions), int(n_bins),
                  1 if normalize_t else 0, tp.data_ptr(), _stream())
    return tp


# ------------------------------------------------------------------------------------------------ sharded composition
class SumPartials(torch.autograd.Function):
    """y = sum over ranks of x (an all-reduce), every rank holding the result.  Dual of `UseReplicated`: under rank-local
    autograd of REPLICATED scalars (every rank differentiates its own copy of the same cost) the cotangent of y is replicated
    and each rank's x simply receives it -- but as an operation in the recorded backward it is a replicated value put to local
    use, whose own adjoint is a sum over ranks again.  Making the two Functions each other's backward keeps derivatives of ANY
    order correct (a Hessian-vector product needs the all-reduce of the tangent images sum_q J_q v; it appears here as the
    backward of the `UseReplicated` in `SumPartials.backward`)."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.detach().clone()
        torch.distributed.all_reduce(y, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return UseReplicated.apply(g, ctx.group), None


class UseReplicated(torch.autograd.Function):
    """Identity on a value every rank holds identically and uses in its own partial computation (the motion; dL/dIWE);
    its adjoint sums the ranks' contributions."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return SumPartials.apply(g, ctx.group), None


# ------------------------------------------------------------------------------------------------ warp
def warp_events(events: torch.Tensor, motion: torch.Tensor, motion_model: str, image_size: Tuple[int, int], tp: torch.Tensor,
                ref_index: int = 0) -> torch.Tensor:
    """fp32 [n,C] -> fp32 [n,C] = (x', y', dt, p).  Raises IndexError on a source pixel outside the image."""
    ev, m = _f32c(events), _f32c(motion)
    H, W = image_size
    out = torch.empty_like(ev)
    with torch.cuda.device(ev.device):
        status = torch.zeros(1, dtype=torch.int32, device=ev.device)
        _lib.call("cmax_warp_events", ev.data_ptr(), ev.shape[0], ev.shape[1], H, W, _lib.MOTION[motion_model], m.data_ptr(),
                  tp.data_ptr(), ref_index, out.data_ptr(), status.data_ptr(), _stream())
    if motion_model in ("dense-flow", "dense-flow-voxel") and int(status.item()) != 0:
        raise IndexError(f"an event's pixel lies outside the {H}x{W} flow field (the reference's torch.gather raises, src/warp.py:305-307)")
    return out


def warp_events_backward(events: torch.Tensor, motion_model: str, motion_shape, image_size, tp: torch.Tensor, ref_index: int,
                         grad_out: torch.Tensor) -> torch.Tensor:
    ev, g = _f32c(events), _f32c(grad_out)
    H, W = image_size
    n_bins = motion_shape[0] if motion_model == "dense-flow-voxel" else 0
    gm = torch.empty(tuple(motion_shape), dtype=torch.float32, device=ev.device)
    with torch.cuda.device(ev.device):
        _lib.call("cmax_warp_events_backward", ev.data_ptr(), ev.shape[0], ev.shape[1], H, W, _lib.MOTION[motion_model], n_bins,
                  tp.data_ptr(), ref_index, g.data_ptr(), gm.data_ptr(), _stream())
    return gm


class WarpFunction(torch.autograd.Function):
    """Differentiable w.r.t. `motion` only (events never need a gradient: scipy_autograd/torch_wrapper.py:38-40)."""

    @staticmethod
    def forward(ctx, events, motion, motion_model, image_size, tp, ref_index):
        out = warp_events(events, motion, motion_model, image_size, tp, ref_index)
        ctx.save_for_backward(events.detach())
        ctx.meta = (motion_model, tuple(motion.shape), tuple(image_size), tp, ref_index, motion.dtype)
        return out.to(events.dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (events,) = ctx.saved_tensors
        gm = _WarpBackward.apply(grad_out, events, ctx.meta)
        return None, gm, None, None, None, None


def warp_events_tangent(events: torch.Tensor, motion_model: str, image_size, tp: torch.Tensor, ref_index: int,
                        tangent_motion: torch.Tensor) -> torch.Tensor:
    """[n,2] = d(x', y') / d motion . tangent_motion (the warp is linear in the motion)."""
    ev, tm = _f32c(events), _f32c(tangent_motion)
    H, W = image_size
    out = torch.empty(ev.shape[0], 2, dtype=torch.float32, device=ev.device)
    with torch.cuda.device(ev.device):
        _lib.call("cmax_warp_events_tangent", ev.data_ptr(), ev.shape[0], ev.shape[1], H, W, _lib.MOTION[motion_model], tp.data_ptr(), ref_index,
                  tm.data_ptr(), out.data_ptr(), _stream())
    return out


class _WarpBackward(torch.autograd.Function):
    """grad_out [n,C] -> grad_motion, as a differentiable (linear) function of grad_out: what a Hessian-vector product
    differentiates (scipy_autograd/torch_wrapper.py:51-73)."""

    @staticmethod
    def forward(ctx, grad_out, events, meta):
        motion_model, mshape, image_size, tp, ref_index, mdtype = meta
        ctx.save_for_backward(events)
        ctx.meta = meta
        ctx.gshape, ctx.gdtype = grad_out.shape, grad_out.dtype
        return warp_events_backward(events, motion_model, mshape, image_size, tp, ref_index, grad_out).to(mdtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        (events,) = ctx.saved_tensors
        motion_model, mshape, image_size, tp, ref_index, mdtype = ctx.meta
        t = warp_events_tangent(events, motion_model, image_size, tp, ref_index, u)
        full = torch.zeros(ctx.gshape, dtype=ctx.gdtype, device=t.device)
        full[:, :2] = t.to(ctx.gdtype)
        return full, None, None


# ------------------------------------------------------------------------------------------------ vote
def vote(xy: torch.Tensor, padded_size: Tuple[int, int], pad: Tuple[int, int], weight: Optional[torch.Tensor],
         method: str = "bilinear_vote") -> torch.Tensor:
    """[n,>=2] -> fp32 [Hp,Wp].  src/event_image_converter.py:316-374 (bilinear_vote), :209-255 (count)."""
    if method not in _lib.VOTE:
        raise NotImplementedError(f"{method = } is not implemented")
    x = _f32c(xy)
    Hp, Wp = padded_size
    img = torch.empty(Hp, Wp, dtype=torch.float32, device=x.device)
    w = _f32c(weight) if weight is not None else None
    with torch.cuda.device(x.device):
        _lib.call("cmax_vote", x.data_ptr(), x.shape[0], x.shape[1], w.data_ptr() if w is not None else None, Hp, Wp, pad[0], pad[1],
                  _lib.VOTE[method], img.data_ptr(), _stream())
    return img


def vote_backward(xy: torch.Tensor, padded_size, pad, weight: Optional[torch.Tensor], grad_image: torch.Tensor,
                  want_weight_grad: bool = False):
    x, g = _f32c(xy), _f32c(grad_image)
    Hp, Wp = padded_size
    gxy = torch.empty(x.shape[0], 2, dtype=torch.float32, device=x.device)
    gw = torch.empty(x.shape[0], dtype=torch.float32, device=x.device) if want_weight_grad else None
    w = _f32c(weight) if weight is not None else None
    with torch.cuda.device(x.device):
        _lib.call("cmax_vote_backward", x.data_ptr(), x.shape[0], x.shape[1], w.data_ptr() if w is not None else None, Hp, Wp,
                  pad[0], pad[1], g.data_ptr(), gxy.data_ptr(), gw.data_ptr() if gw is not None else None, _stream())
    return gxy, gw


class VoteFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xy, weight, padded_size, pad, method):
        img = vote(xy, padded_size, pad, weight, method)
        ctx.save_for_backward(xy, weight.detach() if weight is not None else None)  # xy stays attached: the second order differentiates through it
        ctx.meta = (tuple(padded_size), tuple(pad), method, xy.dtype, xy.shape, weight is not None and ctx.needs_input_grad[1])
        return img.to(xy.dtype)

    @staticmethod
    def backward(ctx, grad_image):
        xy, weight = ctx.saved_tensors
        padded_size, pad, method, dtype, shape, want_w = ctx.meta
        if method != "bilinear_vote":  # the count image is piecewise constant in the coordinates
            return torch.zeros(shape, dtype=dtype, device=xy.device), None, None, None, None
        if want_w:  # weights with a gradient: first order only
            gxy, gw = vote_backward(xy, padded_size, pad, weight, grad_image, True)
            full = torch.zeros(shape, dtype=dtype, device=xy.device)
            full[:, :2] = gxy.to(dtype)
            return full, gw.to(weight.dtype), None, None, None
        return _VoteBackward.apply(xy, grad_image, weight, (padded_size, pad, dtype, shape)), None, None, None, None


def vote_backward2(xy, padded_size, pad, weight, grad_image, u):
    """Adjoint of `vote_backward` w.r.t. (grad_image, xy) for the cotangent u [n,>=2] of grad_xy."""
    x, g, uu = _f32c(xy), _f32c(grad_image), _f32c(u)
    Hp, Wp = padded_size
    g_img = torch.empty(Hp, Wp, dtype=torch.float32, device=x.device)
    g_xy = torch.empty(x.shape[0], 2, dtype=torch.float32, device=x.device)
    w = _f32c(weight) if weight is not None else None
    with torch.cuda.device(x.device):
        _lib.call("cmax_vote_backward2", x.data_ptr(), x.shape[0], x.shape[1], w.data_ptr() if w is not None else None, Hp, Wp, pad[0], pad[1],
                  g.data_ptr(), uu.data_ptr(), uu.shape[1], g_img.data_ptr(), g_xy.data_ptr(), _stream())
    return g_img, g_xy


class _VoteBackward(torch.autograd.Function):
    """(xy, grad_image) -> grad_xy, differentiable once more (bilinear in grad_image, piecewise bilinear in xy)."""

    @staticmethod
    def forward(ctx, xy, grad_image, weight, meta):
        padded_size, pad, dtype, shape = meta
        ctx.save_for_backward(xy, grad_image, weight)
        ctx.meta = meta
        ctx.gi_dtype = grad_image.dtype
        gxy, _ = vote_backward(xy, padded_size, pad, weight, grad_image, False)
        full = torch.zeros(shape, dtype=dtype, device=xy.device)
        full[:, :2] = gxy.to(dtype)
        return full

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        xy, grad_image, weight = ctx.saved_tensors
        padded_size, pad, dtype, shape = ctx.meta
        g_img, g_xy = vote_backward2(xy, padded_size, pad, weight, grad_image, u)
        full = torch.zeros(shape, dtype=dtype, device=xy.device)
        full[:, :2] = g_xy.to(dtype)
        return full, g_img.to(ctx.gi_dtype), None, None


# ------------------------------------------------------------------------------------------------ blur
def blur3(images: torch.Tensor, sigma: float, transpose: bool = False) -> torch.Tensor:
    """[k,Hp,Wp] fp32 -> 3x3 Gaussian, reflect padding (torchvision gaussian_blur(kernel_size=3) at
    src/event_image_converter.py:153-158); transpose=True applies the adjoint."""
    x = _f32c(images)
    k, Hp, Wp = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.call("cmax_blur3", x.data_ptr(), out.data_ptr(), k, Hp, Wp, float(sigma), 1 if transpose else 0, _stream())
    return out


class BlurFunction(torch.autograd.Function):
    """Linear, so its backward is the same Function with the transpose flag flipped (differentiable to any order)."""

    @staticmethod
    def forward(ctx, image, sigma, transpose=False):
        ctx.sigma, ctx.transpose = sigma, transpose
        ctx.dtype = image.dtype
        return blur3(image[None], sigma, transpose=transpose)[0].to(image.dtype)

    @staticmethod
    def backward(ctx, g):
        return BlurFunction.apply(g, ctx.sigma, not ctx.transpose).to(ctx.dtype), None, None


# ------------------------------------------------------------------------------------------------ statistics
def image_stats(images: torch.Tensor, stat: str, omit_boundary: bool, want_grad: bool = True):
    """[k,Hp,Wp] -> (stats float64 [k,4] = {value, mean, M, 0}, d value / d image fp32 [k,Hp,Wp] or None).
    value = unbiased variance of the crop (src/costs/image_variance.py:37-58) or mean (Sobel/8)^2 magnitude
    (src/costs/gradient_magnitude.py:60-76)."""
    x = _f32c(images)
    k, Hp, Wp = x.shape
    lib = _lib.load()
    stats = torch.empty(k, 4, dtype=torch.float64, device=x.device)
    grad = torch.empty_like(x) if want_grad else None
    with torch.cuda.device(x.device):
        ws = torch.empty(lib.cmax_stats_workspace_bytes(k, Hp, Wp) + 256, dtype=torch.uint8, device=x.device)
        ws_ptr = (ws.data_ptr() + 255) // 256 * 256
        _lib.call("cmax_image_stats", x.data_ptr(), k, Hp, Wp, _lib.STAT[stat], 1 if omit_boundary else 0, stats.data_ptr(),
                  grad.data_ptr() if grad is not None else None, ws_ptr, _stream())
    return stats, grad


class ImageStatFunction(torch.autograd.Function):
    """0-dim statistic of one image, differentiable through the kernel's closed-form image gradient."""

    @staticmethod
    def forward(ctx, image, stat, omit_boundary):
        stats, _ = image_stats(image[None], stat, omit_boundary, want_grad=False)
        ctx.save_for_backward(image)
        ctx.meta = (stat, omit_boundary)
        return stats[0, 0].to(image.dtype)

    @staticmethod
    def backward(ctx, g):
        (image,) = ctx.saved_tensors
        return g * _StatGrad.apply(image, *ctx.meta), None, None


class _StatGrad(torch.autograd.Function):
    """image -> d stat / d image.  Both statistics are quadratic forms in the image (unbiased variance, mean squared Sobel
    magnitude), so this map is LINEAR and symmetric: its adjoint is itself."""

    @staticmethod
    def forward(ctx, image, stat, omit_boundary):
        ctx.meta = (stat, omit_boundary)
        return image_stats(image[None], stat, omit_boundary, want_grad=True)[1][0].to(image.dtype)

    @staticmethod
    def backward(ctx, u):
        return _StatGrad.apply(u, *ctx.meta), None, None


# ------------------------------------------------------------------------------------------------ tile flow
def tile_flow_geometry(image_shape, patch_size, sliding_window, patch_shift):
    """(pad_h, pad_w) of the replicate padding, src/solver/patch_contrast_base.py:470-479."""
    pad_h = int(patch_size[0] / 2 // sliding_window[0]) + patch_shift[0] // sliding_window[0] + 1
    pad_w = int(patch_size[1] / 2 // sliding_window[1]) + patch_shift[1] // sliding_window[1] + 1
    return pad_h, pad_w


def tile_flow_upsample(motion: torch.Tensor, image_shape, pad, window) -> torch.Tensor:
    m = _f32c(motion)
    _, hp, wp = m.shape
    H, W = int(image_shape[0]), int(image_shape[1])
    dense = torch.empty(2, H, W, dtype=torch.float32, device=m.device)
    with torch.cuda.device(m.device):
        _lib.call("cmax_tile_flow_upsample", m.data_ptr(), hp, wp, int(pad[0]), int(pad[1]), int(window[0]), int(window[1]), H, W,
                  dense.data_ptr(), _stream())
    return dense


def tile_flow_upsample_backward(grad_dense: torch.Tensor, grid, pad, window) -> torch.Tensor:
    g = _f32c(grad_dense)
    _, H, W = g.shape
    hp, wp = int(grid[0]), int(grid[1])
    gm = torch.empty(2, hp, wp, dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        _lib.call("cmax_tile_flow_upsample_backward", g.data_ptr(), hp, wp, int(pad[0]), int(pad[1]), int(window[0]), int(window[1]), H, W,
                  gm.data_ptr(), _stream())
    return gm


class TileFlowFunction(torch.autograd.Function):
    """[2,hp,wp] patch motion -> [2,H,W] dense flow (negated, replicate-padded, bilinear, cropped), differentiable."""

    @staticmethod
    def forward(ctx, motion, image_shape, pad, window):
        ctx.meta = (tuple(motion.shape[-2:]), tuple(pad), tuple(window), motion.dtype)
        return tile_flow_upsample(motion, image_shape, pad, window).to(motion.dtype)

    @staticmethod
    def backward(ctx, g):
        return _TileFlowBackward.apply(g, ctx.meta, tuple(g.shape[-2:])), None, None, None


class _TileFlowBackward(torch.autograd.Function):
    """The (linear) adjoint of the upsample; its own adjoint is the upsample."""

    @staticmethod
    def forward(ctx, g, meta, image_shape):
        grid, pad, window, dtype = meta
        ctx.meta, ctx.image_shape, ctx.gdtype = meta, image_shape, g.dtype
        return tile_flow_upsample_backward(g, grid, pad, window).to(dtype)

    @staticmethod
    @once_differentiable
    def backward(ctx, u):
        grid, pad, window, dtype = ctx.meta
        return tile_flow_upsample(u, ctx.image_shape, pad, window).to(ctx.gdtype), None, None


# ------------------------------------------------------------------------------------------------ time-aware flow voxel
def _voxel_args(scheme: str, t0_location: str):
    if t0_location not in ("first", "middle"):
        raise NotImplementedError(f"{t0_location =} not supported")  # src/utils/flow_utils.py:119-122
    if scheme not in _lib.SCHEME:
        raise NotImplementedError(f"flow-voxel scheme {scheme!r} has no CUDA form (available: {sorted(_lib.SCHEME)})")
    return _lib.SCHEME[scheme], 1 if t0_location == "middle" else 0


def flow_voxel(dense: torch.Tensor, time_bin: int, scheme: str = "upwind", t0_location: str = "middle") -> torch.Tensor:
    """[2,H,W] flow at t0 -> [time_bin,2,H,W] (construct_dense_flow_voxel_torch, src/utils/flow_utils.py:99-161)."""
    sc, mid = _voxel_args(scheme, t0_location)
    d = _f32c(dense)
    if d.dim() != 3 or d.shape[0] != 2:
        raise ValueError(f"dense flow must be [2,H,W], got {tuple(d.shape)}")
    _, H, W = d.shape
    vox = torch.empty(int(time_bin), 2, H, W, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.call("cmax_flow_voxel", d.data_ptr(), H, W, int(time_bin), sc, mid, vox.data_ptr(), _stream())
    return vox


def flow_voxel_backward(dense: torch.Tensor, voxel: torch.Tensor, grad_voxel: torch.Tensor, scheme: str = "upwind",
                        t0_location: str = "middle") -> torch.Tensor:
    """Adjoint of `flow_voxel`: grad_voxel [T,2,H,W] -> grad_dense [2,H,W]; `dense`, `voxel` = the forward's input / output."""
    sc, mid = _voxel_args(scheme, t0_location)
    d, v, g = _f32c(dense), _f32c(voxel), _f32c(grad_voxel)
    T, _, H, W = v.shape
    if g.shape != v.shape:
        raise ValueError(f"grad_voxel must have the voxel's shape {tuple(v.shape)}, got {tuple(g.shape)}")
    out = torch.empty(2, H, W, dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        ws = torch.empty(_lib.load().cmax_flow_voxel_workspace_bytes(H, W), dtype=torch.uint8, device=d.device)
        _lib.call("cmax_flow_voxel_backward", d.data_ptr(), v.data_ptr(), g.data_ptr(), H, W, T, sc, mid, out.data_ptr(), ws.data_ptr(),
                  _stream())
    return out


# ---- second order of the voxel propagation.  First order runs on the hand-written kernels above (forward bit-identical to the
# reference, exact adjoint).  A Hessian-vector product (Newton-CG / trust-* on the time-aware configurations,
# configs/mvsec_indoor_burgers.yaml:48 through scipy_autograd/torch_wrapper.py:51-73) differentiates the ADJOINT once more; the
# propagation is piecewise quadratic, and rather than a third and fourth kernel family that derivative is taken by torch autograd
# over the same explicit steps written with torch CUDA ops (T - 1 steps of a dozen elementwise kernels on [2,H,W] images -- a
# few hundred microseconds, paid only inside hessp calls).  Formulas: src/utils/flow_utils.py:439-493 (upwind), :567-639 (Burgers).
def _shift(a: torch.Tensor, k: int, dim: int) -> torch.Tensor:
    """a[i + k] along `dim` with the border replicated."""
    n = a.shape[dim]
    idx = torch.clamp(torch.arange(n, device=a.device) + k, 0, n - 1)
    return a.index_select(dim, idx)


def _voxel_step_torch(flow: torch.Tensor, dt: float, scheme: str) -> torch.Tensor:
    if dt == 0:
        return flow
    sgn = 1.0 if dt > 0 else -1.0
    h = abs(dt)
    f = flow * sgn
    u, v = f[0], f[1]
    zero = torch.zeros_like(u)
    up, um, vp, vm = torch.maximum(u, zero), torch.minimum(u, zero), torch.maximum(v, zero), torch.minimum(v, zero)

    def one_sided(c):  # (row-backward, row-forward, col-backward, col-forward); zero across the image border
        return c - _shift(c, -1, -2), _shift(c, 1, -2) - c, c - _shift(c, -1, -1), _shift(c, 1, -1) - c

    if scheme == "upwind":
        out = []
        for c in (u, v):
            rb, rf, cb, cf = one_sided(c)
            out.append(c - h * (((up * rb + um * rf) + vp * cb) + vm * cf))
        return torch.stack(out) * sgn
    u_b, u_f, v_b, v_f = _shift(u, -1, -2), _shift(u, 1, -2), _shift(v, -1, -1), _shift(v, 1, -1)
    bu = ((u * u) * torch.sign(u) + torch.maximum(torch.sign(u_b), zero) * ((-u_b) * u_b) - torch.minimum(torch.sign(u_f), zero) * (u_f * u_f)) / 2.0
    bv = ((v * v) * torch.sign(v) + torch.maximum(torch.sign(v_b), zero) * ((-v_b) * v_b) - torch.minimum(torch.sign(v_f), zero) * (v_f * v_f)) / 2.0
    _, _, u_cb, u_cf = one_sided(u)
    v_rb, v_rf, _, _ = one_sided(v)
    return torch.stack([u - h * ((vp * u_cb + vm * u_cf) + bu), v - h * ((up * v_rb + um * v_rf) + bv)]) * sgn


def flow_voxel_torch(dense: torch.Tensor, time_bin: int, scheme: str, t0_location: str) -> torch.Tensor:
    """The propagation as a twice-differentiable composition of torch ops (same levels as `flow_voxel`)."""
    _voxel_args(scheme, t0_location)
    T = int(time_bin)
    h = 1.0 / T
    t0 = 0 if t0_location == "first" else T // 2
    levels = [None] * T
    levels[t0] = dense
    for i in range(t0, 0, -1):
        levels[i - 1] = _voxel_step_torch(levels[i], -h, scheme)
    if scheme == "burgers":  # the reference's backward loop also writes level -1 (flow_utils.py:140-141)
        levels[T - 1] = _voxel_step_torch(levels[0], -h, scheme)
    for i in range(t0, T - 1):
        levels[i + 1] = _voxel_step_torch(levels[i], h, scheme)
    return torch.stack(levels)


class FlowVoxelFunction(torch.autograd.Function):
    """Differentiable `flow_voxel` in the caller's dtype."""

    @staticmethod
    def forward(ctx, dense, time_bin, scheme, t0_location):
        vox = flow_voxel(dense, time_bin, scheme, t0_location)
        ctx.save_for_backward(dense, vox)  # (the input itself: a recorded backward differentiates through it)
        ctx.meta = (scheme, t0_location, dense.dtype)
        return vox.to(dense.dtype)

    @staticmethod
    def backward(ctx, g):
        dense, vox = ctx.saved_tensors
        scheme, t0_location, dtype = ctx.meta
        if torch.is_grad_enabled() and dense.requires_grad:
            # the backward itself is being recorded (create_graph=True: a Hessian-vector product): J^T g as a differentiable
            # function of (dense, g), through the torch-op restatement of the propagation
            with torch.enable_grad():
                v = flow_voxel_torch(dense, vox.shape[0], scheme, t0_location)
                (gd,) = torch.autograd.grad(v, dense, g.to(v.dtype), create_graph=True)
            return gd, None, None, None
        return flow_voxel_backward(dense, vox, g, scheme, t0_location).to(dtype), None, None, None
