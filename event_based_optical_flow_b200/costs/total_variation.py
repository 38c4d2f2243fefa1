"""`total_variation`: smoothness regulariser on the patch-grid flow (src/costs/total_variation.py:14-126).

It acts on the [2, hp, wp] motion grid (at most 16x16 nodes), not on events or images, so it stays outside the CUDA library
(SURVEY.md section 8 row a14): the mean absolute Sobel/8 response of both flow components in both directions, boundary ring
dropped when `omit_boundary`; positive when minimising.

What matters at this size is the NUMBER of torch operators, not their work: through the reference's class one value + gradient
is ~200 aten operators (two Conv2d modules on channel slices, cat, autograd through all of it), ~0.5 ms of host time -- several
times the whole contrast term of a 30 000-event batch.  `_TotalVariationFunction` evaluates value and analytic gradient in one
forward (conv2d -> |.| -> sum, sign -> conv_transpose2d: 8 operators, no autograd graph) and its backward is one multiplication.
The function is piecewise linear, so its second derivative is zero wherever it exists: the saved gradient is a constant of the
recorded backward, which is exactly what autograd through `abs` yields (Newton-CG's Hessian-vector products see no TV term).
"""
from __future__ import annotations

import logging

import torch
import torch.nn.functional as F

from .base import CostBase

logger = logging.getLogger(__name__)

# Sobel pair of the reference (src/utils/stat_utils.py:51-52): derivative along rows, derivative along columns
_SOBEL = (((-1.0, -2.0, -1.0), (0.0, 0.0, 0.0), (1.0, 2.0, 1.0)),
          ((-1.0, 0.0, 1.0), (-2.0, 0.0, 2.0), (-1.0, 0.0, 1.0)))


_TAPS: dict = {}


def _taps(dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    """[2,1,3,3]: the Sobel pair already divided by 8 (a power of two: bit-identical to dividing the response)."""
    key = (dtype, device)
    if key not in _TAPS:
        _TAPS[key] = (torch.tensor(_SOBEL, dtype=torch.float64) / 8.0).to(dtype=dtype, device=device).unsqueeze(1)
    return _TAPS[key]


class _TotalVariationFunction(torch.autograd.Function):
    """mean |Sobel/8| of [(b,) 2, h, w] with its analytic gradient computed in the forward."""

    @staticmethod
    def forward(ctx, flow: torch.Tensor, omit_boundary: bool):
        batch = flow if flow.dim() == 4 else flow.unsqueeze(0)
        b, c, h, w = batch.shape
        taps = _taps(batch.dtype, batch.device)
        response = F.conv2d(batch.reshape(b * c, 1, h, w), taps, padding=1)  # [b*c, 2, h, w]
        crop = bool(omit_boundary) and h > 2 and w > 2
        inner = response[..., 1:-1, 1:-1] if crop else response
        n = inner.numel()
        value = inner.abs().sum() / n
        if ctx.needs_input_grad[0]:
            s = torch.sign(inner) / n
            if crop:
                s = F.pad(s, (1, 1, 1, 1))
            ctx.save_for_backward(F.conv_transpose2d(s, taps, padding=1).reshape(flow.shape), flow)
        return value

    @staticmethod
    def backward(ctx, g):
        grad, flow = ctx.saved_tensors
        if torch.is_grad_enabled():  # a recorded backward (Hessian-vector products): connected to `flow`, with a zero derivative
            return grad * g + flow * 0, None
        return grad * g, None


def total_variation_loss(flow: torch.Tensor, omit_boundary: bool, direction: str = "minimize") -> torch.Tensor:
    """`TotalVariation(direction).calculate({"flow": flow, "omit_boundary": omit_boundary})` of the reference for a tensor."""
    tv = _TotalVariationFunction.apply(flow, omit_boundary)
    return tv if direction == "minimize" else -tv


class TotalVariation(CostBase):
    name = "total_variation"
    required_keys = ["flow", "omit_boundary"]

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32", *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    def _loss(self, arg: dict) -> torch.Tensor:
        flow = arg["flow"]
        if not isinstance(flow, torch.Tensor):
            msg = f"Unsupported input type. {type(flow)}."
            logger.error(msg)
            raise NotImplementedError(msg)
        return self.calculate_torch(flow, arg["omit_boundary"])

    def calculate_torch(self, flow: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        if self.direction != "minimize":
            logger.warning("The loss is specified as maximize direction")
        return total_variation_loss(flow, omit_boundary, self.direction)

    def get_sobel_image_torch(self, flow: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        """[(b,) 2, h, w] -> [b, 4, h, w] = per flow component (d/d row, d/d col), Sobel/8 with zero padding
        (src/utils/stat_utils.py:64-83 applied with in_channels=2)."""
        batch = flow if flow.dim() == 4 else flow.unsqueeze(0)
        b, c, h, w = batch.shape
        taps = torch.tensor(_SOBEL, dtype=batch.dtype, device=batch.device).unsqueeze(1)  # [2,1,3,3]
        response = F.conv2d(batch.reshape(b * c, 1, h, w), taps, padding=1).reshape(b, 2 * c, h, w) / 8.0
        if omit_boundary and h > 2 and w > 2:
            response = response[..., 1:-1, 1:-1]
        return response
