"""`total_variation`: smoothness regulariser on the patch-grid flow (src/costs/total_variation.py:14-126).

It acts on the [2, hp, wp] motion grid (at most 16x16 nodes), not on events or images, so it is a handful of torch ops on a
few hundred floats and stays outside the CUDA library: the mean absolute Sobel/8 response of
both flow components in both directions, boundary ring dropped when `omit_boundary`; positive when minimising.
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
        tv = self.get_sobel_image_torch(flow, omit_boundary).abs().mean()
        if self.direction != "minimize":
            logger.warning("The loss is specified as maximize direction")
            tv = -tv
        return tv

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
