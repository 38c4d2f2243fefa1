"""Total variation of the (tiny, <=16x16) patch-grid flow: reference src/costs/total_variation.py:14-126.
Acts on [2,hp,wp], not on events or images, so it stays in torch."""
from __future__ import annotations

import logging

import torch

from .base import CostBase

logger = logging.getLogger(__name__)

_KX = [[-1.0, -2.0, -1.0], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0]]
_KY = [[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]


class TotalVariation(CostBase):
    name = "total_variation"
    required_keys = ["flow", "omit_boundary"]

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32", *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    @CostBase.register_history
    @CostBase.catch_key_error
    def calculate(self, arg: dict) -> torch.Tensor:
        flow = arg["flow"]
        if not isinstance(flow, torch.Tensor):
            e = f"Unsupported input type. {type(flow)}."
            logger.error(e)
            raise NotImplementedError(e)
        return self.calculate_torch(flow, arg["omit_boundary"])

    def calculate_torch(self, flow: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        loss = torch.mean(torch.abs(self.get_sobel_image_torch(flow, omit_boundary)))
        if self.direction == "minimize":
            return loss
        logger.warning("The loss is specified as maximize direction")
        return -loss

    def get_sobel_image_torch(self, flow: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        """[(b,) 2, h, w] -> [b, 4, h, w]: (x-comp d/dx, x-comp d/dy, y-comp d/dx, y-comp d/dy), Sobel/8, zero pad
        (src/utils/stat_utils.py:64-83 with in_channels=2)."""
        if flow.dim() == 3:
            flow = flow[None]
        k = torch.tensor([_KX, _KY], dtype=flow.dtype, device=flow.device)[:, None]  # [2,1,3,3]
        b, c, h, w = flow.shape
        sobel = torch.nn.functional.conv2d(flow.reshape(b * c, 1, h, w), k, padding=1).reshape(b, 2 * c, h, w) / 8.0
        if omit_boundary and sobel.shape[2] > 2 and sobel.shape[3] > 2:
            sobel = sobel[..., 1:-1, 1:-1]
        return sobel
