"""Contrast cost plugins on the CUDA statistics kernel (reference: src/costs/image_variance.py,
gradient_magnitude.py, normalized_*.py, multi_focal_normalized_*.py).  The scalar compositions (ratios, sums) are
0-dim torch ops so that torch autograd chains them exactly like the reference; the image-sized work
(variance / Sobel magnitude and their image gradients) is one `cmax_image_stats` call per image."""
from __future__ import annotations

import logging
from typing import Optional

import torch

from .. import ops
from .base import CostBase, require_cuda_image

logger = logging.getLogger(__name__)


def _stat(iwe: torch.Tensor, stat: str, omit_boundary: bool) -> torch.Tensor:
    """Statistic of one image [H,W] or of a batch [b,H,W].  The reference reduces over the WHOLE batch: `torch.var(iwe)` after
    the crop (src/costs/image_variance.py:52-58) is the pooled unbiased variance, `torch.mean(gx^2 + gy^2)`
    (src/costs/gradient_magnitude.py:67-76) the mean of the per-image means."""
    omit = bool(omit_boundary)
    if iwe.dim() == 2:
        return ops.ImageStatFunction.apply(iwe, stat, omit)
    per_image = torch.stack([ops.ImageStatFunction.apply(i, stat, omit) for i in iwe])
    if stat != "variance" or iwe.shape[0] == 1:
        return per_image.mean()
    # pooled variance from the per-image (unbiased variance, mean, M): within-image + between-image sums of squares
    crop = iwe[..., 1:-1, 1:-1] if omit else iwe
    b, m = crop.shape[0], crop.shape[-2] * crop.shape[-1]
    means = crop.mean(dim=(-2, -1))
    within = (m - 1) * per_image.sum()
    between = m * ((means - means.mean()) ** 2).sum()
    return (within + between) / (b * m - 1)


class ImageVariance(CostBase):
    """src/costs/image_variance.py:12-58."""

    name = "image_variance"
    required_keys = ["iwe", "omit_boundary"]

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    def _loss(self, arg: dict) -> torch.Tensor:
        iwe = require_cuda_image(arg["iwe"], self.name)
        return self.calculate_torch(iwe, arg["omit_boundary"])

    def calculate_torch(self, iwe: torch.Tensor, omit_boundary: bool = False) -> torch.Tensor:
        loss = _stat(iwe, "variance", omit_boundary)
        if self.direction == "minimize":
            return -loss
        return loss


class GradientMagnitude(CostBase):
    """src/costs/gradient_magnitude.py:14-76 (Sobel/8, zero padding: src/utils/stat_utils.py:51-83)."""

    name = "gradient_magnitude"
    required_keys = ["iwe", "omit_boundary"]

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32", *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)
        self.precision = precision

    def _loss(self, arg: dict) -> torch.Tensor:
        iwe = require_cuda_image(arg["iwe"], self.name)
        return self.calculate_torch(iwe, arg["omit_boundary"])

    def calculate_torch(self, iwe: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        magnitude = _stat(iwe, "gradmag", omit_boundary)
        if self.precision == "64":
            magnitude = magnitude.double()
        if self.direction == "minimize":
            return -magnitude
        return magnitude


class NormalizedImageVariance(CostBase):
    """src/costs/normalized_image_variance.py:12-64.  NB: the warped IWE is cropped, the original is not (:38-41)."""

    name = "normalized_image_variance"
    required_keys = ["iwe", "omit_boundary", "orig_iwe"]

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)

    def _loss(self, arg: dict) -> torch.Tensor:
        iwe = require_cuda_image(arg["iwe"], self.name)
        orig_iwe = require_cuda_image(arg["orig_iwe"], self.name)
        return self.calculate_torch(iwe, orig_iwe, arg["omit_boundary"])

    def calculate_torch(self, iwe: torch.Tensor, orig_iwe: torch.Tensor, omit_boundary: bool = False) -> torch.Tensor:
        loss1 = _stat(iwe, "variance", omit_boundary)
        loss2 = _stat(orig_iwe, "variance", False)
        if self.direction == "minimize":
            return loss2 / loss1
        logger.warning("The loss is specified as maximize direction")
        return loss1 / loss2


class NormalizedGradientMagnitude(CostBase):
    """src/costs/normalized_gradient_magnitude.py:14-79."""

    name = "normalized_gradient_magnitude"
    required_keys = ["iwe", "omit_boundary", "orig_iwe"]

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32", *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)
        self.gradient_magnitude = GradientMagnitude(direction=direction, store_history=store_history,
                                                    cuda_available=cuda_available, precision=precision)

    def _loss(self, arg: dict) -> torch.Tensor:
        iwe = require_cuda_image(arg["iwe"], self.name)
        orig_iwe = require_cuda_image(arg["orig_iwe"], self.name)
        return self.calculate_torch(iwe, orig_iwe, arg["omit_boundary"])

    def calculate_torch(self, iwe: torch.Tensor, orig_iwe: torch.Tensor, omit_boundary: bool) -> torch.Tensor:
        loss1 = self.gradient_magnitude.calculate_torch(iwe, omit_boundary)
        loss2 = self.gradient_magnitude.calculate_torch(orig_iwe, omit_boundary)
        if self.direction == "minimize":
            return loss2 / loss1
        logger.warning("The loss is specified as maximize direction")
        return loss1 / loss2


class _MultiFocal(CostBase):
    """N(forward) + N(backward) + 2 N(middle).  src/costs/multi_focal_normalized_*.py."""

    required_keys = ["forward_iwe", "backward_iwe", "middle_iwe", "omit_boundary", "orig_iwe"]
    _inner_cls = None

    def __init__(self, direction="minimize", store_history: bool = False, cuda_available=False, precision="32", *args, **kwargs):
        super().__init__(direction=direction, store_history=store_history)
        self._inner = self._inner_cls(direction=direction, cuda_available=cuda_available, precision=precision)

    def _loss(self, arg: dict) -> torch.Tensor:
        orig_iwe = arg["orig_iwe"]
        forward_iwe = require_cuda_image(arg["forward_iwe"], self.name)
        middle_iwe = arg["middle_iwe"] if "middle_iwe" in arg.keys() else None
        backward_iwe = arg["backward_iwe"]
        return self.calculate_torch(orig_iwe, forward_iwe, backward_iwe, middle_iwe, arg["omit_boundary"])

    def calculate_torch(self, orig_iwe, forward_iwe, backward_iwe, middle_iwe: Optional[torch.Tensor], omit_boundary: bool):
        forward_loss = self._inner.calculate_torch(forward_iwe, orig_iwe, omit_boundary)
        backward_loss = self._inner.calculate_torch(backward_iwe, orig_iwe, omit_boundary)
        loss = forward_loss + backward_loss
        if middle_iwe is not None:
            loss = loss + self._inner.calculate_torch(middle_iwe, orig_iwe, omit_boundary) * 2
        if self.direction in ["minimize", "natural"]:
            return loss
        logger.warning("The loss is specified as maximize direction")
        return -loss


class MultiFocalNormalizedImageVariance(_MultiFocal):
    """src/costs/multi_focal_normalized_image_variance.py:13-91."""

    name = "multi_focal_normalized_image_variance"
    _inner_cls = NormalizedImageVariance

    @property
    def variance_loss(self):
        return self._inner


class MultiFocalNormalizedGradientMagnitude(_MultiFocal):
    """src/costs/multi_focal_normalized_gradient_magnitude.py:14-101."""

    name = "multi_focal_normalized_gradient_magnitude"
    _inner_cls = NormalizedGradientMagnitude

    @property
    def gradient_loss(self):
        return self._inner
