"""`EventImageConverter` on the sm_100a kernels: events -> image (IWE, count image, event mask).

Interface of the reference's class (src/event_image_converter.py:14-374), tensors only: `image_size` / `outer_padding`
attributes (image_size includes the padding on both sides), `create_iwe(events, method, sigma)`, `create_eventmask`,
`create_image_from_events_tensor`, `bilinear_vote_tensor(events, weight)`, `count_event_tensor`, `update_property`.
Batched events `[b, n, 4]` give `[b, H, W]`; anything but a tensor raises RuntimeError, an unknown method
NotImplementedError -- the reference's behaviour.  The arithmetic is one `cmax_vote` launch per image (+ `cmax_blur3`).
"""
from __future__ import annotations

import logging
from typing import Optional, Tuple, Union

import torch

from . import ops

logger = logging.getLogger(__name__)

_METHODS = ("bilinear_vote", "count")


def _pair(padding) -> Tuple[int, int]:
    return (int(padding), int(padding)) if isinstance(padding, (int, float)) else (int(padding[0]), int(padding[1]))


class EventImageConverter:
    def __init__(self, image_size: tuple, outer_padding: Union[int, Tuple[int, int]] = 0):
        self.outer_padding = _pair(outer_padding)
        self.image_size = (int(image_size[0]) + 2 * self.outer_padding[0], int(image_size[1]) + 2 * self.outer_padding[1])

    def update_property(self, image_size: Optional[tuple] = None, outer_padding=None):
        """Same quirk as src/event_image_converter.py:30-43: the padding is added ONCE (not on both sides) here."""
        size = tuple(image_size) if image_size is not None else self.image_size
        if outer_padding is not None:
            self.outer_padding = _pair(outer_padding)
        self.image_size = (size[0] + self.outer_padding[0], size[1] + self.outer_padding[1])

    # ---- public entry points
    def create_iwe(self, events: torch.Tensor, method: str = "bilinear_vote", sigma: int = 1) -> torch.Tensor:
        self._need_tensor(events)
        return self.create_image_from_events_tensor(events, method, sigma=sigma)

    def create_eventmask(self, events: torch.Tensor) -> torch.Tensor:
        """[(b,) 1, H, W] bool: pixels hit by at least one event (src/event_image_converter.py:69-82)."""
        self._need_tensor(events, silent=True)
        hit = self.create_image_from_events_tensor(events, sigma=0) != 0
        return hit.unsqueeze(-3)

    def create_image_from_events_tensor(self, events: torch.Tensor, method: str = "bilinear_vote", weight=1.0, sigma: int = 0):
        if method not in _METHODS:
            msg = f"{method = } is not implemented"
            logger.error(msg)
            raise NotImplementedError(msg)
        image = self._vote(events, weight if method == "bilinear_vote" else 1.0, method)
        if sigma > 0:  # 3x3 Gaussian, reflect padding: what torchvision's gaussian_blur(kernel_size=3) does at :153-158
            planes = image if image.dim() == 3 else image[None]
            planes = torch.stack([ops.BlurFunction.apply(p, float(sigma)) for p in planes])
            image = planes if image.dim() == 3 else planes[0]
        return image.squeeze()

    def bilinear_vote_tensor(self, events: torch.Tensor, weight=1.0) -> torch.Tensor:
        """Each event adds its 4 bilinear weights (times `weight`) to the corners that lie inside the image (:316-374)."""
        return self._vote(events, weight, "bilinear_vote")

    def count_event_tensor(self, events: torch.Tensor) -> torch.Tensor:
        """Each event adds 1 to every corner that lies inside the image (:209-255)."""
        return self._vote(events, 1.0, "count")

    # ---- helpers
    @staticmethod
    def _need_tensor(events, silent: bool = False) -> None:
        if isinstance(events, torch.Tensor):
            return
        if silent:
            raise RuntimeError
        msg = f"Non-supported type of events. {type(events)}"
        logger.error(msg)
        raise RuntimeError(msg)

    def _vote(self, events: torch.Tensor, weight, method: str) -> torch.Tensor:
        if events.dim() == 3:  # a batch: one image per entry, per-entry weights if given as [b, n]
            per_entry = weight if isinstance(weight, torch.Tensor) and weight.dim() == 2 else [weight] * events.shape[0]
            return torch.stack([self._vote(ev, w, method) for ev, w in zip(events, per_entry)])
        w = None
        if isinstance(weight, torch.Tensor):
            w = weight.to(events.device).expand(events.shape[0]) if weight.dim() == 0 else weight
        elif float(weight) != 1.0:
            w = torch.full((events.shape[0],), float(weight), dtype=events.dtype, device=events.device)
        return ops.VoteFunction.apply(events, w, tuple(self.image_size), tuple(self.outer_padding), method)
