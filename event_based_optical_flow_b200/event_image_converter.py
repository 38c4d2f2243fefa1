"""Drop-in `EventImageConverter` (reference: src/event_image_converter.py:14-374) on the sm_100a kernels."""
from __future__ import annotations

import logging
from typing import Optional, Tuple, Union

import torch

from . import ops

logger = logging.getLogger(__name__)


class EventImageConverter(object):
    """Events -> image.  Args as src/event_image_converter.py:23-28: image_size (H, W), outer_padding."""

    def __init__(self, image_size: tuple, outer_padding: Union[int, Tuple[int, int]] = 0):
        if isinstance(outer_padding, (int, float)):
            self.outer_padding = (int(outer_padding), int(outer_padding))
        else:
            self.outer_padding = outer_padding
        self.image_size = tuple(int(i + p * 2) for i, p in zip(image_size, self.outer_padding))

    def update_property(self, image_size: Optional[tuple] = None, outer_padding=None):
        # mirrors src/event_image_converter.py:30-43 (including that it adds the padding only once)
        if image_size is not None:
            self.image_size = image_size
        if outer_padding is not None:
            if isinstance(outer_padding, int):
                self.outer_padding = (outer_padding, outer_padding)
            else:
                self.outer_padding = outer_padding
        self.image_size = tuple(i + p for i, p in zip(self.image_size, self.outer_padding))

    def create_iwe(self, events: torch.Tensor, method: str = "bilinear_vote", sigma: int = 1) -> torch.Tensor:
        """[(b,) n, >=2] -> [(b,) H, W].  src/event_image_converter.py:45-67."""
        if isinstance(events, torch.Tensor):
            return self.create_image_from_events_tensor(events, method, sigma=sigma)
        e = f"Non-supported type of events. {type(events)}"
        logger.error(e)
        raise RuntimeError(e)

    def create_eventmask(self, events: torch.Tensor) -> torch.Tensor:
        """[(b,) 1, H, W] boolean: at least one event.  src/event_image_converter.py:69-82."""
        if isinstance(events, torch.Tensor):
            return (0 != self.create_image_from_events_tensor(events, sigma=0))[..., None, :, :]
        raise RuntimeError

    def create_image_from_events_tensor(self, events: torch.Tensor, method: str = "bilinear_vote", weight=1.0, sigma: int = 0):
        """src/event_image_converter.py:126-159."""
        if method == "count":
            image = self.count_event_tensor(events)
        elif method == "bilinear_vote":
            image = self.bilinear_vote_tensor(events, weight=weight)
        else:
            e = f"{method = } is not implemented"
            logger.error(e)
            raise NotImplementedError(e)
        if sigma > 0:
            if image.dim() == 2:
                image = ops.BlurFunction.apply(image, float(sigma))
            else:
                image = torch.stack([ops.BlurFunction.apply(im, float(sigma)) for im in image], dim=0)
        return torch.squeeze(image)

    def _vote(self, events: torch.Tensor, weight, method: str) -> torch.Tensor:
        if events.dim() == 3:
            ws = weight if isinstance(weight, torch.Tensor) and weight.dim() == 2 else [weight] * events.shape[0]
            return torch.stack([self._vote(events[b], ws[b], method) for b in range(events.shape[0])], dim=0)
        w = None
        if isinstance(weight, torch.Tensor):
            w = weight.to(events.device).expand(events.shape[0]) if weight.dim() == 0 else weight
        elif float(weight) != 1.0:
            w = torch.full((events.shape[0],), float(weight), dtype=events.dtype, device=events.device)
        return ops.VoteFunction.apply(events, w, tuple(self.image_size), tuple(self.outer_padding), method)

    def bilinear_vote_tensor(self, events: torch.Tensor, weight=1.0) -> torch.Tensor:
        """src/event_image_converter.py:316-374."""
        return self._vote(events, weight, "bilinear_vote")

    def count_event_tensor(self, events: torch.Tensor) -> torch.Tensor:
        """src/event_image_converter.py:209-255."""
        return self._vote(events, 1.0, "count")
