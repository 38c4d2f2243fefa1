"""Base of the cost plugins.

Public contract (what the reference's solvers rely on, src/costs/base.py:11-77): construct with
`(direction, store_history)`; class attributes `name` and `required_keys`; `calculate(arg: dict)` returns the loss;
a missing arg-dict key is logged together with the required keys and re-raised as KeyError; with `store_history` every
loss is appended to `history["loss"]` as a Python float; `get_history / clear_history / enable_history_register /
disable_history_register`.

Design here: a plugin only writes `_loss(self, arg)`.  `__init_subclass__` registers every named plugin in `REGISTRY` (the
`costs.functions` table) and `calculate` is ONE template method on the base that adds the key check and the bookkeeping,
so no plugin has to decorate anything.
"""
from __future__ import annotations

import logging
from typing import Callable, Dict, List

import torch

logger = logging.getLogger(__name__)

DIRECTIONS = ("minimize", "maximize", "natural")
REGISTRY: Dict[str, type] = {}


class CostBase:
    required_keys: List[str] = []

    def __init_subclass__(cls, register: bool = True, **kw):
        super().__init_subclass__(**kw)
        plugin_name = cls.__dict__.get("name")
        if register and plugin_name is not None:
            REGISTRY[plugin_name] = cls

    def __init__(self, direction: str = "minimize", store_history: bool = False, *_, **__):
        if direction not in DIRECTIONS:
            msg = f"direction should be minimize, maximize, and natural. Got {direction}."  # the reference's wording (base.py:22)
            logger.error(msg)
            raise ValueError(msg)
        self.direction = direction
        self.store_history = bool(store_history)
        self.history: Dict[str, list] = {}
        self.clear_history()

    # ---- what a plugin implements
    def _loss(self, arg: dict):
        raise NotImplementedError(f"{type(self).__name__} does not implement a loss")

    # ---- the public entry point
    def calculate(self, arg: dict):
        try:
            loss = self._loss(arg)
        except KeyError:
            logger.error("Input for the cost needs keys of:")
            logger.error(self.required_keys)
            raise
        if self.store_history:
            self.history["loss"].append(self.get_item(loss))  # one host sync per call, like the reference's bookkeeping
        return loss

    @staticmethod
    def get_item(loss) -> float:
        return loss.item() if isinstance(loss, torch.Tensor) else loss

    # ---- history bookkeeping
    def clear_history(self) -> None:
        self.history = {"loss": []}

    def get_history(self) -> dict:
        return dict(self.history)

    def _set_recording(self, on: bool) -> None:
        self.store_history = on

    def enable_history_register(self) -> None:
        self._set_recording(True)

    def disable_history_register(self) -> None:
        self._set_recording(False)


def require_cuda_image(iwe, who: str) -> torch.Tensor:
    if not isinstance(iwe, torch.Tensor):
        msg = f"Unsupported input type. {type(iwe)}."
        logger.error(msg)
        raise NotImplementedError(msg)
    if not iwe.is_cuda:
        raise RuntimeError(f"{who}: the B200 cost plugins take CUDA tensors (no CPU fallback)")
    return iwe
