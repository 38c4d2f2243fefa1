"""Cost plugin base (reference: src/costs/base.py:11-77): direction check, required_keys contract, KeyError logging,
loss history."""
from __future__ import annotations

import logging
from typing import Dict, List

import torch

logger = logging.getLogger(__name__)


class CostBase(object):
    """direction: 'minimize' | 'maximize' | 'natural' (src/costs/base.py:20-25)."""

    required_keys: List[str] = []

    def __init__(self, direction="minimize", store_history: bool = False, *args, **kwargs):
        if direction not in ["minimize", "maximize", "natural"]:
            e = f"direction should be minimize, maximize, and natural. Got {direction}."
            logger.error(e)
            raise ValueError(e)
        self.direction = direction
        self.store_history = store_history
        self.clear_history()

    def catch_key_error(func):
        def wrapper(self, arg: dict):
            try:
                return func(self, arg)
            except KeyError as e:
                logger.error("Input for the cost needs keys of:")
                logger.error(self.required_keys)
                raise e

        return wrapper

    def register_history(func):
        def wrapper(self, arg: dict):
            loss = func(self, arg)
            if self.store_history:
                self.history["loss"].append(self.get_item(loss))  # one host sync per call, as in the reference
            return loss

        return wrapper

    def get_item(self, loss) -> float:
        if isinstance(loss, torch.Tensor):
            return loss.item()
        return loss

    def clear_history(self) -> None:
        self.history: Dict[str, list] = {"loss": []}

    def get_history(self) -> dict:
        return self.history.copy()

    def enable_history_register(self) -> None:
        self.store_history = True

    def disable_history_register(self) -> None:
        self.store_history = False

    @register_history
    @catch_key_error
    def calculate(self, arg: dict):
        raise NotImplementedError

    catch_key_error = staticmethod(catch_key_error)
    register_history = staticmethod(register_history)


def require_cuda_image(iwe, who: str) -> torch.Tensor:
    if not isinstance(iwe, torch.Tensor):
        e = f"Unsupported input type. {type(iwe)}."
        logger.error(e)
        raise NotImplementedError(e)
    if not iwe.is_cuda:
        raise RuntimeError(f"{who}: the B200 cost plugins take CUDA tensors (no CPU fallback)")
    return iwe
