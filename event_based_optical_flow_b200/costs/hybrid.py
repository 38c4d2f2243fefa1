"""`hybrid`: a weighted combination of other cost plugins.

Contract of the reference's HybridCost (src/costs/hybrid.py:12-79): built from `{plugin name: weight}`; `cost_func[name]`
is `{"func": plugin, "weight": w}` (the solver seam iterates over it); the loss is `sum w * plugin(arg)`, a weight of
"inv" meaning `1 / plugin(arg)`; `update_weight`; the history also carries every member's own losses.
"""
from __future__ import annotations

import logging
from typing import Dict, Iterator, Tuple

from .base import REGISTRY, CostBase

logger = logging.getLogger(__name__)


class HybridCost(CostBase, register=False):  # the reference's `functions` table does not list the combination itself
    name = "hybrid"

    def __init__(self, direction: str, cost_with_weight: Dict[str, object], store_history: bool = False, *args, **kwargs):
        logger.info(f"Log functions are mix of {cost_with_weight}")
        self.cost_func: Dict[str, dict] = {}
        for plugin, weight in cost_with_weight.items():
            member = REGISTRY[plugin](direction=direction, store_history=store_history, *args, **kwargs)
            self.cost_func[plugin] = {"func": member, "weight": weight}
        super().__init__(direction=direction, store_history=store_history)
        self.required_keys = [key for _, member, _ in self._members() for key in member.required_keys]

    def _members(self) -> Iterator[Tuple[str, CostBase, object]]:
        for plugin, entry in getattr(self, "cost_func", {}).items():
            yield plugin, entry["func"], entry["weight"]

    def update_weight(self, cost_with_weight: Dict[str, object]) -> None:
        if set(cost_with_weight) != set(self.cost_func):
            raise AssertionError(f"update_weight needs exactly the members {sorted(self.cost_func)}, got {sorted(cost_with_weight)}")
        for plugin, weight in cost_with_weight.items():
            self.cost_func[plugin]["weight"] = weight

    def _loss(self, arg: dict):
        total = 0.0
        for _, member, weight in self._members():
            value = member.calculate(arg)
            total = total + (1.0 / value if weight == "inv" else weight * value)
        return total

    def clear_history(self) -> None:
        super().clear_history()
        for _, member, _ in self._members():
            member.clear_history()

    def get_history(self) -> dict:
        merged = super().get_history()
        for plugin, member, _ in self._members():
            merged[plugin] = member.get_history()["loss"]
        return merged

    def _set_recording(self, on: bool) -> None:
        super()._set_recording(on)
        for _, member, _ in self._members():
            member.store_history = on
