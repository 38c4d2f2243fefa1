"""Plugging the fused CUDA objective into the reference's solver seam.

The reference composes one objective evaluation in `PatchContrastMaximization.calculate_cost` /
`get_arg_for_cost` (src/solver/patch_contrast_base.py:273-352): up to three warps + four IWEs + the cost plugin, every
call, differentiated by torch autograd.  `B200CostMixin.calculate_cost` has the same signature and the same return value
(a 0-dim tensor in the dtype of `warp`, connected to `warp` in the autograd graph), but evaluates every contrast
term with the fused kernels and keeps the event batch resident between calls.  Everything around it -- the pyramid, the
tile-flow upsample, the scipy / optuna drivers -- is the reference's own, unchanged code:

    from src.solver import PyramidalPatchContrastMaximization          # the reference
    from event_based_optical_flow_b200.solver import B200CostMixin, use_b200_operators

    class B200Pyramidal(B200CostMixin, PyramidalPatchContrastMaximization):
        pass
    solver.collections["b200_pyramidal_patch_contrast_maximization"] = B200Pyramidal   # selectable from the YAML

Two levels of drop-in exist:
  * `use_b200_operators(solver)` only swaps `solver.warper` / `solver.imager` / `solver.cost_func` for the CUDA
    operator classes; the reference's own `get_arg_for_cost` keeps composing them (one kernel per operator call).
  * the mixin replaces the composition itself by the fused path (3 kernels + 1 per CM iteration).
"""
from __future__ import annotations

import logging
from typing import Dict, Optional, Tuple

import torch

from . import costs as b200_costs
from .event_image_converter import EventImageConverter
from .objective import COST_TABLE, ContrastObjective, EventPlan
from .warp import Warp

logger = logging.getLogger(__name__)


def use_b200_operators(solver) -> None:
    """Swap the solver's duck-typed seam objects (src/solver/base.py:139-147, :185-204) for the CUDA ones."""
    image_shape = tuple(solver.imager.image_size)
    pad = tuple(getattr(solver.imager, "outer_padding", (0, 0)))
    unpadded = tuple(int(s - 2 * p) for s, p in zip(image_shape, pad))
    solver.imager = EventImageConverter(unpadded, outer_padding=pad)
    solver.warper = Warp(tuple(solver.warper.image_size), calculate_feature=getattr(solver.warper, "calculate_feature", False),
                         normalize_t=getattr(solver.warper, "normalize_t", True), calib_param=getattr(solver.warper, "calib_param", None))
    old = solver.cost_func
    if getattr(old, "name", None) == "hybrid":
        weights = {k: v["weight"] for k, v in old.cost_func.items()}
        solver.cost_func = b200_costs.HybridCost(direction=old.direction, cost_with_weight=weights, store_history=old.store_history,
                                                 precision="64")
    else:
        solver.cost_func = b200_costs.functions[old.name](direction=old.direction, store_history=old.store_history, precision="64")


class B200CostMixin:
    """Mix in BEFORE a reference solver class.  Needs from the host class only what the reference's own
    `get_arg_for_cost` uses: `self.cost_func`, `self.iwe_config`, `self.imager`, `self.warper`."""

    b200_event_order = "pixel"
    b200_process_group = None  # set to a torch.distributed group to shard the events of every rank

    # -- cache: one resident plan per event tensor, one fused objective per (plan, cost term, motion model)
    def _b200_cache(self) -> dict:
        if not hasattr(self, "_b200_objectives"):
            self._b200_objectives: Dict[tuple, ContrastObjective] = {}
            self._b200_plans: Dict[tuple, EventPlan] = {}
        return self._b200_objectives

    def b200_release(self) -> None:
        """Drop the resident event copies (call when `optimize()` is done with a batch)."""
        self._b200_objectives = {}
        self._b200_plans = {}

    def _b200_image_geometry(self) -> Tuple[Tuple[int, int], Tuple[int, int]]:
        pad = tuple(int(p) for p in getattr(self.imager, "outer_padding", (0, 0)))
        padded = tuple(int(s) for s in self.imager.image_size)
        return (padded[0] - 2 * pad[0], padded[1] - 2 * pad[1]), pad

    def _b200_objective(self, events: torch.Tensor, cost_name: str, direction: str, motion_model: str, n_bins: Optional[int]):
        cache = self._b200_cache()
        ev_key = (events.data_ptr(), tuple(events.shape), events._version, str(events.device))
        key = ev_key + (cost_name, direction, motion_model, n_bins, float(self.iwe_config["blur_sigma"]))
        obj = cache.get(key)
        if obj is None:
            if len(self._b200_plans) > 4:  # a new batch arrived: forget the old ones
                self.b200_release()
                cache = self._b200_cache()
            image_size, pad = self._b200_image_geometry()
            plan = self._b200_plans.get(ev_key)
            t_range = None
            if self.b200_process_group is not None:
                from .distributed import global_time_range
                t_range = global_time_range(events, self.b200_process_group)
            if plan is None:
                plan = EventPlan(events, image_size, pad, self.b200_event_order, t_range)
                self._b200_plans[ev_key] = plan
            obj = ContrastObjective(plan, image_size, cost=cost_name, motion_model=motion_model, sigma=float(self.iwe_config["blur_sigma"]),
                                    omit_boundary=True, direction=direction, n_bins=n_bins, orig_events=events,
                                    process_group=self.b200_process_group)
            cache[key] = obj
        return obj

    @staticmethod
    def _b200_register(cost, loss) -> None:
        if getattr(cost, "store_history", False):
            cost.history["loss"].append(cost.get_item(loss))

    def _b200_term(self, cost, events, warp, motion_model, coarse_flow):
        """One (non-hybrid) cost plugin evaluated for this call, or None if it has no fused form here."""
        name = getattr(cost, "name", None)
        if name in COST_TABLE:
            n_bins = int(warp.shape[0]) if motion_model == "dense-flow-voxel" else None
            obj = self._b200_objective(events, name, cost.direction, motion_model, n_bins)
            loss = obj(warp)
            self._b200_register(cost, loss)
            return loss
        if name == "total_variation":
            return cost.calculate({"flow": coarse_flow, "omit_boundary": True})
        return None

    def interpolate_dense_flow_from_patch_tensor(self, motion_array: torch.Tensor) -> torch.Tensor:
        """Same contract as src/solver/patch_contrast_base.py:462-506, one CUDA kernel (and one for the adjoint) instead
        of pad + torchvision resize + crop.  Falls back to the reference for CPU tensors / the 'nearest' filter."""
        if not (isinstance(motion_array, torch.Tensor) and motion_array.is_cuda and getattr(self, "filter_type", "bilinear") == "bilinear"):
            return super().interpolate_dense_flow_from_patch_tensor(motion_array)
        from . import ops
        pad = ops.tile_flow_geometry(self.image_shape, self.patch_size, self.sliding_window, self.patch_shift)
        m = motion_array.reshape((self.motion_vector_size,) + tuple(self.patch_image_size))
        return ops.TileFlowFunction.apply(m, tuple(self.image_shape), pad, tuple(self.sliding_window))

    def motion_to_dense_flow(self, motion, t_scale: float = 1.0):
        """Same contract as the reference's two `motion_to_dense_flow` methods -- the pyramid's
        (src/solver/patch_contrast_pyramid.py:464-516: dict of per-scale motions, `t_scale`, scale = max of the dense flow)
        and the time-aware solver's (src/solver/time_aware_patch_contrast.py:42-80: one motion array, scale = max of the
        motion) -- with the upwind / Burgers voxel propagation (src/utils/flow_utils.py:99-161) as ONE CUDA launch per side
        of t0 instead of ~25 torch kernels per time bin (and as many again in autograd).  Anything else (numpy callers,
        CPU tensors, the other interpolation schemes) goes to the reference's own method."""
        is_pyramid = isinstance(motion, dict)
        finest = motion[self.current_scale] if is_pyramid else motion
        fast = (isinstance(finest, torch.Tensor) and finest.is_cuda and getattr(self, "is_time_aware", False)
                and getattr(self, "flow_interpolation", None) in ("upwind", "burgers"))
        if not fast:
            return super().motion_to_dense_flow(motion, t_scale) if is_pyramid else super().motion_to_dense_flow(motion)
        from . import ops
        dense = self.interpolate_dense_flow_from_patch_tensor(finest)
        if is_pyramid:
            scale = dense.max() if self.scale_later else 1.0
            voxel = ops.FlowVoxelFunction.apply(dense * t_scale / scale, self.time_bin, self.flow_interpolation, self.t0_flow_location)
            return voxel * scale / t_scale
        scale = finest.max() if self.scale_later else 1.0
        voxel = ops.FlowVoxelFunction.apply(dense / scale, self.time_bin, self.flow_interpolation, self.t0_flow_location)
        return voxel * scale

    def calculate_cost(self, events, warp, motion_model: str, coarse_flow=None, save_intermediate_result: bool = True):
        """Same contract as src/solver/patch_contrast_base.py:273-287."""
        fusable = (isinstance(events, torch.Tensor) and events.is_cuda and isinstance(warp, torch.Tensor)
                   and self.iwe_config.get("method", "bilinear_vote") == "bilinear_vote"
                   and motion_model in ("dense-flow", "dense-flow-voxel", "2d-translation", "rigid-optical-flow"))
        if not fusable:  # numpy callers (metrics, visualisation, the Optuna initialiser) keep the reference path
            return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
        if warp.device != events.device:
            warp = warp.to(events.device)
        cost = self.cost_func
        if getattr(cost, "name", None) == "hybrid":
            loss = 0.0
            for name, entry in cost.cost_func.items():
                term = self._b200_term(entry["func"], events, warp, motion_model, coarse_flow)
                if term is None:
                    return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
                loss = loss + (1.0 / term if entry["weight"] == "inv" else entry["weight"] * term)
            self._b200_register(cost, loss)
            return loss
        loss = self._b200_term(cost, events, warp, motion_model, coarse_flow)
        if loss is None:
            return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
        return loss
