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
  * `use_b200_operators(solver)` only swaps `solver.warper` / `solver.imager` / `solver.cost_func` for dual operators (the
    CUDA classes for CUDA tensors, the reference's own objects for numpy input); the reference's own `get_arg_for_cost`
    keeps composing them (one kernel per operator call).
  * the mixin replaces the composition itself by the fused path (3 kernels per CM iteration).
"""
from __future__ import annotations

import logging
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import costs as b200_costs
from .event_image_converter import EventImageConverter
from .objective import COST_TABLE, ContrastObjective, EventPlan
from .warp import Warp

logger = logging.getLogger(__name__)


def _has_cuda_tensor(obj) -> bool:
    if isinstance(obj, torch.Tensor):
        return obj.is_cuda
    if isinstance(obj, dict):
        return any(_has_cuda_tensor(v) for v in obj.values())
    if isinstance(obj, (list, tuple)):
        return any(_has_cuda_tensor(v) for v in obj)
    return False


class _DualOperator:
    """A seam object that runs the CUDA operator for CUDA tensors and the reference's own object for everything else.

    The reference keeps calling `solver.imager` / `solver.warper` / `solver.cost_func` with numpy arrays outside the
    scipy objective (`visualize_one_batch_warp`, `create_clipped_iwe_for_visualization`, the metrics, the Optuna
    initialiser through `calculate_cost`: src/solver/base.py:280-330, src/solver/patch_contrast_pyramid.py:320-415), so a swap
    that only understood CUDA tensors would break its run loop.  Methods dispatch per call on their arguments; attributes
    (image_size, direction, history, required_keys, ...) are the reference object's, and the two histories are kept as one."""

    def __init__(self, cuda_obj, reference_obj):
        object.__setattr__(self, "_cuda", cuda_obj)
        object.__setattr__(self, "_reference", reference_obj)
        if hasattr(reference_obj, "history") and hasattr(cuda_obj, "history"):
            cuda_obj.history = reference_obj.history  # one shared bookkeeping dict (src/costs/base.py:53-56)

    def __getattr__(self, name):
        ref = object.__getattribute__(self, "_reference")
        cuda = object.__getattribute__(self, "_cuda")
        target = getattr(ref, name) if hasattr(ref, name) else getattr(cuda, name)
        fast = getattr(cuda, name, None)
        if not callable(target) or not callable(fast):
            return target

        def dispatch(*args, **kw):
            if _has_cuda_tensor(args) or _has_cuda_tensor(kw):
                return fast(*args, **kw)
            return target(*args, **kw)

        return dispatch

    def __setattr__(self, name, value):  # e.g. solver code that toggles store_history: keep both sides in step
        for side in (object.__getattribute__(self, "_reference"), object.__getattribute__(self, "_cuda")):
            if hasattr(side, name):
                setattr(side, name, value)

    def clear_history(self):
        ref = object.__getattribute__(self, "_reference")
        ref.clear_history()
        object.__getattribute__(self, "_cuda").history = ref.history


def use_b200_operators(solver) -> None:
    """Swap the solver's duck-typed seam objects (src/solver/base.py:139-147, :185-204) for dual operators: the CUDA classes
    for CUDA tensors, the reference's own objects (kept) for numpy / CPU input."""
    image_shape = tuple(solver.imager.image_size)
    pad = tuple(getattr(solver.imager, "outer_padding", (0, 0)))
    unpadded = tuple(int(s - 2 * p) for s, p in zip(image_shape, pad))
    solver.imager = _DualOperator(EventImageConverter(unpadded, outer_padding=pad), solver.imager)
    solver.warper = _DualOperator(Warp(tuple(solver.warper.image_size), calculate_feature=getattr(solver.warper, "calculate_feature", False),
                                       normalize_t=getattr(solver.warper, "normalize_t", True),
                                       calib_param=getattr(solver.warper, "calib_param", None)), solver.warper)
    old = solver.cost_func
    if getattr(old, "name", None) == "hybrid":
        weights = {k: v["weight"] for k, v in old.cost_func.items()}
        fast = b200_costs.HybridCost(direction=old.direction, cost_with_weight=weights, store_history=old.store_history, precision="64")
    else:
        fast = b200_costs.functions[old.name](direction=old.direction, store_history=old.store_history, precision="64")
    solver.cost_func = _DualOperator(fast, old)


class _Batch:
    """One cached event batch: the tensor itself (a strong reference, so that neither its storage nor its id() can be
    recycled while the entry lives), the plans keyed by what they were packed for, the objectives."""

    def __init__(self, events: torch.Tensor):
        self.events = events
        self.version = events._version
        self.t_range = None
        self.t_scale = None
        self.plans: Dict[tuple, EventPlan] = {}
        self.objectives: Dict[tuple, ContrastObjective] = {}
        self.tile_objectives: dict = {}

    def matches(self, events: torch.Tensor) -> bool:
        return self.events is events and self.version == events._version

    def close(self) -> None:
        self.tile_objectives.clear()
        self.objectives.clear()
        for plan in self.plans.values():
            plan.close()
        self.plans.clear()


def _weighted_sum(total, term, weight):
    """total + weight * term as src/costs/hybrid.py:40-52 forms it ('inv' = the reciprocal), without the torch operators that
    change nothing (0.0 + x, 1.0 * x): at the shipped batch size every operator is ~10 us of host time, forward and backward."""
    if isinstance(weight, str):
        if weight != "inv":
            raise ValueError(f"unknown hybrid weight {weight!r}")
        term = 1.0 / term
    elif weight != 1.0:
        term = weight * term
    return term if total is None else total + term


class B200CostMixin:
    """Mix in BEFORE a reference solver class.  Needs from the host class only what the reference's own
    `get_arg_for_cost` uses: `self.cost_func`, `self.iwe_config`, `self.imager`, `self.warper`."""

    b200_event_order = "pixel"
    b200_process_group = None  # set to a torch.distributed group to shard the events of every rank
    b200_exchange = "nccl"     # "peer": the two sums over NVLink peer memory inside the kernels (needs symmetric memory)
    b200_cuda_graph = False     # True: single-GPU objectives replay a CUDA graph per evaluation (small, launch-bound batches)
    b200_fuse_tile_flow = None  # None: TileFlowObjective's default (fused when sharded); True / False force the tile-flow model on / off

    b200_fast_total_variation = True  # the hybrid's total-variation term through costs.total_variation.total_variation_loss
    b200_defer_history = True   # cost histories are materialised when read (get_history / clear_history), not with one .item() per call

    b200_max_batches = 2  # resident event batches (the current one and its predecessor); older ones are closed

    # -- cache, keyed on the IDENTITY of the event tensor (the reference hands the same tensor to every objective call of an
    #    optimize(), src/solver/patch_contrast_pyramid.py:186, 252-277).  A data_ptr()-based key would collide as soon as
    #    the caching allocator recycles the storage of the previous frame's batch.  A solver that clones its events per call
    #    (src/solver/patch_contrast_base.py:252) gets a fresh, correct plan per call.
    def _b200_cache(self) -> "OrderedDict[int, _Batch]":
        if not hasattr(self, "_b200_batches"):
            self._b200_batches: "OrderedDict[int, _Batch]" = OrderedDict()
        return self._b200_batches

    def b200_release(self) -> None:
        """Drop the resident event copies (call when `optimize()` is done with a batch)."""
        for batch in self._b200_cache().values():
            batch.close()
        self._b200_batches = OrderedDict()

    def _b200_batch(self, events: torch.Tensor) -> _Batch:
        cache = self._b200_cache()
        batch = cache.get(id(events))
        if batch is not None and not batch.matches(events):  # same object, modified in place since: rebuild
            batch.close()
            del cache[id(events)]
            batch = None
        if batch is None:
            batch = _Batch(events)
            cache[id(events)] = batch
            while len(cache) > self.b200_max_batches:
                _, old = cache.popitem(last=False)
                old.close()
        else:
            cache.move_to_end(id(events))
        return batch

    # -- the per-patch initialiser of the pyramid (src/solver/patch_contrast_pyramid.py:320-362): same studies, same sampling
    #    ranges, same candidate cost -- but trial t of ALL patches is evaluated by one batched CUDA call instead of one numpy /
    #    scipy / cv2 chain per patch and trial
    def initialize_guess_from_optuna_sampling(self, events, motion0):
        from .patch_init import PatchCandidateEvaluator, n_trials_at, run_patch_studies
        evaluator = PatchCandidateEvaluator(events, [self.patches[i] for i in range(self.n_patch)], self.scaled_patch_size[self.current_scale],
                                            outer_padding=self.padding, sigma=float(self.iwe_config["blur_sigma"]),
                                            normalize_t=self.normalize_t_in_batch)
        n_iter = self.opt_config["n_iter"]
        return run_patch_studies(evaluator, np.asarray(motion0, dtype=np.float64).reshape(2, -1), n_trials_at(n_iter, self.current_scale, self.coarest_scale),
                                 min(10, n_iter // 5), suggest=lambda trial, key, m0: self.sampling_initial(trial, key, m0))

    def _b200_image_geometry(self) -> Tuple[Tuple[int, int], Tuple[int, int]]:
        pad = tuple(int(p) for p in getattr(self.imager, "outer_padding", (0, 0)))
        padded = tuple(int(s) for s in self.imager.image_size)
        return (padded[0] - 2 * pad[0], padded[1] - 2 * pad[1]), pad

    def _b200_objective(self, events: torch.Tensor, cost_name: str, direction: str, motion_model: str, n_bins: Optional[int]):
        batch = self._b200_batch(events)
        sigma = float(self.iwe_config["blur_sigma"])
        key = (cost_name, direction, motion_model, n_bins, sigma)
        obj = batch.objectives.get(key)
        if obj is None:
            image_size, pad = self._b200_image_geometry()
            if self.b200_process_group is not None and batch.t_range is None:
                from .distributed import global_time_range
                batch.t_range = global_time_range(events, self.b200_process_group)
            # one plan per (reference times, voxel depth): two hybrid terms with different reference sets never re-pack each other's
            plan_key = (tuple(d for _, d in COST_TABLE[cost_name][2]), n_bins if motion_model == "dense-flow-voxel" else 0)
            plan = batch.plans.get(plan_key)
            if plan is None:
                plan = EventPlan(events, image_size, pad, self.b200_event_order, batch.t_range)
                batch.plans[plan_key] = plan
            obj = ContrastObjective(plan, image_size, cost=cost_name, motion_model=motion_model, sigma=sigma, omit_boundary=True,
                                    direction=direction, n_bins=n_bins, orig_events=events, process_group=self.b200_process_group,
                                    exchange=self.b200_exchange, cuda_graph=self.b200_cuda_graph)
            batch.objectives[key] = obj
        return obj

    # -- cost history without a host sync per call (src/costs/base.py:42-56 does `loss.item()` inside
    #    every `calculate`: three device round trips per objective call for a hybrid of two terms).  The losses stay on the device
    #    until somebody asks: `get_history()` / `clear_history()` of every cost object that has deferred entries are wrapped (on the
    #    INSTANCE) to materialise them first, in call order, with one D2H copy for all of them.
    def _b200_register(self, cost, loss) -> None:
        if not getattr(cost, "store_history", False):
            return
        if not (self.b200_defer_history and isinstance(loss, torch.Tensor)):
            cost.history["loss"].append(cost.get_item(loss))
            return
        self.__dict__.setdefault("_b200_pending", []).append((cost, loss.detach()))
        if cost.__dict__.get("_b200_history_owner") is not self:
            for name in ("get_history", "clear_history"):
                def flushing(*a, _inner=getattr(cost, name), **k):
                    self.b200_flush_history()
                    return _inner(*a, **k)
                setattr(cost, name, flushing)
            cost._b200_history_owner = self

    def b200_flush_history(self) -> None:
        """Append the deferred losses to their cost objects' `history["loss"]` (one device -> host copy)."""
        pending = self.__dict__.get("_b200_pending")
        if not pending:
            return
        self._b200_pending = []
        values = torch.stack([loss.double().reshape(()) for _, loss in pending]).cpu().tolist()
        for (cost, _), value in zip(pending, values):
            cost.history["loss"].append(value)

    def _b200_unrecorded(self, cost, arg: dict):
        """`cost.calculate(arg)` of a plugin that stays in torch (total variation) with ITS `.item()` deferred too."""
        if not (self.b200_defer_history and getattr(cost, "store_history", False)):
            return cost.calculate(arg)
        cost.store_history = False
        try:
            loss = cost.calculate(arg)
        finally:
            cost.store_history = True
        self._b200_register(cost, loss)
        return loss

    def _b200_total_variation(self, cost, flow):
        """The total-variation term of a hybrid cost for a tensor flow: value + analytic gradient in 8 torch operators
        (costs/total_variation.py) instead of ~200 through the plugin's two Conv2d modules and autograd; same numbers."""
        if not (self.b200_fast_total_variation and isinstance(flow, torch.Tensor) and flow.dim() in (3, 4) and flow.shape[-3] == 2):
            return self._b200_unrecorded(cost, {"flow": flow, "omit_boundary": True})
        loss = b200_costs.total_variation.total_variation_loss(flow, True, cost.direction)
        self._b200_register(cost, loss)
        return loss

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
            return self._b200_total_variation(cost, coarse_flow)
        return None

    # -- the pyramid's objective with the tile-flow map INSIDE the event kernels
    def objective_scipy(self, motion_array, *args, **kw):
        """Same contract as `PyramidalPatchContrastMaximization.objective_scipy(motion_array, events, coarser_motion,
        suppress_log)` (src/solver/patch_contrast_pyramid.py:430-462).  For the configuration that method spends its time in --
        CUDA events, bilinear patch interpolation, not time-aware -- the loss is evaluated from the patch motion directly through
        `TileFlowObjective`: fused, the strip kernels compute interpolate(motion) * t_scale per source
        pixel and return dL/d(patch motion), no dense flow or dense gradient is ever built; composed, the up-sampling and its
        adjoint are two small kernels around the dense model.  Anything else goes to the reference's own method."""
        events = args[0] if len(args) >= 1 else kw.get("events")
        pyramid_call = len(args) >= 2 and isinstance(args[1], dict) or "coarser_motion" in kw
        fast = (pyramid_call and isinstance(events, torch.Tensor) and events.is_cuda and isinstance(motion_array, torch.Tensor)
                and not getattr(self, "is_time_aware", False) and getattr(self, "filter_type", "bilinear") == "bilinear"
                and self.iwe_config.get("method", "bilinear_vote") == "bilinear_vote"
                and getattr(self, "motion_model_for_dense_warp", None) == "dense-flow" and int(getattr(self, "motion_vector_size", 2)) == 2)
        if not fast:
            return super().objective_scipy(motion_array, *args, **kw)
        batch = self._b200_batch(events)
        if getattr(self, "normalize_t_in_batch", False):
            if batch.t_scale is None:  # events are constant over an optimize(): one host sync per batch instead of one per call
                t = events.detach()[:, 2]
                batch.t_scale = float(t.max() - t.min())
            t_scale = batch.t_scale
        else:
            t_scale = 1.0
        motion = motion_array.reshape((self.motion_vector_size,) + tuple(self.patch_image_size))
        if motion.device != events.device:
            motion = motion.to(events.device)
        geometry = (tuple(self.patch_size), tuple(self.sliding_window), tuple(self.patch_shift), t_scale)

        def term(cost):
            name = getattr(cost, "name", None)
            if name in COST_TABLE:
                obj = self._b200_objective(events, name, cost.direction, "dense-flow", None)
                key = (name, cost.direction) + geometry
                tile = batch.tile_objectives.get(key)
                if tile is None:
                    from .objective import TileFlowObjective
                    fuse = self.b200_fuse_tile_flow
                    if fuse and obj.plan.n_strips == 0:
                        fuse = False  # (a sparse batch has no strips: the composition is the only form)
                    tile = TileFlowObjective(obj, geometry[0], geometry[1], geometry[2], t_scale, fused=fuse)
                    batch.tile_objectives[key] = tile
                loss = tile(motion)
                self._b200_register(cost, loss)
                return loss
            if name == "total_variation":
                return self._b200_total_variation(cost, motion)
            return None

        cost = self.cost_func
        if getattr(cost, "name", None) == "hybrid":
            loss = None
            for entry in cost.cost_func.values():
                t = term(entry["func"])
                if t is None:
                    self.b200_flush_history()
                    return super().objective_scipy(motion_array, *args, **kw)
                loss = _weighted_sum(loss, t, entry["weight"])
            self._b200_register(cost, loss)
        else:
            loss = term(cost)
            if loss is None:
                return super().objective_scipy(motion_array, *args, **kw)
        if not (args[2] if len(args) >= 3 else kw.get("suppress_log", False)) and logger.isEnabledFor(logging.INFO):
            logger.info(f"{loss = }")  # (formatting a CUDA tensor is a host sync: only when somebody listens)
        return loss

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
        if not fusable:  # numpy callers (metrics, visualisation) keep the reference path -- which registers its history at once
            self.b200_flush_history()
            return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
        if warp.device != events.device:
            warp = warp.to(events.device)
        cost = self.cost_func
        if getattr(cost, "name", None) == "hybrid":
            loss = None
            for name, entry in cost.cost_func.items():
                term = self._b200_term(entry["func"], events, warp, motion_model, coarse_flow)
                if term is None:
                    self.b200_flush_history()
                    return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
                loss = _weighted_sum(loss, term, entry["weight"])
            self._b200_register(cost, loss)
            return loss
        loss = self._b200_term(cost, events, warp, motion_model, coarse_flow)
        if loss is None:
            self.b200_flush_history()
            return super().calculate_cost(events, warp, motion_model, coarse_flow, save_intermediate_result)
        return loss
