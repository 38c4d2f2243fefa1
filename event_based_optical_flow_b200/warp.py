"""Drop-in `Warp` (reference: src/warp.py:24-522) running on the sm_100a kernels.

Same constructor, same `warp_event(events, motion, motion_model, direction)` -> `(warped, feature_dict)` contract,
same exceptions.  Tensors must live on a CUDA device; the arithmetic is fp32 (bit-exact with the reference's torch
branch run in fp32), results are returned in the dtype of `events`.
"""
from __future__ import annotations

import logging
from typing import Optional, Tuple, Union

import numpy as np
import torch

from . import ops

logger = logging.getLogger(__name__)


class MotionModelKeyError(Exception):
    """Unknown motion model (reference: src/warp.py:15-21)."""

    def __init__(self, message):
        e = f"{message = } not supported"
        logger.error(e)
        super().__init__(e)


class FeatureCalculatorMock:
    """src/feature_calculator.py:8-22 -- the second return value of every warp."""

    def __init__(self, *args, **kwargs):
        self.skip = True

    def calculate_feature(self, *args, **kwargs) -> dict:
        return {"none": None}


_DIRECTIONS = ("first", "middle", "last", "random", "before", "after")


def _resolve_direction(direction):
    if type(direction) is float:
        return direction
    if direction == "random":
        return float(np.random.uniform(low=0.0, high=1.0))
    if direction in _DIRECTIONS:
        return direction
    e = f"direction argument should be first, middle, last. Or float. {direction}"
    logger.error(e)
    raise ValueError(e)


# motion models whose parameters are the two translation components (the only parametric family of this release of the
# reference, src/warp.py:64-128); "dense-flow" is accepted by the bookkeeping helpers as "a rigid translation everywhere"
_TRANSLATION_MODELS = ("2d-translation", "rigid-optical-flow")
_TRANSLATION_KEYS = ("trans_x", "trans_y")


def _check_parametric(motion_model: str, allow_dense: bool) -> None:
    if motion_model in _TRANSLATION_MODELS:
        return
    if allow_dense and motion_model == "dense-flow":
        logger.warning(f"Assume only rigid transformation {motion_model = }, not meaningful.")
        return
    raise MotionModelKeyError(motion_model)


class Warp:
    """Event warping on the CUDA kernels.  Constructor arguments as the reference's (src/warp.py:35-44):
    image_size (H, W), calculate_feature, normalize_t, calib_param."""

    _SETTABLE = ("image_size", "calculate_feature", "normalize_t", "calib_param")

    def __init__(self, image_size: tuple, calculate_feature: bool = False, normalize_t: bool = False,
                 calib_param: Optional[np.ndarray] = None):
        self.update_property(image_size, calculate_feature, normalize_t, calib_param)
        self.feature_2dof = self.feature_dense = FeatureCalculatorMock()

    def update_property(self, image_size=None, calculate_feature=None, normalize_t=None, calib_param=None):
        """Set whichever properties are given (None = keep)."""
        for attr, value in zip(self._SETTABLE, (image_size, calculate_feature, normalize_t, calib_param)):
            if value is None:
                continue
            if attr == "calib_param":
                logger.info("Set camera matrix K.")
            setattr(self, attr, value)

    # -- host-side bookkeeping between parameter dicts, motion vectors and flow fields (src/warp.py:64-153)
    def get_key_names(self, motion_model: str) -> list:
        _check_parametric(motion_model, allow_dense=True)
        return list(_TRANSLATION_KEYS)

    def get_motion_vector_size(self, motion_model: str) -> int:
        zeros = dict.fromkeys(self.get_key_names(motion_model), 0.0)
        return len(self.motion_model_to_motion(motion_model, zeros))

    def motion_model_to_motion(self, motion_model: str, params: dict) -> np.ndarray:
        _check_parametric(motion_model, allow_dense=True)
        theta = np.array([params[k] for k in _TRANSLATION_KEYS])
        return self.get_flow_from_motion(theta, "2d-translation") if motion_model == "dense-flow" else theta

    def motion_model_from_motion(self, motion: np.ndarray, motion_model: str) -> dict:
        if motion_model != "dense-flow":
            _check_parametric(motion_model, allow_dense=False)
        return {k: motion[i] for i, k in enumerate(_TRANSLATION_KEYS)}

    def get_flow_from_motion(self, motion, motion_model: str):
        """Dense flow [2,H,W] equivalent to a parametric motion: the displacement a unit-dt event undergoes
        (src/warp.py:130-153).  For the 2-dof model this is the constant field -theta."""
        _check_parametric(motion_model, allow_dense=False)
        H, W = self.image_size
        if isinstance(motion, torch.Tensor):
            return -(motion.reshape(2, 1, 1).expand(2, H, W)).clone()
        m = np.asarray(motion, dtype=np.float64).reshape(2, 1, 1)
        return -np.broadcast_to(m, (2, H, W)).copy()

    # -- the hot-path entry point
    def warp_event(self, events: torch.Tensor, motion: torch.Tensor, motion_model: str,
                   direction: Union[str, float] = "first", flow_propagate_bin: Optional[int] = None) -> Tuple[torch.Tensor, dict]:
        """events [(b,) n, C>=3], motion as the model needs -> (warped [(b,) n, C] = (x', y', dt, p), feature dict).
        src/warp.py:156-199."""
        direction = _resolve_direction(direction)
        if motion_model == "dense-flow-voxel-optimized":
            # dead code in the reference (reads an undefined attribute, src/warp.py:422); no solver selects it
            raise MotionModelKeyError(motion_model)
        if motion_model not in ("dense-flow", "dense-flow-voxel", "2d-translation", "rigid-optical-flow"):
            raise MotionModelKeyError(motion_model)
        if not isinstance(events, torch.Tensor):
            raise RuntimeError("the B200 Warp takes CUDA torch tensors; use the reference's numpy branch on the host")
        if motion_model in ("2d-translation", "rigid-optical-flow"):
            assert motion.shape[-1] == 2
        if events.dim() == 3:
            outs = [self._warp_one(events[b], motion[b], motion_model, direction) for b in range(events.shape[0])]
            warped = torch.stack(outs, dim=0)
        else:
            warped = self._warp_one(events, motion, motion_model, direction)
        feat = (self.feature_2dof if motion.shape[-1] == 2 and motion.dim() <= 2 else self.feature_dense).calculate_feature()
        return warped.squeeze(), feat

    def _warp_one(self, events, motion, motion_model, direction):
        if not isinstance(motion, torch.Tensor):
            motion = torch.as_tensor(np.asarray(motion), device=events.device)
        motion = motion.to(events.device)
        H, W = int(self.image_size[0]), int(self.image_size[1])
        if motion_model == "dense-flow":
            if motion.dim() != 3 or motion.shape[0] != 2:
                raise ValueError(f"dense-flow motion must be [2,H,W], got {tuple(motion.shape)}")
            H, W = int(motion.shape[1]), int(motion.shape[2])
        n_bins = 0
        if motion_model == "dense-flow-voxel":
            if motion.dim() != 4 or motion.shape[1] != 2:
                raise ValueError(f"dense-flow-voxel motion must be [T,2,H,W], got {tuple(motion.shape)}")
            n_bins, H, W = int(motion.shape[0]), int(motion.shape[2]), int(motion.shape[3])
        if events.shape[0] == 0:
            return events.clone()
        tp = ops.time_params(events, (direction,), n_bins, self.normalize_t)
        return ops.WarpFunction.apply(events, motion, motion_model, (H, W), tp, 0)

    # reference-time helpers kept for API compatibility (src/warp.py:201-259); torch ops on 0-dim tensors
    def calculate_reftime(self, events: torch.Tensor, direction: Union[str, float] = "first"):
        direction = _resolve_direction(direction)
        t = events[..., 2]
        lo, hi = t.min(-1).values if t.dim() > 1 else t.min(), t.max(-1).values if t.dim() > 1 else t.max()
        if type(direction) is float:
            return lo + (hi - lo) * direction
        if direction == "first":
            return lo
        if direction == "last":
            return hi
        return lo + (hi - lo) * {"middle": 0.5, "before": -1.0, "after": 2.0}[direction]

    def calculate_dt(self, event: torch.Tensor, reference_time, time_period=None):
        dt = event[..., 2] - reference_time
        if self.normalize_t:
            if time_period is None:
                time_period = (dt.max(-1).values - dt.min(-1).values) if dt.dim() > 1 else dt.max() - dt.min()
            dt = dt / (time_period[..., None] if isinstance(time_period, torch.Tensor) and time_period.dim() > 0 else time_period)
        return dt
