"""event_based_optical_flow_b200 -- the contrast-maximization inner loop (warp -> IWE -> cost -> gradient) of
tub-rip/event_based_optical_flow as hand-written sm_100a CUDA kernels behind the reference's Python interfaces.

Layout:
  csrc/ + include/cmax_b200.h   the kernels and the C ABI (the product)
  _lib.py / _build.py           ctypes binding and the in-tree nvcc build
  ops.py                        one-call-per-operator wrappers + autograd Functions
  warp.py, event_image_converter.py, costs/   drop-in mirrors of the reference's duck-typed seam objects
  objective.py                  the fused per-iteration objective (EventPlan, ContrastObjective, cm_objective)
  solver.py                     the mixin that plugs the fused objective into the reference's solver seam
  patch_init.py                 batched 2-dof candidate costs of the pyramid's per-patch (Optuna) initialiser
  distributed.py                event sharding + the two all-reduces per CM iteration
"""
from . import _lib
from .costs import functions as cost_functions
from .event_image_converter import EventImageConverter
from .objective import COST_TABLE, ContrastObjective, EventPlan, TileFlowObjective, TimeAwareObjective, cm_objective
from .warp import MotionModelKeyError, Warp

__all__ = ["Warp", "EventImageConverter", "MotionModelKeyError", "cost_functions", "EventPlan", "ContrastObjective",
           "cm_objective", "COST_TABLE", "TileFlowObjective", "TimeAwareObjective"]
