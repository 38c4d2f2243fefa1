"""ctypes binding of include/cmax_b200.h (the C ABI is the product boundary; this file is the thin binding).

Fails loudly: if the shared library is missing (not built) `load()` raises -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import _build

ABI_VERSION = 2
PATCH_GLOBAL_IMAGES, PATCH_KEEP_IMAGES = 1, 2  # flags of cmax_patch_candidates
MAX_REFS = 4
MAX_BINS = 64
MAX_PEERS = 8

OK, ERR_ARG, ERR_CUDA, ERR_SOURCE_OOB, ERR_WORKSPACE = 0, 1, 2, 3, 4
MOTION = {"dense-flow": 0, "dense-flow-voxel": 1, "2d-translation": 2, "rigid-optical-flow": 2, "tile-flow": 3}
VOTE = {"bilinear_vote": 0, "count": 1}
STAT = {"variance": 0, "gradmag": 1}
FORM = {"plain": 0, "normalized": 1, "multifocal": 2}
ORDER = {"asis": 0, "tile": 1, "pixel": 2}
SCHEME = {"upwind": 0, "burgers": 1}


class Ref(C.Structure):
    _fields_ = [("mode", C.c_int32), ("fraction", C.c_float)]


class Peers(C.Structure):
    """cmax_peers: symmetric-memory addresses of every rank's partial IWE stack, partial gradient and flag block."""
    _fields_ = [("n_peers", C.c_int32), ("rank", C.c_int32), ("iwe", C.c_void_p * MAX_PEERS), ("grad", C.c_void_p * MAX_PEERS),
                ("flags", C.c_void_p * MAX_PEERS)]


class CostSpec(C.Structure):
    _fields_ = [("stat", C.c_int32), ("form", C.c_int32), ("direction_sign", C.c_int32), ("omit_boundary", C.c_int32),
                ("sigma", C.c_float), ("weights", C.c_float * MAX_REFS)]


TIME_PARAMS_BYTES = 4 * (4 * MAX_REFS + MAX_REFS * (MAX_BINS + 1) + 4)

_p, _i, _i64, _f, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t
_SIGNATURES = {
    # name: (restype, argtypes)   -- order and meaning exactly as in include/cmax_b200.h
    "cmax_abi_version": (_i, []),
    "cmax_last_error": (C.c_char_p, []),
    "cmax_build_arch": (C.c_char_p, []),
    "cmax_time_range": (_i, [_p, _i64, _i, _p, _p]),
    "cmax_time_params": (_i, [_p, C.POINTER(Ref), _i, _i, _i, _p, _p]),
    "cmax_warp_events": (_i, [_p, _i64, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p]),
    "cmax_warp_events_backward": (_i, [_p, _i64, _i, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    "cmax_vote": (_i, [_p, _i64, _i, _p, _i, _i, _i, _i, _i, _p, _p]),
    "cmax_vote_backward": (_i, [_p, _i64, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "cmax_vote_backward2": (_i, [_p, _i64, _i, _p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p]),
    "cmax_warp_events_tangent": (_i, [_p, _i64, _i, _i, _i, _i, _p, _i, _p, _p, _p]),
    "cmax_blur3": (_i, [_p, _p, _i, _i, _i, _f, _i, _p]),
    "cmax_tile_flow_upsample": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "cmax_tile_flow_upsample_backward": (_i, [_p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
    "cmax_patch_candidates_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "cmax_patch_candidates": (_i, [_p, _p, _i64, _i, _p, _p, _i, _i, _i, _i, _i, _f, _p, _i, _p, _sz, _p, _p]),
    "cmax_flow_voxel_workspace_bytes": (_sz, [_i, _i]),
    "cmax_flow_voxel": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "cmax_flow_voxel_backward": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "cmax_stats_workspace_bytes": (_sz, [_i, _i, _i]),
    "cmax_image_stats": (_i, [_p, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "cmax_plan_workspace_bytes": (_sz, [_i64, _i, _i, _i]),
    "cmax_plan_create": (_i, [C.POINTER(_p), _p, _i64, _i, _i, _i, _i, _i, _f, _f, _i, _p, _sz, _p]),
    "cmax_plan_destroy": (None, [_p]),
    "cmax_plan_info": (_i, [_p, C.POINTER(_f), C.POINTER(_f), C.POINTER(_i64), C.POINTER(C.c_int32)]),
    "cmax_plan_strips": (_i, [_p, C.POINTER(_i64)]),
    "cmax_plan_set_refs": (_i, [_p, C.POINTER(Ref), _i, _i, _p]),
    "cmax_plan_set_tile_flow": (_i, [_p, _i, _i, _i, _i, _i, _i, _f]),
    "cmax_plan_set_variant": (_i, [_p, _i, _i]),
    "cmax_plan_set_stage_mask": (_i, [_p, _i]),
    "cmax_plan_set_compact": (_i, [_p, _i, C.POINTER(C.c_int32), _p]),
    "cmax_objective_workspace_bytes": (_sz, [_p, C.POINTER(CostSpec)]),
    "cmax_objective_workspace_init": (_i, [_p, _p, _p]),
    "cmax_objective_vote": (_i, [_p, _i, _p, _p, _p]),
    "cmax_objective_fold": (_i, [_p, _p, C.POINTER(_p), _p]),
    "cmax_objective_cost": (_i, [_p, C.POINTER(CostSpec), _p, _p, _i, _p, _p, _i64, _p]),
    "cmax_objective_grad": (_i, [_p, _i, _p, _p, _p, _i, _p]),
    "cmax_objective": (_i, [_p, _i, _p, C.POINTER(CostSpec), _p, _p, _p, _p, _p]),
    "cmax_objective_iwe_offset": (_sz, [_p]),
    "cmax_objective_full_iwe_offset": (_sz, [_p]),
    "cmax_objective_probe_offset": (_sz, [_p]),
    "cmax_objective_sharded": (_i, [_p, _i, _p, C.POINTER(CostSpec), _p, _p, C.POINTER(Peers), _p, _p, _p]),
    "cmax_combine_cost": (_i, [_p, _i, _i, _i, _p, C.POINTER(_f), _i, _i, _p, _p, _p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lock = threading.Lock()
_lib = None


class CmaxError(RuntimeError):
    """A cmax_* entry point returned a non-zero status."""

    def __init__(self, fn: str, status: int, message: str):
        self.status = status
        super().__init__(f"{fn} failed with status {status}: {message}")


_use_measure = False


def use_measure_library() -> None:
    """Probe scripts only: load lib/libcmax_b200_measure.so (`python -m event_based_optical_flow_b200._build --measure`)
    instead of the release library.  Must be called before the first `load()`; the product never calls it."""
    global _use_measure
    if _lib is not None:
        raise RuntimeError("use_measure_library() must be called before the library is loaded")
    _use_measure = True


def library_path() -> str:
    return _build.MEASURE_LIB_PATH if _use_measure else _build.LIB_PATH


def load() -> C.CDLL:
    """dlopen lib/libcmax_b200.so and declare every prototype.  Raises if the library has not been built."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -m event_based_optical_flow_b200._build` "
                "(or __graft_entry__.build()).  There is no CPU / torch fallback for the contrast-maximization path.")
        lib = C.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header and library out of sync
            fn.restype = res
            fn.argtypes = args
        if lib.cmax_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {lib.cmax_abi_version()} != {ABI_VERSION} (stale build: run `python -m event_based_optical_flow_b200._build --force`)")
        _lib = lib
        return lib


def check(fn: str, status: int) -> None:
    """Map a cmax_status to the exception type the reference raises in the same situation."""
    if status == OK:
        return
    msg = load().cmax_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(f"{fn}: {msg}")
    if status == ERR_SOURCE_OOB:
        raise IndexError(f"{fn}: {msg}")  # torch.gather raises on the reference path (src/warp.py:305-307)
    raise CmaxError(fn, status, msg)


def call(name: str, *args) -> None:
    check(name, getattr(load(), name)(*args))


def refs_array(directions) -> "C.Array[Ref]":
    """Reference-time selectors from the reference's `direction` vocabulary (src/warp.py:201-233)."""
    arr = (Ref * MAX_REFS)()
    if len(directions) < 1 or len(directions) > MAX_REFS:
        raise ValueError(f"between 1 and {MAX_REFS} reference times are supported, got {len(directions)}")
    for k, d in enumerate(directions):
        if isinstance(d, float):
            arr[k] = Ref(2, d)
        elif d == "first":
            arr[k] = Ref(0, 0.0)
        elif d == "last":
            arr[k] = Ref(1, 1.0)
        elif d == "middle":
            arr[k] = Ref(2, 0.5)
        elif d == "before":
            arr[k] = Ref(2, -1.0)
        elif d == "after":
            arr[k] = Ref(2, 2.0)
        else:
            raise ValueError(f"direction argument should be first, middle, last. Or float. {d}")
    return arr
