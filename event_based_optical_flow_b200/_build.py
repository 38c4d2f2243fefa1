"""In-tree nvcc build of libcmax_b200.so (sm_100a only).

The library is the product: there is no CPU or torch fallback, so a missing toolchain or a failed build raises.
-fmad=false keeps ptxas from contracting `x - dt*f` into an FMA, which would break bit-exact warped coordinates
(hard part 1); the event math additionally uses explicit __f*_rn intrinsics.
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcmax_b200.so")
SOURCES = ("cmax_events.cu", "cmax_ops.cu", "cmax_cost.cu", "cmax_fused.cu", "cmax_mid.cu", "cmax_tileflow.cu", "cmax_flowvoxel.cu", "cmax_lean.cu", "cmax_patchinit.cu")
HEADERS = ("cmax_common.cuh", "cmax_plan.cuh", "cmax_stats.cuh", "cmax_runs.cuh", "cmax_objective.cuh", "cmax_tile.cuh", os.path.join("..", "..", "include", "cmax_b200.h"))
# `--measure` builds a SECOND library, lib/libcmax_b200_measure.so, with the measurement aids compiled in (-DCMAX_MEASURE:
# partial stage masks, CMAX_PDL=0, phase stamps of the image / exchange kernels).  The release library has none of them and
# nothing in a process's environment selects the other file: a probe script asks for it explicitly (_lib.use_measure_library()).
NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libcmax_b200.so cannot be built (set NVCC=/path/to/nvcc)")


OBJ_DIR = os.path.join(LIB_DIR, "obj")
COMPILE_FLAGS = tuple(f for f in NVCC_FLAGS if f not in ("-shared",))


def _header_mtime() -> float:
    return max(os.path.getmtime(os.path.join(CSRC, h)) for h in HEADERS if os.path.exists(os.path.join(CSRC, h)))


def _obj_path(src: str) -> str:
    return os.path.join(OBJ_DIR, src.replace(".cu", ".o"))


def _obj_stale(src: str) -> bool:
    obj = _obj_path(src)
    if not os.path.exists(obj):
        return True
    built = os.path.getmtime(obj)
    return os.path.getmtime(os.path.join(CSRC, src)) > built or _header_mtime() > built


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.exists(d) and os.path.getmtime(d) > built for d in deps)


MEASURE_LIB_PATH = os.path.join(LIB_DIR, "libcmax_b200_measure.so")


def build_measure_library(verbose: bool = False) -> str:
    """lib/libcmax_b200_measure.so: every source recompiled with -DCMAX_MEASURE (its own object directory)."""
    obj_dir = os.path.join(LIB_DIR, "obj_measure")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        procs.append((src, obj, subprocess.Popen([nvcc, "-DCMAX_MEASURE", *COMPILE_FLAGS, "-c", "-o", obj, os.path.join(CSRC, src)],
                                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for src, _, proc in procs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({proc.returncode}):\n{out}\n{err}")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", MEASURE_LIB_PATH, *[o for _, o, _ in procs]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    return MEASURE_LIB_PATH


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into lib/libcmax_b200.so; returns the path.  No-op when up to date.  Every source is its own
    translation unit (no relocatable device code), so stale objects are recompiled in parallel and then linked."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    todo = [s for s in SOURCES if force or _obj_stale(s)]
    procs = []
    for src in todo:
        cmd = [nvcc, *COMPILE_FLAGS, "-c", "-o", _obj_path(src), os.path.join(CSRC, src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for src, proc in procs:
        out, err = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} ({proc.returncode}):\n{out}\n{err}")
        if verbose:
            print(err)
    tmp = LIB_PATH + ".tmp"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared", "-o", tmp, *[_obj_path(s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"nvcc link failed ({proc.returncode}):\n{proc.stdout}\n{proc.stderr}")
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    if "--measure" in sys.argv:
        print(build_measure_library(verbose="-v" in sys.argv))
    else:
        print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
