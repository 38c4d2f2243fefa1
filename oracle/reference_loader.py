"""The UNMODIFIED reference (tub-rip/event_based_optical_flow) as a checker / CPU baseline -- TEST INFRASTRUCTURE ONLY.

`install()` copies the reference's pure-Python sources (src/, configs/) from /root/reference into `baseline/_ref/` (git-ignored:
the repository's history stays free of reference code; the directory travels to the GPU box with the gpurun payload, where
/root/reference does not exist).  `load()` imports it from there -- or straight from /root/reference when present -- after
putting import-only stand-ins into `sys.modules` for the third-party packages this image lacks and the hot path never calls
(optuna, plotly, matplotlib, skimage, h5py, hdf5plugin, torch_scatter).  Only tests/, __graft_entry__ and bench.py's
`--impl reference` / cpu_baseline legs may use this module; nothing under event_based_optical_flow_b200/ imports it.
"""
from __future__ import annotations

import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_SRC = os.environ.get("CM_REFERENCE", "/root/reference")
INSTALL_DIR = os.path.join(ROOT, "baseline", "_ref")
_STUBBED = ("optuna", "optuna.storages", "optuna.distributions", "optuna.samplers", "optuna.study", "optuna.logging", "matplotlib",
            "matplotlib.pyplot", "plotly", "plotly.graph_objects", "skimage", "skimage.transform", "h5py", "hdf5plugin", "torch_scatter")


class _Anything:
    """Stands in for any attribute / class / call result of a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return sys.modules.get(self.__name__ + "." + name, _Anything)


def install(force: bool = False) -> str | None:
    """Copy the reference's Python sources into baseline/_ref (no-op without /root/reference).  Returns the directory."""
    if not os.path.isdir(os.path.join(REFERENCE_SRC, "src")):
        return INSTALL_DIR if os.path.isdir(os.path.join(INSTALL_DIR, "src")) else None
    if os.path.isdir(os.path.join(INSTALL_DIR, "src")) and not force:
        return INSTALL_DIR
    os.makedirs(INSTALL_DIR, exist_ok=True)
    for sub in ("src", "configs"):
        dst = os.path.join(INSTALL_DIR, sub)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REFERENCE_SRC, sub), dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    return INSTALL_DIR


def location() -> str | None:
    for cand in (INSTALL_DIR, REFERENCE_SRC):
        if os.path.isdir(os.path.join(cand, "src")):
            return cand
    return None


def load():
    """-> namespace(solver, costs, warp, event_image_converter, utils, root) of the unmodified reference, or None if it is
    neither installed under baseline/_ref nor mounted at /root/reference."""
    root = location()
    if root is None:
        return None
    def missing(top: str) -> bool:
        if top in sys.modules:
            return isinstance(sys.modules[top], _StubModule)
        try:
            return importlib.util.find_spec(top) is None
        except (ValueError, ImportError):
            return True

    for name in _STUBBED:
        if name not in sys.modules and missing(name.split(".")[0]):
            sys.modules[name] = _StubModule(name)
    for name in _STUBBED:  # parents expose their stubbed children as attributes
        if "." in name and isinstance(sys.modules.get(name), _StubModule):
            parent, child = name.rsplit(".", 1)
            if isinstance(sys.modules.get(parent), _StubModule):
                setattr(sys.modules[parent], child, sys.modules[name])
    if root not in sys.path:
        sys.path.insert(0, root)
    from src import costs, event_image_converter, solver, utils, warp  # noqa: E402
    return types.SimpleNamespace(solver=solver, costs=costs, warp=warp, event_image_converter=event_image_converter, utils=utils, root=root)
