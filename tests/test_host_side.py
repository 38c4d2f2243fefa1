"""CPU-only checks: the C-ABI library builds, loads and exports every symbol include/*.h declares; host-side
logic of the drop-in classes (registries, error conventions, shard arithmetic); and the multi-GPU composition
(contiguous shards + GLOBAL time range + sum all-reduce) on a world_size-2 gloo group, with the oracle standing in
for the kernels (it is the checker here, never the product path)."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_loads_and_exports_header_symbols():
    from event_based_optical_flow_b200 import _build, _lib
    path = _build.build_library()
    assert os.path.exists(path)
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "cmax_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(cmax_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    raw = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/cmax_b200.h but not exported"
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.cmax_abi_version() == _lib.ABI_VERSION == 2
    assert lib.cmax_build_arch() == b"sm_100a"


def test_library_is_sm100a_only():
    import shutil
    import subprocess
    from event_based_optical_flow_b200 import _build
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _build.build_library()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_errors_need_no_gpu():
    """Argument validation happens before any CUDA call, so it is testable here."""
    from event_based_optical_flow_b200 import _lib
    lib = _lib.load()
    assert lib.cmax_time_range(None, 10, 2, None, None) == _lib.ERR_ARG
    assert b"d_minmax" in lib.cmax_last_error()
    assert lib.cmax_vote(None, 5, 4, None, 8, 8, 0, 0, 7, None, None) == _lib.ERR_ARG
    assert lib.cmax_blur3(None, None, 1, 8, 8, 1.0, 0, None) == _lib.ERR_ARG
    assert lib.cmax_objective_workspace_bytes(None, None) == 0
    with pytest.raises(ValueError):
        _lib.check("cmax_vote", _lib.ERR_ARG)
    with pytest.raises(IndexError):
        _lib.check("cmax_plan_create", _lib.ERR_SOURCE_OOB)
    with pytest.raises(ValueError):
        _lib.refs_array(("sideways",))
    with pytest.raises(ValueError):
        _lib.refs_array(())


def test_no_cpu_fallback():
    import event_based_optical_flow_b200 as B
    ev = torch.zeros(4, 4)
    with pytest.raises(RuntimeError):
        B.ContrastObjective(ev, (8, 8))
    with pytest.raises(RuntimeError):
        B.Warp((8, 8)).warp_event(ev, torch.zeros(2, 8, 8), "dense-flow")
    with pytest.raises(RuntimeError):
        B.EventImageConverter((8, 8)).create_iwe(ev)
    with pytest.raises(RuntimeError):
        B.cost_functions["image_variance"]().calculate({"iwe": torch.zeros(8, 8), "omit_boundary": True})
    if not torch.cuda.is_available():
        from event_based_optical_flow_b200.patch_init import PatchCandidateEvaluator
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            PatchCandidateEvaluator(ev, [(0, 8, 0, 8)], (8, 8))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "event_based_optical_flow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cm_oracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def test_cost_registry_and_host_logic():
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.costs import HybridCost
    assert set(B.cost_functions) == {"image_variance", "gradient_magnitude", "normalized_image_variance",
                                     "normalized_gradient_magnitude", "multi_focal_normalized_image_variance",
                                     "multi_focal_normalized_gradient_magnitude", "total_variation"}
    assert set(B.COST_TABLE) == set(B.cost_functions) - {"total_variation"}
    with pytest.raises(ValueError):
        B.cost_functions["gradient_magnitude"](direction="up")
    h = HybridCost("minimize", {"multi_focal_normalized_gradient_magnitude": 1.0, "total_variation": 0.01}, store_history=True)
    assert set(h.required_keys) == {"forward_iwe", "backward_iwe", "middle_iwe", "omit_boundary", "orig_iwe", "flow"}
    assert set(h.get_history()) == {"loss", "multi_focal_normalized_gradient_magnitude", "total_variation"}
    w = B.Warp((4, 6), normalize_t=True)
    assert w.get_key_names("2d-translation") == ["trans_x", "trans_y"]
    assert w.get_motion_vector_size("rigid-optical-flow") == 2
    with pytest.raises(B.MotionModelKeyError):
        w.get_key_names("affine")
    flow = w.get_flow_from_motion(np.array([1.5, -2.0]), "2d-translation")
    assert flow.shape == (2, 4, 6) and np.all(flow[0] == -1.5) and np.all(flow[1] == 2.0)
    im = B.EventImageConverter((4, 6), outer_padding=2)
    assert im.image_size == (8, 10) and im.outer_padding == (2, 2)
    # total variation stays in torch (tiny patch grid): check against the oracle-free closed form on a ramp
    tv = B.cost_functions["total_variation"]()
    ramp = torch.arange(5.0)[None, :, None].expand(2, 5, 5).contiguous()
    assert abs(float(tv.calculate({"flow": ramp, "omit_boundary": True})) - 0.5) < 1e-6  # d/dx = 1 on 2 of 4 channels


def test_shard_bounds_partition():
    from event_based_optical_flow_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from event_based_optical_flow_b200.distributed import global_time_range, shard_events
        from oracle import cm_oracle as O
        rng = np.random.default_rng(5)
        H, W, n = 24, 32, 4001
        ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
        ev = torch.from_numpy(ev.astype(np.float32))
        flow = torch.from_numpy(rng.uniform(-4, 4, (2, H, W)).astype(np.float32))
        mine = shard_events(ev, world, rank)
        tmin, tmax = global_time_range(mine)
        assert tmin == float(ev[:, 2].min()) and tmax == float(ev[:, 2].max())
        # partial IWE of this shard with the GLOBAL reference time / period, then the sum all-reduce
        t = mine[:, 2]
        dt = (t - tmin) / torch.tensor(tmax - tmin, dtype=torch.float32)
        src = mine[:, 0].long() * W + mine[:, 1].long()
        warped = mine.clone()
        warped[:, 0] = mine[:, 0] - dt * flow[0].reshape(-1)[src]
        warped[:, 1] = mine[:, 1] - dt * flow[1].reshape(-1)[src]
        part = O.bilinear_vote(warped, (H, W))
        dist.all_reduce(part)
        full = O.bilinear_vote(O.warp_dense(ev, flow, "first"), (H, W))
        torch.testing.assert_close(part, full, rtol=1e-5, atol=1e-5)
        # every rank computes the identical cost from the identical reduced image
        cost = -O.image_variance(part)
        gathered = [torch.zeros_like(cost) for _ in range(world)]
        dist.all_gather(gathered, cost)
        assert all(torch.equal(gathered[0], c) for c in gathered)
        out[rank] = float(cost)
    finally:
        dist.destroy_process_group()


def test_sharded_composition_gloo_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_gloo_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world and out[0] == out[1]


def _dual_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from event_based_optical_flow_b200.ops import SumPartials, UseReplicated
        torch.manual_seed(0)
        A = torch.randn(world, 7, 5, dtype=torch.float64)  # every rank draws the same numbers; rank r USES A[r]
        m0 = torch.randn(5, dtype=torch.float64)
        v = torch.randn(5, dtype=torch.float64)

        def part(m, r):  # a nonlinear 'partial image' of rank r
            return torch.sin(A[r] @ m) + (A[r] @ m) ** 2

        def cost_of_image(img):
            return (img ** 3).sum() + img.var()

        def sharded(m):  # what every rank records: its own partial, summed over the ranks
            return cost_of_image(SumPartials.apply(part(UseReplicated.apply(m, None), rank), None))

        def whole(m):
            return cost_of_image(sum(part(m, r) for r in range(world)))

        m = m0.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(sharded(m), m, create_graph=True)
        (hv,) = torch.autograd.grad(g, m, v)
        mr = m0.clone().requires_grad_(True)
        (gr,) = torch.autograd.grad(whole(mr), mr, create_graph=True)
        (hvr,) = torch.autograd.grad(gr, mr, v)
        torch.testing.assert_close(g.detach(), gr.detach(), rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(hv, hvr, rtol=1e-12, atol=1e-12)
        # the path scipy_autograd takes (torch.autograd.functional.vhp), and a third derivative for good measure
        _, hv2 = torch.autograd.functional.vhp(sharded, m0, v)
        torch.testing.assert_close(hv2, hvr, rtol=1e-12, atol=1e-12)
        m = m0.clone().requires_grad_(True)
        (g,) = torch.autograd.grad(sharded(m), m, create_graph=True)
        (h1,) = torch.autograd.grad(g, m, v, create_graph=True)
        (t3,) = torch.autograd.grad(h1, m, v)
        mr = m0.clone().requires_grad_(True)
        (gr,) = torch.autograd.grad(whole(mr), mr, create_graph=True)
        (h1r,) = torch.autograd.grad(gr, mr, v, create_graph=True)
        (t3r,) = torch.autograd.grad(h1r, mr, v)
        torch.testing.assert_close(t3, t3r, rtol=1e-11, atol=1e-11)
        out[rank] = hv.tolist()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", (2, 3))
def test_sharded_second_order_autograd_pair_gloo(world):
    """`ops.SumPartials` / `ops.UseReplicated` (each the other's backward): gradient, Hessian-vector product and a third
    derivative recorded rank-locally on a sharded composition equal the single-process ones, and every rank holds the same."""
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_dual_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world and all(out[r] == out[0] for r in range(world))


def test_cost_plugin_contract_on_the_host():
    """The plugin contract the reference's solvers rely on (src/costs/base.py:11-77, src/costs/hybrid.py:12-79), as far as it
    needs no GPU: KeyError for a missing arg-dict key, history bookkeeping, hybrid weights / 'inv' / member histories."""
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.costs import CostBase, HybridCost
    for name, cls in B.cost_functions.items():
        assert issubclass(cls, CostBase) and cls.name == name
        with pytest.raises(KeyError):
            cls().calculate({})
    assert "hybrid" not in B.cost_functions  # like the reference's table, which is built before HybridCost is imported
    with pytest.raises(NotImplementedError):
        B.cost_functions["image_variance"]().calculate({"iwe": np.zeros((8, 8)), "omit_boundary": True})  # numpy: not a CUDA plugin's job
    flow = torch.arange(5.0)[None, :, None].expand(2, 5, 5).contiguous()
    tv = B.cost_functions["total_variation"](direction="minimize", store_history=True)
    a = float(tv.calculate({"flow": flow, "omit_boundary": True}))
    tv.disable_history_register()
    tv.calculate({"flow": flow, "omit_boundary": True})
    tv.enable_history_register()
    tv.calculate({"flow": 2 * flow, "omit_boundary": True})
    hist = tv.get_history()["loss"]
    assert len(hist) == 2 and abs(hist[0] - a) < 1e-12 and abs(hist[1] - 2 * a) < 1e-6
    tv.clear_history()
    assert tv.get_history() == {"loss": []}
    assert float(B.cost_functions["total_variation"](direction="maximize").calculate({"flow": flow, "omit_boundary": True})) == -a
    h = HybridCost("minimize", {"total_variation": 3.0}, store_history=True)
    assert abs(float(h.calculate({"flow": flow, "omit_boundary": True})) - 3.0 * a) < 1e-6
    h.update_weight({"total_variation": "inv"})
    assert abs(float(h.calculate({"flow": flow, "omit_boundary": True})) - 1.0 / a) < 1e-5
    assert [len(v) for v in h.get_history().values()] == [2, 2]
    with pytest.raises(AssertionError):
        h.update_weight({"image_variance": 1.0})
    h.disable_history_register()
    assert h.cost_func["total_variation"]["func"].store_history is False


# ------------------------------------------------------------------------------------------------ total variation vs golden
def test_total_variation_matches_reference_golden():
    """`TotalVariation` (torch, on the <=16x16 patch grid) against outputs of the reference's own class
    (tests/golden/make_golden_tv.py): value and autograd gradient, both directions, both precisions, with and without the
    boundary ring, and the batched form."""
    import os
    from conftest import GOLDEN_DIR
    from event_based_optical_flow_b200.costs import functions
    g = np.load(os.path.join(GOLDEN_DIR, "reference_tv.npz"))
    for k, (h, w) in enumerate(g["shapes"]):
        flow = g[f"{k}/flow"]
        for direction in ("minimize", "maximize"):
            for omit in (True, False):
                for prec, dt, tol in (("32", torch.float32, 1e-6), ("64", torch.float64, 1e-13)):
                    cost = functions["total_variation"](direction=direction, precision=prec)
                    f = torch.from_numpy(flow).to(dt).requires_grad_(True)
                    val = cost.calculate({"flow": f, "omit_boundary": omit})
                    (grad,) = torch.autograd.grad(val, f)
                    key = f"{k}/{direction}/{int(omit)}/{prec}"
                    np.testing.assert_allclose(val.detach().numpy(), g[key + "/value"], rtol=tol, atol=tol, err_msg=key)
                    np.testing.assert_allclose(grad.numpy(), g[key + "/grad"], rtol=tol, atol=tol, err_msg=key)
        fb = torch.from_numpy(np.stack([flow, -0.5 * flow])).double()
        val = functions["total_variation"](direction="minimize", precision="64").calculate({"flow": fb, "omit_boundary": True})
        np.testing.assert_allclose(val.numpy(), g[f"{k}/batched"], rtol=1e-13, atol=1e-13)


# ------------------------------------------------------------------------------------------------ dual operators
def test_dual_operator_dispatches_numpy_to_the_reference_object():
    """`use_b200_operators` must leave the reference's run loop intact: numpy / CPU input goes to the solver's original
    objects, attributes come from them, and the history stays one dict."""
    from event_based_optical_flow_b200.solver import _DualOperator, _has_cuda_tensor

    class Ref:
        image_size = (4, 5)
        direction = "minimize"
        store_history = True

        def __init__(self):
            self.history = {"loss": []}
            self.calls = []

        def create_iwe(self, events, method="bilinear_vote", sigma=1):
            self.calls.append(("ref", type(events).__name__))
            return "ref-result"

        def clear_history(self):
            self.history = {"loss": []}

    class Fast:
        store_history = True

        def __init__(self):
            self.history = {"loss": []}

        def create_iwe(self, events, method="bilinear_vote", sigma=1):
            raise AssertionError("the CUDA operator must not see numpy input")

    ref, fast = Ref(), Fast()
    dual = _DualOperator(fast, ref)
    assert dual.image_size == (4, 5) and dual.direction == "minimize"
    assert dual.create_iwe(np.zeros((3, 4))) == "ref-result"
    assert dual.create_iwe(torch.zeros(3, 4), sigma=0) == "ref-result"  # a CPU tensor is the reference's business too
    assert ref.calls == [("ref", "ndarray"), ("ref", "Tensor")]
    assert fast.history is ref.history
    dual.store_history = False
    assert ref.store_history is False and fast.store_history is False
    dual.clear_history()
    assert fast.history is ref.history
    assert not _has_cuda_tensor({"a": [np.zeros(2), torch.zeros(2)]})


# ------------------------------------------------------------------------------------------------ pixel re-sharding (gloo)
def _reshard_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from event_based_optical_flow_b200.distributed import pixel_sort_key, reshard_events_by_pixel, shard_events
        rng = np.random.default_rng(9)
        H, W, n = 70, 100, 9001
        ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
        ev = torch.from_numpy(ev.astype(np.float32))
        mine = shard_events(ev, world, rank)               # what every rank is handed: a contiguous time slice
        got = reshard_events_by_pixel(mine, (H, W))
        keys = pixel_sort_key(got, (H, W))
        out[rank] = (got.numpy(), int(keys.min()), int(keys.max()))
        # events of one pixel keep their time order
        order = torch.sort(keys, stable=True).indices
        k_sorted, t_sorted = keys[order], got[order, 2]
        same = k_sorted[1:] == k_sorted[:-1]
        assert bool((t_sorted[1:][same] >= t_sorted[:-1][same]).all())
    finally:
        dist.destroy_process_group()


def test_reshard_events_by_pixel_gloo_world3():
    """One all-to-all turns time slices into contiguous slices of the pixel-ordered stream: nothing lost or duplicated, the
    key ranges of the ranks are disjoint and ordered, the counts are balanced, time order inside a pixel is kept."""
    world = 3
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_reshard_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        parts = [out[r] for r in range(world)]
    rng = np.random.default_rng(9)
    H, W, n = 70, 100, 9001
    ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1).astype(np.float32)
    allev = np.concatenate([p[0] for p in parts])
    assert allev.shape == ev.shape
    canon = lambda a: a[np.lexsort((a[:, 3], a[:, 2], a[:, 1], a[:, 0]))]  # noqa: E731
    np.testing.assert_array_equal(canon(allev), canon(ev))
    for r in range(world - 1):
        assert parts[r][2] < parts[r + 1][1], "key ranges must be disjoint and in rank order"
    counts = [len(p[0]) for p in parts]
    assert max(counts) - min(counts) <= 0.02 * n + 8, counts


def test_mixin_defers_cost_history_until_it_is_read():
    """SURVEY.md section 8f row 4 (drop the per-call `.item()`): losses registered through the mixin stay tensors until
    `get_history()` / `clear_history()` of the cost object is called; order and values are those of immediate registration."""
    from event_based_optical_flow_b200.costs import CostBase, HybridCost
    from event_based_optical_flow_b200.solver import B200CostMixin

    class Plain(CostBase):
        name = "plain_for_test"

        def _loss(self, arg):
            return arg["x"]

    mixin = B200CostMixin()
    a, b = Plain(store_history=True), Plain(store_history=True)
    hybrid = HybridCost("minimize", {"total_variation": 1.0}, store_history=True)
    for k in range(3):
        mixin._b200_register(a, torch.tensor(float(k)))
        mixin._b200_register(b, torch.tensor(10.0 + k, dtype=torch.float64))
        mixin._b200_register(hybrid, torch.tensor(100.0 + k))
    assert a.history["loss"] == [] and len(mixin._b200_pending) == 9
    assert a.get_history()["loss"] == [0.0, 1.0, 2.0]          # reading one flushes all (one copy)
    assert b.history["loss"] == [10.0, 11.0, 12.0] and mixin._b200_pending == []
    assert hybrid.get_history()["loss"] == [100.0, 101.0, 102.0]
    mixin._b200_register(a, torch.tensor(7.0))
    a.clear_history()                                           # pending entries belong to the history that is being cleared
    assert a.get_history()["loss"] == [] and mixin._b200_pending == []
    a.disable_history_register()
    mixin._b200_register(a, torch.tensor(8.0))
    assert a.get_history()["loss"] == []
    # a plugin that stays in torch: its own bookkeeping is bypassed for the call and deferred as well
    flow = torch.arange(5.0)[None, :, None].expand(2, 5, 5).contiguous()
    tv = hybrid.cost_func["total_variation"]["func"]
    loss = mixin._b200_unrecorded(tv, {"flow": flow, "omit_boundary": True})
    assert tv.store_history and tv.history["loss"] == [] and tv.get_history()["loss"] == [float(loss)]
    # switched off: immediate, like the reference
    eager = B200CostMixin()
    eager.b200_defer_history = False
    eager._b200_register(b, torch.tensor(1.5))
    assert b.history["loss"][-1] == 1.5


def test_total_variation_single_function_second_order_is_zero_like_autograd():
    """The one-pass TV (value + analytic gradient in the forward) under a recorded backward: connected to the flow with a zero
    Hessian, exactly what autograd through conv -> abs -> mean gives (torch.autograd.functional.vhp is what Newton-CG calls)."""
    from event_based_optical_flow_b200.costs import functions
    from event_based_optical_flow_b200.costs.total_variation import total_variation_loss
    tv = functions["total_variation"](direction="minimize")
    torch.manual_seed(1)
    m0 = torch.randn(2, 9, 13, dtype=torch.float64)
    v = torch.randn_like(m0)
    val, hv = torch.autograd.functional.vhp(lambda x: tv.calculate({"flow": x, "omit_boundary": True}), m0, v)
    assert float(hv.abs().max()) == 0.0
    # inside a sum with a curved term the Hessian is the curved term's alone
    _, hv2 = torch.autograd.functional.vhp(lambda x: total_variation_loss(x, True) * 3.0 + (x ** 3).sum(), m0, v)
    torch.testing.assert_close(hv2, 6 * m0 * v)
    assert float(total_variation_loss(m0, True, "maximize")) == -float(val)
    assert not total_variation_loss(m0, True).requires_grad  # no graph, no saved gradient, when nothing asks for one


def test_binding_constants_equal_the_header():
    """Every enum value / limit `_lib.py` hard-codes is the one include/cmax_b200.h declares (the binding a maintainer would
    write from the header alone must agree with the one shipped)."""
    from event_based_optical_flow_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cmax_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    values = {k: int(v) for k, v in re.findall(r"\b(CMAX_[A-Z0-9_]+)\s*=\s*(\d+)", header)}
    values.update({k: int(v) for k, v in re.findall(r"#define\s+(CMAX_[A-Z0-9_]+)\s+(\d+)", header)})
    assert values["CMAX_ABI_VERSION"] == _lib.ABI_VERSION
    assert (values["CMAX_MAX_REFS"], values["CMAX_MAX_BINS"], values["CMAX_MAX_PEERS"]) == (_lib.MAX_REFS, _lib.MAX_BINS, _lib.MAX_PEERS)
    assert [values[k] for k in ("CMAX_OK", "CMAX_ERR_ARG", "CMAX_ERR_CUDA", "CMAX_ERR_SOURCE_OOB", "CMAX_ERR_WORKSPACE")] == \
        [_lib.OK, _lib.ERR_ARG, _lib.ERR_CUDA, _lib.ERR_SOURCE_OOB, _lib.ERR_WORKSPACE]
    assert _lib.MOTION == {"dense-flow": values["CMAX_MOTION_DENSE"], "dense-flow-voxel": values["CMAX_MOTION_VOXEL"],
                           "2d-translation": values["CMAX_MOTION_2DOF"], "rigid-optical-flow": values["CMAX_MOTION_2DOF"],
                           "tile-flow": values["CMAX_MOTION_TILE"]}
    assert _lib.STAT == {"variance": values["CMAX_STAT_VARIANCE"], "gradmag": values["CMAX_STAT_GRADMAG"]}
    assert _lib.FORM == {"plain": values["CMAX_COST_PLAIN"], "normalized": values["CMAX_COST_NORMALIZED"], "multifocal": values["CMAX_COST_MULTIFOCAL"]}
    assert _lib.ORDER == {"asis": values["CMAX_ORDER_ASIS"], "tile": values["CMAX_ORDER_TILE"], "pixel": values["CMAX_ORDER_PIXEL"]}
    assert _lib.SCHEME == {"upwind": values["CMAX_SCHEME_UPWIND"], "burgers": values["CMAX_SCHEME_BURGERS"]}
    assert (_lib.PATCH_GLOBAL_IMAGES, _lib.PATCH_KEEP_IMAGES) == (values["CMAX_PATCH_GLOBAL_IMAGES"], values["CMAX_PATCH_KEEP_IMAGES"])


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/cmax_b200.h compiles as C (gcc, no CUDA headers: what a cgo / JNI / ctypes binding sees), and the structs the
    Python binding mirrors have the C compiler's sizes and field offsets."""
    import shutil
    import subprocess
    from event_based_optical_flow_b200 import _lib
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "layout.c"
    src.write_text(r"""
#include <stddef.h>
#include <stdio.h>
#include "cmax_b200.h"
int main(void) {
  printf("ref %zu %zu %zu\n", sizeof(cmax_ref), offsetof(cmax_ref, mode), offsetof(cmax_ref, fraction));
  printf("spec %zu %zu %zu %zu %zu %zu %zu\n", sizeof(cmax_cost_spec), offsetof(cmax_cost_spec, stat), offsetof(cmax_cost_spec, form),
         offsetof(cmax_cost_spec, direction_sign), offsetof(cmax_cost_spec, omit_boundary), offsetof(cmax_cost_spec, sigma),
         offsetof(cmax_cost_spec, weights));
  printf("peers %zu %zu %zu %zu %zu %zu\n", sizeof(cmax_peers), offsetof(cmax_peers, n_peers), offsetof(cmax_peers, rank),
         offsetof(cmax_peers, iwe), offsetof(cmax_peers, grad), offsetof(cmax_peers, flags));
  printf("time %zu\n", sizeof(cmax_time_params_t));
  return 0;
}
""")
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict((line.split()[0], [int(v) for v in line.split()[1:]]) for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())

    def layout(struct):
        return [ctypes.sizeof(struct)] + [getattr(struct, name).offset for name, _ in struct._fields_]

    assert out["ref"] == layout(_lib.Ref)
    assert out["spec"] == layout(_lib.CostSpec)
    assert out["peers"] == layout(_lib.Peers)
    assert out["time"] == [_lib.TIME_PARAMS_BYTES]


def test_every_entry_point_rejects_null_arguments_without_a_gpu():
    """Argument validation comes before any CUDA call: with NULL pointers and zero sizes every status-returning entry point of
    the C ABI answers CMAX_ERR_ARG and names itself in cmax_last_error() -- on a machine without a GPU, without crashing."""
    from event_based_optical_flow_b200 import _lib
    lib = _lib.load()
    scalars = {ctypes.c_int: 0, ctypes.c_int32: 0, ctypes.c_int64: 0, ctypes.c_size_t: 0, ctypes.c_float: 0.0, ctypes.c_double: 0.0}
    checked = 0
    for name, (restype, argtypes) in _lib._SIGNATURES.items():
        if restype is not ctypes.c_int or not argtypes:
            continue
        args = [scalars.get(t) for t in argtypes]  # None = NULL for every pointer type
        status = getattr(lib, name)(*args)
        message = lib.cmax_last_error().decode()
        assert status == _lib.ERR_ARG, (name, status, message)
        assert message.startswith(name), (name, message)
        checked += 1
    assert checked >= 25
    with pytest.raises(ValueError, match="cmax_vote"):  # the mapping of the status to the reference's exception type
        _lib.call("cmax_vote", *[scalars.get(t) for t in _lib._SIGNATURES["cmax_vote"][1]])
