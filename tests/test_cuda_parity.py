"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Tolerances (BASELINE.json north_star): warped coordinates and floor indices BIT-EXACT vs the fp32 reference branch;
IWE, cost and gradient within 1e-5 relative (fp32 summation order differs from the reference's sequential
scatter_add_, so these cannot be bit-exact).  Gradients are compared with the SAME-dtype (fp32) oracle, norm-wise:
the objective's gradient is discontinuous at pixel borders, so fp32 and fp64 disagree wherever a floor index flips
(hard part 2).
"""
import numpy as np
import pytest
import torch

from oracle import cm_oracle as O

pytestmark = pytest.mark.gpu

CASES = ("small", "frac", "tiny")
DIRS = (("first", "first"), ("middle", "middle"), ("last", "last"), ("d0.3", 0.3))
MODELS = ("dense-flow", "dense-flow-voxel", "2d-translation")
RTOL = 1e-5


@pytest.fixture(scope="module")
def B():
    import event_based_optical_flow_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _inputs(g, case):
    H, W, n, T = (int(v) for v in g[f"{case}/meta"])
    t = lambda k: torch.from_numpy(g[f"{case}/{k}"]).float()
    return H, W, t("events"), {"dense-flow": t("flow"), "dense-flow-voxel": t("voxel"), "2d-translation": t("theta")}


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


def _cost_close(val, ref32, ref64=None, tol=RTOL):
    """cost within `tol` of the fp64 reference (the north-star bound); against the fp32 reference the bound is widened by
    exactly how far the fp32 reference itself sits from fp64 (its sequential fp32 scatter_add_ / fp32 statistics drift by
    up to ~1e-5 on these batches, while the CUDA path accumulates its statistics in fp64)."""
    val, ref32 = float(val), float(ref32)
    if ref64 is None:
        return abs(val - ref32) <= tol * abs(ref32)
    ref64 = float(ref64)
    return abs(val - ref64) <= tol * abs(ref64) and abs(val - ref32) <= tol * abs(ref32) + abs(ref32 - ref64)


def _oracle_warp(model, ev, motion, d):
    return {"dense-flow": O.warp_dense, "dense-flow-voxel": O.warp_voxel, "2d-translation": O.warp_2dof}[model](ev, motion, d)


# ------------------------------------------------------------------------------------------------ modular operators
@pytest.mark.parametrize("case", CASES)
def test_warp_bit_exact_vs_reference_golden(B, dev, golden_random, case):
    g = golden_random
    H, W, ev, motions = _inputs(g, case)
    warper = B.Warp((H, W), normalize_t=True)
    for model in MODELS:
        for dtag, d in DIRS:
            got, feat = warper.warp_event(ev.to(dev), motions[model].to(dev), model, direction=d)
            np.testing.assert_array_equal(got.cpu().numpy(), g[f"{case}/f32/warp/{model}/{dtag}"], err_msg=f"{model} {dtag}")
            assert feat == {"none": None}


@pytest.mark.parametrize("case", CASES)
def test_iwe_vs_reference_golden(B, dev, golden_random, case):
    g = golden_random
    H, W, ev, motions = _inputs(g, case)
    for pad in (0, 3):
        imager = B.EventImageConverter((H, W), outer_padding=pad)
        for model in MODELS:
            for dtag, d in DIRS:
                w = torch.from_numpy(g[f"{case}/f32/warp/{model}/{dtag}"]).to(dev)
                for s in (0, 1):
                    got = imager.create_iwe(w, "bilinear_vote", s).cpu().numpy()
                    ref = g[f"{case}/f32/iwe/p{pad}/{model}/{dtag}/s{s}"]
                    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-5 * max(1.0, np.abs(ref).max()))
        wgt = torch.from_numpy(g[f"{case}/weight"]).float().to(dev)
        got = imager.bilinear_vote_tensor(ev.to(dev), weight=wgt).cpu().numpy()
        ref = g[f"{case}/f32/vote_weighted/p{pad}"]
        np.testing.assert_allclose(got, ref, rtol=RTOL, atol=1e-5 * np.abs(ref).max())
        got = imager.count_event_tensor(ev.to(dev)).cpu().numpy()
        np.testing.assert_array_equal(got, g[f"{case}/f32/count/p{pad}"])  # integer counts: exact


def test_reference_hand_vectors(B, dev, golden_hand):
    """tests/test_warp.py:96-139 and tests/test_event_image_converter.py:17-69 of the reference."""
    g = golden_hand
    ev = torch.from_numpy(g["warp34/events"]).float()
    w, _ = B.Warp((3, 4), normalize_t=True).warp_event(ev.to(dev), torch.from_numpy(g["warp34/flow"]).float().to(dev), "dense-flow")
    np.testing.assert_allclose(w.cpu().numpy()[:, :3], g["warp34/expected"], rtol=1e-6, atol=1e-6)
    imager = B.EventImageConverter((3, 4))
    for k in ("int", "frac"):
        img = imager.bilinear_vote_tensor(torch.from_numpy(g[f"vote34/{k}/events"]).float().to(dev),
                                          weight=torch.from_numpy(g[f"vote34/{k}/weight"]).float().to(dev))
        np.testing.assert_allclose(img.cpu().numpy(), g[f"vote34/{k}/expected"], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("sigma", (0, 1))
def test_modular_pipeline_autograd(B, dev, golden_random, case, sigma):
    """Warp -> EventImageConverter -> cost plugin -> torch.autograd.grad, composed the way the reference's solver
    composes them (src/solver/patch_contrast_base.py:289-352), against the reference's own autograd (golden)."""
    g = golden_random
    H, W, ev, motions = _inputs(g, case)
    warper, imager = B.Warp((H, W), normalize_t=True), B.EventImageConverter((H, W))
    evd = ev.to(dev)
    for model in MODELS:
        for cn in O.COSTS:
            cost = B.cost_functions[cn](direction="minimize", store_history=True)
            motion = motions[model].to(dev).requires_grad_(True)
            arg = {"omit_boundary": True, "clip": True}
            if "orig_iwe" in cost.required_keys:
                arg["orig_iwe"] = imager.create_iwe(evd, "bilinear_vote", sigma)
            for key, d in (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")):
                if key in cost.required_keys or (key == "backward_iwe" and "iwe" in cost.required_keys):
                    w, _ = warper.warp_event(evd, motion, model, direction=d)
                    arg[key] = imager.create_iwe(w, "bilinear_vote", sigma)
            if "backward_iwe" in arg:
                arg["iwe"] = arg["backward_iwe"]
            loss = cost.calculate(arg)
            (grad,) = torch.autograd.grad(loss, motion)
            ref_v = float(g[f"{case}/f32/cost/{model}/{cn}/s{sigma}"])
            ref_v64 = float(g[f"{case}/f64/cost/{model}/{cn}/s{sigma}"])
            ref_g = g[f"{case}/f32/grad/{model}/{cn}/s{sigma}"]
            assert _cost_close(loss, ref_v, ref_v64), (model, cn, float(loss), ref_v, ref_v64)
            assert _rel(grad.cpu().numpy(), ref_g) <= RTOL, (model, cn, _rel(grad.cpu().numpy(), ref_g))
            assert len(cost.get_history()["loss"]) == 1


# ------------------------------------------------------------------------------------------------ fused objective
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("sigma", (0, 1))
@pytest.mark.parametrize("order", ("asis", "tile", "pixel"))
def test_fused_objective_vs_reference_golden(B, dev, golden_random, case, sigma, order):
    g = golden_random
    H, W, ev, motions = _inputs(g, case)
    T = int(g[f"{case}/meta"][3])
    for model in MODELS:
        for cn in O.COSTS:
            obj = B.ContrastObjective(ev.to(dev), (H, W), cost=cn, motion_model=model, sigma=float(sigma), n_bins=T, order=order)
            if order == "tile":
                obj.plan.set_variant(0, 0)  # the per-event kernels; the other orders run the default run-walk kernels
            val, grad = obj.value_and_grad(motions[model].to(dev))
            ref_v = float(g[f"{case}/f32/cost/{model}/{cn}/s{sigma}"])
            ref_v64 = float(g[f"{case}/f64/cost/{model}/{cn}/s{sigma}"])
            ref_g = g[f"{case}/f32/grad/{model}/{cn}/s{sigma}"]
            assert _cost_close(val, ref_v, ref_v64), (model, cn, float(val), ref_v, ref_v64)
            assert _rel(grad.cpu().numpy(), ref_g) <= RTOL, (model, cn, _rel(grad.cpu().numpy(), ref_g))
            assert grad.shape == motions[model].shape
            # value-only evaluation and the autograd wrapper agree with value_and_grad
            assert abs(float(obj.value(motions[model].to(dev))) - float(val)) <= 1e-6 * abs(float(val))  # atomics: order varies
            m = motions[model].to(dev).double().requires_grad_(True)
            loss = obj(m)
            assert loss.dtype == torch.float64
            (g2,) = torch.autograd.grad(loss * 2.0, m)
            assert _rel(g2.cpu().numpy(), 2.0 * grad.double().cpu().numpy()) <= 1e-5
            assert abs(float(val) - ref_v64) <= RTOL * abs(ref_v64)


@pytest.mark.parametrize("pad", (0, 3))
def test_fused_iwe_and_indices_bit_exact(B, dev, golden_random, pad):
    """Bit-exact warped floor indices: an IWE built by `count`-style integer votes would hide fractional errors, so
    compare (a) the fused IWE with the oracle IWE built from the golden warped coordinates and (b) floor indices."""
    g = golden_random
    for case in CASES:
        H, W, ev, motions = _inputs(g, case)
        obj = B.ContrastObjective(ev.to(dev), (H, W), cost="multi_focal_normalized_image_variance", motion_model="dense-flow",
                                  outer_padding=pad, order="asis")
        iwes = obj.iwe(motions["dense-flow"].to(dev)).cpu()
        for k, d in enumerate(("first", "last", "middle")):
            w = torch.from_numpy(g[f"{case}/f32/warp/dense-flow/{d}"])
            ref = O.create_iwe(w, (H, W), (pad, pad), "bilinear_vote", 0)
            np.testing.assert_allclose(iwes[k].numpy(), ref.numpy(), rtol=RTOL, atol=1e-5 * float(ref.abs().max()))
            got_w, _ = B.Warp((H, W), normalize_t=True).warp_event(ev.to(dev), motions["dense-flow"].to(dev), "dense-flow", direction=d)
            fl = torch.floor(got_w[:, :2] + 1e-6).long().cpu()
            np.testing.assert_array_equal(fl.numpy(), torch.floor(w[:, :2] + 1e-6).long().numpy())


def test_fractional_coordinates_keep_the_16_byte_format(B, dev, golden_random):
    H, W, ev, motions = _inputs(golden_random, "frac")
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="dense-flow")
    assert obj.plan.set_compact(True) is False
    H, W, ev, motions = _inputs(golden_random, "small")
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="dense-flow")
    assert obj.plan.set_compact(True) is (bool((ev[:, :2] == ev[:, :2].floor()).all()))


@pytest.mark.parametrize("order", ("asis", "pixel"))
@pytest.mark.parametrize("fractional", (False, True))
def test_fused_path_is_bit_exact_when_votes_do_not_collide(B, dev, order, fractional):
    """One event per 4x4 block and |flow*dt| < 0.9: no two events share a pixel, so every IWE pixel is a single bilinear
    weight and must equal the reference arithmetic bit for bit (warp product rounded before the subtraction, floor(x+1e-6),
    fractions, weight products) -- in both packed-event formats and for every reference time."""
    rng = np.random.default_rng(5)
    H, W = 64, 96
    rows, cols = np.meshgrid(np.arange(1, H - 2, 4), np.arange(1, W - 2, 4), indexing="ij")
    n = rows.size
    ev = np.zeros((n, 4), dtype=np.float32)
    ev[:, 0] = rows.ravel() + (rng.uniform(0, 0.05, n) if fractional else 0)
    ev[:, 1] = cols.ravel() + (rng.uniform(0, 0.05, n) if fractional else 0)
    ev[:, 2] = np.sort(rng.uniform(0, 0.05, n))
    ev = torch.from_numpy(ev)
    flow = torch.from_numpy(rng.uniform(-0.9, 0.9, (2, H, W)).astype(np.float32))
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="multi_focal_normalized_image_variance", motion_model="dense-flow", order=order)
    assert obj.plan.set_compact(True) is (not fractional)
    iwes = obj.iwe(flow.to(dev)).cpu()
    for k, d in enumerate(("first", "last", "middle")):
        ref = O.create_iwe(O.warp_dense(ev, flow, d), (H, W))
        assert torch.equal(iwes[k], ref), (d, float((iwes[k] - ref).abs().max()))
    theta = torch.tensor([0.7, -0.6])
    obj2 = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="2d-translation", order=order)
    assert torch.equal(obj2.iwe(theta.to(dev))[0].cpu(), O.create_iwe(O.warp_2dof(ev, theta, "first"), (H, W)))


def test_c1_config(B, dev, golden_c1):
    """BASELINE config 1: 30k events, 346x260, 2-dof warp + variance; and the dense-flow metric path."""
    g = golden_c1
    ev = torch.from_numpy(g["events"]).to(dev)
    th = torch.from_numpy(g["theta"]).float().to(dev)
    obj = B.ContrastObjective(ev, (260, 346), cost="image_variance", motion_model="2d-translation")
    val, grad = obj.value_and_grad(th)
    assert abs(float(val) - float(g["c1/2dof/cost"])) <= RTOL * abs(float(g["c1/2dof/cost"]))
    np.testing.assert_allclose(grad.cpu().numpy(), g["c1/2dof/grad"], rtol=RTOL)
    flow = torch.from_numpy(g["flow"]).to(dev)
    w, _ = B.Warp((260, 346), normalize_t=True).warp_event(ev, flow, "dense-flow")
    np.testing.assert_array_equal(w[:, :2].cpu().numpy(), g["c1/dense/warped_xy"])
    for order in ("asis", "pixel"):
        obj = B.ContrastObjective(ev, (260, 346), cost="image_variance", motion_model="dense-flow", order=order)
        iwe = obj.iwe(flow)[0].cpu().numpy()
        np.testing.assert_allclose(iwe, g["c1/dense/iwe"], rtol=RTOL, atol=1e-5)
        val, grad = obj.value_and_grad(flow)
        assert abs(float(val) - float(g["c1/dense/cost"])) <= RTOL * abs(float(g["c1/dense/cost"]))
        assert _rel(grad.cpu().numpy(), g["c1/dense/grad"]) < RTOL


# ------------------------------------------------------------------------------------------------ scale + properties
def _synthetic(n, H, W, seed=0, max_flow=10.0):
    rng = np.random.default_rng(seed)
    ev = np.empty((n, 4), dtype=np.float32)
    ev[:, 0] = rng.integers(0, H, n)
    ev[:, 1] = rng.integers(0, W, n)
    ev[:, 2] = np.sort(rng.uniform(0, 0.05, n))
    ev[:, 3] = rng.integers(0, 2, n)
    flow = rng.uniform(-max_flow, max_flow, (2, H, W)).astype(np.float32)
    return torch.from_numpy(ev), torch.from_numpy(flow)


@pytest.mark.parametrize("variants", ((5, 5), (5, 2), (2, 5), (2, 2), (3, 3), (4, 4), (2, 4), (0, 0), (1, 0), (0, 1), (2, 1), (0, 2)))
def test_one_million_events_vs_oracle(B, dev, variants):
    H, W = 260, 346
    ev, flow = _synthetic(1_000_000, H, W, seed=1)
    ref_v, ref_g = O.objective_value_and_grad(ev, flow, (H, W), motion_model="dense-flow", cost="image_variance")
    ref_v64, _ = O.objective_value_and_grad(ev.double(), flow.double(), (H, W), motion_model="dense-flow", cost="image_variance")
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="dense-flow", order="pixel")
    obj.plan.set_variant(*variants)
    assert obj.plan.set_compact(True) is True  # integer pixel coordinates -> 8-byte packed events
    for compact in (True, False):              # ... and the 16-byte format must give the same answer
        assert obj.plan.set_compact(compact) is compact
        val, grad = obj.value_and_grad(flow.to(dev))
        assert abs(float(val) - float(ref_v64)) <= RTOL * abs(float(ref_v64))
        assert _cost_close(val, ref_v, ref_v64)
        assert _rel(grad.cpu().numpy(), ref_g.numpy()) <= RTOL, _rel(grad.cpu().numpy(), ref_g.numpy())


def test_full_size_properties(B, dev):
    """BASELINE config 2 size (5M events, 260x346): properties that need no oracle run."""
    H, W, n = 260, 346, 5_000_000
    ev, flow = _synthetic(n, H, W, seed=2)
    evd, flowd = ev.to(dev), flow.to(dev)
    obj = B.ContrastObjective(evd, (H, W), cost="image_variance", motion_model="dense-flow", order="pixel")
    # (1) zero flow: the IWE is the exact event histogram (integer coordinates -> weights 1,0,0,0), exact in fp32
    iwe0 = obj.iwe(torch.zeros_like(flowd))[0]
    hist = torch.bincount((evd[:, 0].long() * W + evd[:, 1].long()), minlength=H * W).reshape(H, W).float()
    assert torch.equal(iwe0, hist)
    # (2) mass: every event spreads weight 1 over its in-bounds corners -> sum(IWE) <= n, and equal when padding is huge
    iwe = obj.iwe(flowd)[0]
    assert float(iwe.double().sum()) <= n * (1 + 1e-6)
    objp = B.ContrastObjective(evd, (H, W), cost="image_variance", motion_model="dense-flow", outer_padding=12, order="tile")
    assert abs(float(objp.iwe(flowd)[0].double().sum()) - n) <= 1e-6 * n
    # (3) the event order does not change the result beyond summation order; nor does the scatter variant
    v_pix, g_pix = obj.value_and_grad(flowd)
    obj_asis = B.ContrastObjective(evd, (H, W), cost="image_variance", motion_model="dense-flow", order="asis")
    v_asis, g_asis = obj_asis.value_and_grad(flowd)
    assert abs(float(v_pix) - float(v_asis)) <= 1e-6 * abs(float(v_asis))
    assert _rel(g_pix.cpu().numpy(), g_asis.cpu().numpy()) <= 1e-5
    # (4) additivity over shards with the GLOBAL time range: IWE(all) == IWE(first half) + IWE(second half)
    tr = (float(ev[:, 2].min()), float(ev[:, 2].max()))
    halves = [B.ContrastObjective(evd[a:b], (H, W), cost="image_variance", motion_model="dense-flow", t_range=tr).iwe(flowd)[0]
              for a, b in ((0, n // 2), (n // 2, n))]
    torch.testing.assert_close(halves[0] + halves[1], iwe, rtol=1e-5, atol=1e-3)
    # (5) linearity of the gradient in dL/dIWE: variance cost 'maximize' is exactly the negated 'minimize'
    obj_max = B.ContrastObjective(evd, (H, W), cost="image_variance", motion_model="dense-flow", direction="maximize")
    v_max, g_max = obj_max.value_and_grad(flowd)
    assert abs(float(v_max) + float(v_pix)) <= 1e-6 * abs(float(v_pix))
    assert _rel(g_max.cpu().numpy(), -g_pix.cpu().numpy()) <= 1e-6


# ------------------------------------------------------------------------------------------------ edge cases / errors
def test_edge_cases(B, dev):
    H, W = 16, 20
    flow = torch.zeros(2, H, W, device=dev)
    # empty batch: zero image, zero variance, zero gradient
    obj = B.ContrastObjective(torch.zeros(0, 4, device=dev), (H, W), cost="image_variance", motion_model="dense-flow",
                              t_range=(0.0, 1.0))
    val, grad = obj.value_and_grad(flow)
    assert float(val) == 0.0 and float(grad.abs().sum()) == 0.0
    # every event warped far outside: all four corners masked, IWE stays zero
    ev = torch.tensor([[3.0, 4.0, 0.0, 1.0], [5.0, 6.0, 1.0, 0.0], [7.0, 8.0, 0.5, 1.0]], device=dev)
    big = torch.full((2, H, W), 1e4, device=dev)
    obj = B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow", order="asis")
    iwe = obj.iwe(big)[0]
    assert float(iwe.sum()) == 1.0  # the t = t_min event has dt = 0 and is not moved
    # partially out of bounds: each corner is masked on its own (src/event_image_converter.py:355-372)
    ev = torch.tensor([[H - 1.0, W - 1.0, 0.0, 1.0], [0.0, 0.0, 1.0, 1.0]], device=dev)
    f = torch.zeros(2, H, W, device=dev)
    f[:, 0, 0] = 0.5   # moves the second event to (-0.5, -0.5): only its (r+1,c+1) corner is inside
    f[:, H - 1, W - 1] = -0.5
    obj = B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow", order="asis")
    iwe = obj.iwe(f)[0].cpu()
    ref = O.create_iwe(O.warp_dense(ev.cpu(), f.cpu(), "first"), (H, W))
    torch.testing.assert_close(iwe, ref, rtol=0, atol=1e-7)
    assert float(iwe[0, 0]) == 0.25 and float(iwe[H - 1, W - 1]) == 1.0


def test_error_conventions(B, dev):
    H, W = 8, 8
    ev = torch.tensor([[1.0, 2.0, 0.0, 1.0], [9.0, 2.0, 1.0, 0.0]], device=dev)  # second event's row is outside
    with pytest.raises(IndexError):
        B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow")
    with pytest.raises(IndexError):
        B.Warp((H, W), normalize_t=True).warp_event(ev, torch.zeros(2, H, W, device=dev), "dense-flow")
    ok = ev[:1].repeat(4, 1)
    with pytest.raises(B.MotionModelKeyError):
        B.Warp((H, W)).warp_event(ok, torch.zeros(2, device=dev), "affine")
    with pytest.raises(ValueError):
        B.Warp((H, W)).warp_event(ok, torch.zeros(2, device=dev), "2d-translation", direction="sideways")
    with pytest.raises(NotImplementedError):
        B.EventImageConverter((H, W)).create_iwe(ok, method="polarity")
    with pytest.raises(ValueError):
        B.cost_functions["image_variance"](direction="sideways")
    with pytest.raises(KeyError):
        B.cost_functions["normalized_image_variance"]().calculate({"iwe": torch.zeros(H, W, device=dev), "omit_boundary": True})
    with pytest.raises(RuntimeError):
        B.ContrastObjective(ev.cpu(), (H, W))  # no CPU fallback


def test_directions_of_fused_costs(B, dev, golden_random):
    """'maximize' / 'natural' sign conventions of src/costs/*.py, fused path vs plugin path."""
    g = golden_random
    H, W, ev, motions = _inputs(g, "small")
    evd, flow = ev.to(dev), motions["dense-flow"].to(dev)
    warper, imager = B.Warp((H, W), normalize_t=True), B.EventImageConverter((H, W))
    for cn in O.COSTS:
        for direction in ("minimize", "maximize", "natural"):
            fused = float(B.ContrastObjective(evd, (H, W), cost=cn, motion_model="dense-flow", direction=direction).value(flow))
            cost = B.cost_functions[cn](direction=direction)
            arg = {"omit_boundary": True, "orig_iwe": imager.create_iwe(evd, "bilinear_vote", 0)}
            for key, d in (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")):
                arg[key] = imager.create_iwe(warper.warp_event(evd, flow, "dense-flow", direction=d)[0], "bilinear_vote", 0)
            arg["iwe"] = arg["backward_iwe"]
            plug = float(cost.calculate(arg))
            assert abs(fused - plug) <= 1e-5 * abs(plug), (cn, direction, fused, plug)


# ------------------------------------------------------------------------------------------------ strip kernels, all instantiations
@pytest.mark.parametrize("model,cost,sigma", [("dense-flow", "multi_focal_normalized_gradient_magnitude", 1.0),
                                              ("dense-flow", "gradient_magnitude", 0.0),
                                              ("dense-flow-voxel", "image_variance", 0.0),
                                              ("dense-flow-voxel", "multi_focal_normalized_image_variance", 1.0),
                                              ("2d-translation", "image_variance", 0.0),
                                              ("2d-translation", "multi_focal_normalized_gradient_magnitude", 1.0)])
def test_strip_kernels_every_model_vs_oracle_and_run_kernels(B, dev, model, cost, sigma):
    """A batch dense enough for the plan to build strips (65 events per pixel): the strip kernels (variant 5, the default
    there) for every motion model with 1 and 3 reference times against the oracle and against the run kernels (variant 2)."""
    rng = np.random.default_rng(17)
    H, W, n, T = 48, 64, 200_000, 6
    ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1).astype(np.float32)
    ev = torch.from_numpy(ev)
    if model == "dense-flow":
        motion = torch.from_numpy(rng.uniform(-6, 6, (2, H, W)).astype(np.float32))
    elif model == "dense-flow-voxel":
        motion = torch.from_numpy(rng.uniform(-6, 6, (T, 2, H, W)).astype(np.float32))
    else:
        motion = torch.tensor([9.0, -5.0])
    kw = dict(cost=cost, motion_model=model, sigma=sigma, n_bins=T if model == "dense-flow-voxel" else None)
    obj = B.ContrastObjective(ev.to(dev), (H, W), order="pixel", **kw)
    assert obj.plan.n_strips > 0, "this batch must qualify for strips"
    assert obj.plan.n_strips * 8 <= 1.5 * n
    v5, g5 = obj.value_and_grad(motion.to(dev))
    obj.plan.set_variant(2, 2)
    v2, g2 = obj.value_and_grad(motion.to(dev))
    ref_v, ref_g = O.objective_value_and_grad(ev, motion, (H, W), **{k: v for k, v in kw.items() if k != "n_bins"})
    ref_v64, _ = O.objective_value_and_grad(ev.double(), motion.double(), (H, W), **{k: v for k, v in kw.items() if k != "n_bins"})
    assert _cost_close(v5, ref_v, ref_v64), (float(v5), float(ref_v), float(ref_v64))
    assert abs(float(v5) - float(v2)) <= 1e-6 * abs(float(v2))
    assert _rel(g5.cpu().numpy(), ref_g.numpy()) <= RTOL, _rel(g5.cpu().numpy(), ref_g.numpy())
    assert _rel(g5.cpu().numpy(), g2.cpu().numpy()) <= 1e-5
    # the un-blurred IWE stack of the strip K1 against the oracle's images
    obj.plan.set_variant(5, 5)
    iwe = obj.iwe(motion.to(dev)).cpu()
    _, images = O.objective(ev, motion, (H, W), motion_model=model, cost=cost, sigma=0.0, return_images=True)
    assert iwe.shape[0] == len(obj.ref_keys)
    for r, key in enumerate(obj.ref_keys):
        torch.testing.assert_close(iwe[r], images[key], rtol=1e-5, atol=2e-4)


def test_very_dense_batch_runs_the_segmented_gradient_reduction(B, dev):
    """>= 16 strips per source pixel (what a spatially compact shard of a multi-GPU run looks like): the strip K3 sums the flow
    gradient over the strips of one pixel inside the warp before its reductions.  Same numbers as the oracle and as the run
    kernels; a batch that covers only a band of the image (rows 8..23) exercises the row bookkeeping of the plan too."""
    rng = np.random.default_rng(23)
    H, W, n = 32, 40, 160_000
    ev = np.stack([rng.integers(8, 24, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1).astype(np.float32)
    ev = torch.from_numpy(ev)
    flow = torch.from_numpy(rng.uniform(-5, 5, (2, H, W)).astype(np.float32))
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="dense-flow")
    assert obj.plan.n_strips >= 16 * 16 * W  # 250 events per pixel of the band -> the segmented path is on
    v5, g5 = obj.value_and_grad(flow.to(dev))
    obj.plan.set_variant(2, 2)
    v2, g2 = obj.value_and_grad(flow.to(dev))
    ref_v, ref_g = O.objective_value_and_grad(ev, flow, (H, W), motion_model="dense-flow", cost="image_variance")
    ref_v64, _ = O.objective_value_and_grad(ev.double(), flow.double(), (H, W), motion_model="dense-flow", cost="image_variance")
    assert _cost_close(v5, ref_v, ref_v64)
    assert _rel(g5.cpu().numpy(), ref_g.numpy()) <= RTOL, _rel(g5.cpu().numpy(), ref_g.numpy())
    assert _rel(g5.cpu().numpy(), g2.cpu().numpy()) <= RTOL
    assert float(g5[:, :8].abs().max()) == 0.0 and float(g5[:, 24:].abs().max()) == 0.0  # no source pixels outside the band


def test_cuda_graph_cached_evaluation(B, dev, golden_c1):
    """`cuda_graph=True`: every evaluation replays a graph captured once per variant (config 1's regime: 30 k events, launch
    bound).  Same numbers as the eager calls, for several motions in a row, value-only and value + gradient interleaved."""
    g = golden_c1
    ev = torch.from_numpy(g["events"]).float().to(dev)
    eager = B.ContrastObjective(ev, (260, 346), cost="image_variance", motion_model="2d-translation")
    graphed = B.ContrastObjective(ev, (260, 346), cost="image_variance", motion_model="2d-translation", cuda_graph=True)
    rng = np.random.default_rng(3)
    for k in range(6):
        theta = torch.from_numpy(rng.uniform(-20, 20, 2).astype(np.float32)).to(dev)
        v0, g0 = eager.value_and_grad(theta)
        if k % 2:
            v1 = graphed.value(theta)
            assert abs(float(v1) - float(v0)) <= 1e-6 * abs(float(v0))
        else:
            v1, g1 = graphed.value_and_grad(theta)
            assert abs(float(v1) - float(v0)) <= 1e-6 * abs(float(v0))
            assert _rel(g1.cpu().numpy(), g0.cpu().numpy()) <= 1e-5
    m = torch.tensor([3.0, -7.0], dtype=torch.float64, device=dev, requires_grad=True)
    (ga,) = torch.autograd.grad(graphed(m), m)
    (gb,) = torch.autograd.grad(eager(m), m)
    assert _rel(ga.cpu().numpy(), gb.cpu().numpy()) <= 1e-5
