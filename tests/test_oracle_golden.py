"""Pin the CPU oracle (oracle/cm_oracle.py) to outputs of the unmodified reference
(tests/golden/*.npz, written by tests/golden/make_golden.py) and to the reference's own
hand-computed vectors.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import cm_oracle as O

CASES = ("small", "frac", "tiny")
DTYPES = (("f32", torch.float32), ("f64", torch.float64))
DIRS = (("first", "first"), ("middle", "middle"), ("last", "last"), ("d0.3", 0.3))
MODELS = ("dense-flow", "dense-flow-voxel", "2d-translation")


def _inputs(g, case, dtype):
    H, W, n, T = (int(v) for v in g[f"{case}/meta"])
    t = lambda k: torch.from_numpy(g[f"{case}/{k}"]).to(dtype)
    return H, W, t("events"), {"dense-flow": t("flow"), "dense-flow-voxel": t("voxel"), "2d-translation": t("theta")}


def _warp(model, ev, motion, d):
    return {"dense-flow": O.warp_dense, "dense-flow-voxel": O.warp_voxel, "2d-translation": O.warp_2dof}[model](ev, motion, d)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag,dtype", DTYPES)
def test_warps_bit_exact(golden_random, case, tag, dtype):
    g = golden_random
    H, W, ev, motions = _inputs(g, case, dtype)
    for model in MODELS:
        for dtag, d in DIRS:
            got = _warp(model, ev, motions[model], d).numpy()
            np.testing.assert_array_equal(got, g[f"{case}/{tag}/warp/{model}/{dtag}"], err_msg=f"{model} {dtag}")


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag,dtype", DTYPES)
def test_iwe_bit_exact(golden_random, case, tag, dtype):
    g = golden_random
    H, W, ev, motions = _inputs(g, case, dtype)
    for pad in (0, 3):
        for model in MODELS:
            for dtag, d in DIRS:
                w = _warp(model, ev, motions[model], d)
                got = O.create_iwe(w, (H, W), (pad, pad), "bilinear_vote", 0).numpy()
                np.testing.assert_array_equal(got, g[f"{case}/{tag}/iwe/p{pad}/{model}/{dtag}/s0"])
                got = O.create_iwe(w, (H, W), (pad, pad), "bilinear_vote", 1).numpy()
                ref = g[f"{case}/{tag}/iwe/p{pad}/{model}/{dtag}/s1"]
                np.testing.assert_allclose(got, ref, rtol=1e-6 if tag == "f32" else 1e-13, atol=1e-7)
        wgt = torch.from_numpy(g[f"{case}/weight"]).to(dtype)
        got = O.bilinear_vote(ev, (H + 2 * pad, W + 2 * pad), (pad, pad), wgt).numpy()
        np.testing.assert_array_equal(got, g[f"{case}/{tag}/vote_weighted/p{pad}"])
        got = O.count_vote(ev, (H + 2 * pad, W + 2 * pad), (pad, pad)).numpy()
        np.testing.assert_array_equal(got, g[f"{case}/{tag}/count/p{pad}"])


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("tag,dtype", DTYPES)
@pytest.mark.parametrize("sigma", (0, 1))
def test_costs_and_autograd_grads(golden_random, case, tag, dtype, sigma):
    g = golden_random
    H, W, ev, motions = _inputs(g, case, dtype)
    rtol = 2e-5 if tag == "f32" else 1e-11
    for model in MODELS:
        for cn in O.COSTS:
            val, grad = O.objective_value_and_grad(ev, motions[model], (H, W), motion_model=model, cost=cn, sigma=float(sigma))
            ref_v = float(g[f"{case}/{tag}/cost/{model}/{cn}/s{sigma}"])
            ref_g = g[f"{case}/{tag}/grad/{model}/{cn}/s{sigma}"]
            assert abs(float(val) - ref_v) <= rtol * abs(ref_v), (model, cn)
            err = np.linalg.norm(grad.numpy() - ref_g) / max(np.linalg.norm(ref_g), 1e-30)
            assert err <= rtol, (model, cn, err)


@pytest.mark.parametrize("case", CASES)
def test_closed_form_gradients_match_reference_autograd(golden_random, case):
    """dL/dIWE images + per-event chain (the formulas the CUDA backward implements), fp64 so that
    the comparison is not limited by rounding."""
    g = golden_random
    H, W, ev, motions = _inputs(g, case, torch.float64)
    flow = motions["dense-flow"]
    for cn, dimg in (("image_variance", O.dvariance_dimage), ("gradient_magnitude", O.dgradmag_dimage)):
        for sigma in (0, 1):
            w = O.warp_dense(ev, flow, "first")
            pre = O.bilinear_vote(w, (H, W))
            iwe = O.gaussian_blur3(pre, sigma) if sigma else pre
            G = -dimg(iwe, True)
            if sigma:
                G = O.blur3_adjoint(G, sigma)
            gx, gy = O.event_gradient(w, G)
            grad = O.flow_gradient_dense(ev, w[:, 2], gx, gy, (H, W)).numpy()
            ref_g = g[f"{case}/f64/grad/dense-flow/{cn}/s{sigma}"]
            err = np.linalg.norm(grad - ref_g) / np.linalg.norm(ref_g)
            assert err < 1e-11, (cn, sigma, err)


def test_cost_without_boundary_crop(golden_random):
    g = golden_random
    for case in CASES:
        H, W, ev, motions = _inputs(g, case, torch.float64)
        iwe = O.create_iwe(O.warp_dense(ev, motions["dense-flow"], "first"), (H, W))
        np.testing.assert_allclose(-float(O.image_variance(iwe, False)), float(g[f"{case}/f64/cost_full/image_variance"]), rtol=1e-12)
        np.testing.assert_allclose(-float(O.gradient_magnitude(iwe, False)), float(g[f"{case}/f64/cost_full/gradient_magnitude"]), rtol=1e-12)


def test_reference_hand_vectors(golden_hand):
    """tests/test_warp.py:96-139 and tests/test_event_image_converter.py:17-69 of the reference."""
    g = golden_hand
    ev = torch.from_numpy(g["warp34/events"])
    ev4 = torch.cat([ev, torch.zeros(len(ev), 1, dtype=ev.dtype)], dim=1)
    w = O.warp_dense(ev4, torch.from_numpy(g["warp34/flow"]), "first")
    assert torch.allclose(w[:, :3], torch.from_numpy(g["warp34/expected"]))
    for k in ("int", "frac"):
        img = O.bilinear_vote(torch.from_numpy(g[f"vote34/{k}/events"]), (3, 4), (0, 0), torch.from_numpy(g[f"vote34/{k}/weight"]))
        assert torch.allclose(img, torch.from_numpy(g[f"vote34/{k}/expected"]))


def test_c1_config(golden_c1):
    """BASELINE config 1: 30k events, 346x260, 2-dof warp + variance; and the dense-flow metric path."""
    g = golden_c1
    ev = torch.from_numpy(g["events"])
    th = torch.from_numpy(g["theta"]).float()
    val, grad = O.objective_value_and_grad(ev, th, (260, 346), motion_model="2d-translation", cost="image_variance")
    assert abs(float(val) - float(g["c1/2dof/cost"])) <= 1e-6 * abs(float(g["c1/2dof/cost"]))
    np.testing.assert_allclose(grad.numpy(), g["c1/2dof/grad"], rtol=1e-4)
    flow = torch.from_numpy(g["flow"])
    w = O.warp_dense(ev, flow, "first")
    np.testing.assert_array_equal(w[:, :2].numpy(), g["c1/dense/warped_xy"])
    np.testing.assert_array_equal(O.bilinear_vote(w, (260, 346)).numpy(), g["c1/dense/iwe"])
    val, grad = O.objective_value_and_grad(ev, flow, (260, 346), motion_model="dense-flow", cost="image_variance")
    assert abs(float(val) - float(g["c1/dense/cost"])) <= 1e-6 * abs(float(g["c1/dense/cost"]))
    err = np.linalg.norm(grad.numpy() - g["c1/dense/grad"]) / np.linalg.norm(g["c1/dense/grad"])
    assert err < 1e-6


def test_voxel_bins_partition_all_events():
    t = torch.sort(torch.rand(1000, dtype=torch.float32) * 0.05).values
    for d in ("first", "middle", "last"):
        dt = O.normalised_dt(t, O.reference_time(t, d))
        b = O.voxel_bin_of(dt, 10)
        assert int(b.min()) == 0 and int(b.max()) == 9
        assert bool((b[1:] >= b[:-1]).all())


@pytest.fixture(scope="module")
def golden_tileflow():
    import os
    from conftest import GOLDEN_DIR
    return np.load(os.path.join(GOLDEN_DIR, "reference_tileflow.npz"))


@pytest.mark.parametrize("case", ("s4", "s3", "s1", "sq", "odd"))
def test_tile_flow_upsample_matches_reference(golden_tileflow, case):
    """Tile-flow -> dense-flow and its autograd adjoint vs the reference's own method (patch_contrast_base.py:462-506)."""
    g = golden_tileflow
    meta = [int(v) for v in g[f"{case}/meta"]]
    image_shape, patch_size, sliding_window, patch_shift = tuple(meta[0:2]), tuple(meta[2:4]), tuple(meta[4:6]), tuple(meta[6:8])
    for tag, dt, tol in (("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-13)):
        m = torch.from_numpy(g[f"{case}/{tag}/motion"]).to(dt).requires_grad_(True)
        dense = O.upsample_tile_flow(m, image_shape, patch_size, sliding_window, patch_shift)
        ref = g[f"{case}/{tag}/dense"]
        assert dense.shape == ref.shape
        np.testing.assert_allclose(dense.detach().numpy(), ref, rtol=tol, atol=tol * 8)
        (gm,) = torch.autograd.grad((dense * torch.from_numpy(g[f"{case}/{tag}/grad_dense"]).to(dt)).sum(), m)
        np.testing.assert_allclose(gm.numpy(), g[f"{case}/{tag}/grad_motion"], rtol=tol * 10, atol=tol * 100)


FLOWVOXEL_CASES = ("up_mid10", "bg_mid10", "up_first7", "bg_first7", "bg_mid2", "bg_first2", "up_mid2", "up_mid3", "bg_mid3", "bg_first1",
                   "up_first1", "bg_mid21", "up_first20", "bg_mvsec")


@pytest.fixture(scope="module")
def golden_flowvoxel():
    import os
    from conftest import GOLDEN_DIR
    return np.load(os.path.join(GOLDEN_DIR, "reference_flowvoxel.npz"))


@pytest.mark.parametrize("case", FLOWVOXEL_CASES)
def test_flow_voxel_matches_reference(golden_flowvoxel, case):
    """Upwind / Burgers flow voxel and its autograd adjoint vs the reference's construct_dense_flow_voxel_torch
    (src/utils/flow_utils.py:99-161): BIT-EXACT voxel in fp32 and fp64, gradient to rounding."""
    g = golden_flowvoxel
    H, W, T, sc, mid = (int(v) for v in g[f"{case}/meta"])
    scheme, t0 = ("burgers" if sc else "upwind"), ("middle" if mid else "first")
    for tag, dt, tol in (("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-14)):
        if f"{case}/voxel_{tag}" not in g.files:
            continue
        x = torch.from_numpy(g[f"{case}/flow"]).to(dt).requires_grad_(True)
        vox = O.flow_voxel(x, T, scheme, t0)
        assert np.array_equal(vox.detach().numpy(), g[f"{case}/voxel_{tag}"]), (case, tag)
        (gr,) = torch.autograd.grad((vox * torch.from_numpy(g[f"{case}/cot"]).to(dt)).sum(), x)
        ref = g[f"{case}/grad_{tag}"]
        assert np.abs(gr.numpy() - ref).max() <= tol * max(np.abs(ref).max(), 1e-30), (case, tag)


def test_flow_voxel_rejects_unknown_arguments():
    x = torch.zeros(2, 4, 5)
    with pytest.raises(NotImplementedError):
        O.flow_voxel(x, 3, "upwind", "last")
    with pytest.raises(NotImplementedError):
        O.flow_voxel(x, 3, "nearest", "middle")


@pytest.mark.parametrize("scheme", ["upwind", "burgers"])
def test_flow_voxel_reference_test_cases(scheme):
    """The reference's own checks of the voxel construction (tests/utils/test_flow_utils.py:52-88), on the oracle: level t0
    holds the input flow for both t0 locations, a single level is the input (upwind), 60 levels on a 100x200 random flow."""
    rng = np.random.default_rng(0)
    flow = torch.from_numpy(rng.uniform(-20, 20, (2, 100, 200)))
    n_bin = 60
    assert torch.equal(O.flow_voxel(flow, 1, "upwind", "middle")[0], flow)
    assert torch.equal(O.flow_voxel(flow, n_bin, scheme, "middle")[n_bin // 2], flow)
    assert torch.equal(O.flow_voxel(flow, n_bin, scheme, "first")[0], flow)
    vox = O.flow_voxel(torch.from_numpy(rng.uniform(-1, 1, (2, 20, 30))), 100, scheme, "first")
    assert vox.shape == (100, 2, 20, 30) and bool(torch.isfinite(vox).all())
