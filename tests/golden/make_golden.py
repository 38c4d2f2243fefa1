"""Generate golden vectors by running the UNMODIFIED reference (tub-rip/event_based_optical_flow,
mounted read-only at /root/reference) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes tests/golden/*.npz.  The only concession to the container is an import-only stub for
`optuna` (src/utils/misc.py:12 imports it; nothing on the hot path calls it).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CM_REFERENCE", "/root/reference")


def _import_reference():
    stub_dir = tempfile.mkdtemp(prefix="cmstub")
    os.makedirs(os.path.join(stub_dir, "optuna"))
    with open(os.path.join(stub_dir, "optuna", "__init__.py"), "w") as f:
        f.write(
            "import types as _t\n"
            "class _S: pass\n"
            "storages = _t.SimpleNamespace(InMemoryStorage=_S)\n"
            "distributions = _t.SimpleNamespace(BaseDistribution=object)\n"
            "samplers = _t.SimpleNamespace(); study = _t.SimpleNamespace(Study=object)\n"
            "logging = _t.SimpleNamespace(set_verbosity=lambda *a, **k: None, WARNING=30)\n"
        )
    sys.path.insert(0, stub_dir)
    sys.path.insert(0, REF)
    from src import costs, event_image_converter, utils, warp  # noqa

    return types.SimpleNamespace(costs=costs, eic=event_image_converter, utils=utils, warp=warp)


def synth_events(rng, n, H, W, tmax=0.05, fractional=False):
    """Same distributions as src/utils/event_utils.py:18-47 but from a seeded Generator."""
    x = rng.integers(0, H, n).astype(np.float64)
    y = rng.integers(0, W, n).astype(np.float64)
    if fractional:  # undistorted-style events: non-integer coordinates strictly inside the frame
        x = np.clip(x + rng.uniform(0, 1, n), 0, H - 1e-3)
        y = np.clip(y + rng.uniform(0, 1, n), 0, W - 1e-3)
    t = np.sort(rng.uniform(0.0, tmax, n))
    p = rng.integers(0, 2, n).astype(np.float64)
    return np.stack([x, y, t, p], axis=1)


def main():
    R = _import_reference()
    torch.manual_seed(0)
    out = {}

    # ---------------------------------------------------------------- random cases
    cases = {
        "small": dict(H=24, W=32, n=1200, fmax=6.0, T=4, frac=False, seed=1),
        "frac": dict(H=20, W=28, n=1000, fmax=4.0, T=5, frac=True, seed=2),
        "tiny": dict(H=5, W=7, n=40, fmax=3.0, T=3, frac=False, seed=3),
    }
    cost_names = ["image_variance", "gradient_magnitude", "normalized_image_variance",
                  "normalized_gradient_magnitude", "multi_focal_normalized_image_variance",
                  "multi_focal_normalized_gradient_magnitude"]
    for cname, c in cases.items():
        rng = np.random.default_rng(c["seed"])
        H, W, n, T = c["H"], c["W"], c["n"], c["T"]
        ev64 = synth_events(rng, n, H, W, fractional=c["frac"])
        flow64 = rng.uniform(-c["fmax"], c["fmax"], (2, H, W))
        voxel64 = rng.uniform(-c["fmax"], c["fmax"], (T, 2, H, W))
        theta64 = rng.uniform(-c["fmax"], c["fmax"], 2)
        wgt64 = rng.uniform(-1.0, 2.0, n)
        out[f"{cname}/events"] = ev64
        out[f"{cname}/flow"] = flow64
        out[f"{cname}/voxel"] = voxel64
        out[f"{cname}/theta"] = theta64
        out[f"{cname}/weight"] = wgt64
        out[f"{cname}/meta"] = np.array([H, W, n, T])
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            ev = torch.from_numpy(ev64).to(dtype)
            flow = torch.from_numpy(flow64).to(dtype)
            voxel = torch.from_numpy(voxel64).to(dtype)
            theta = torch.from_numpy(theta64).to(dtype)
            wgt = torch.from_numpy(wgt64).to(dtype)
            warper = R.warp.Warp((H, W), normalize_t=True)
            for pad in (0, 3):
                imager = R.eic.EventImageConverter((H, W), outer_padding=pad)
                for d in ("first", "middle", "last", 0.3):
                    dtag = d if isinstance(d, str) else f"d{d}"
                    for model, motion in (("dense-flow", flow), ("dense-flow-voxel", voxel), ("2d-translation", theta)):
                        w, _ = warper.warp_event(ev, motion, model, direction=d)
                        if pad == 0:
                            out[f"{cname}/{tag}/warp/{model}/{dtag}"] = w.numpy()
                        for sigma in (0, 1):
                            iwe = imager.create_iwe(w, "bilinear_vote", sigma)
                            out[f"{cname}/{tag}/iwe/p{pad}/{model}/{dtag}/s{sigma}"] = iwe.numpy()
                # raw converter products on the un-warped events
                # method="count": the reference's torch branch raises (int64 `vals` into a float image,
                # src/event_image_converter.py:251-254), so the numpy branch (:161-207) is the only live oracle
                out[f"{cname}/{tag}/count/p{pad}"] = imager.create_iwe(ev.numpy(), "count", 0)
                out[f"{cname}/{tag}/vote_weighted/p{pad}"] = imager.bilinear_vote_tensor(ev, weight=wgt).numpy()
                out[f"{cname}/{tag}/orig_iwe/p{pad}/s0"] = imager.create_iwe(ev, "bilinear_vote", 0).numpy()
                out[f"{cname}/{tag}/orig_iwe/p{pad}/s1"] = imager.create_iwe(ev, "bilinear_vote", 1).numpy()

            # costs + autograd gradients through the reference (pad 0), the way get_arg_for_cost composes them
            imager = R.eic.EventImageConverter((H, W), outer_padding=0)
            precision = "64" if dtype == torch.float64 else "32"
            for sigma in (0, 1):
                for model, motion in (("dense-flow", flow), ("dense-flow-voxel", voxel), ("2d-translation", theta)):
                    for cn in cost_names:
                        m = motion.clone().requires_grad_(True)
                        fn = R.costs.functions[cn](direction="minimize", store_history=False, precision=precision)
                        arg = {"omit_boundary": True, "clip": True}
                        arg["orig_iwe"] = imager.create_iwe(ev, "bilinear_vote", sigma)
                        for key, d in (("backward_iwe", "first"), ("forward_iwe", "last"), ("middle_iwe", "middle")):
                            w, _ = warper.warp_event(ev, m, model, direction=d)
                            arg[key] = imager.create_iwe(w, "bilinear_vote", sigma)
                        arg["iwe"] = arg["backward_iwe"]
                        loss = fn.calculate(arg)
                        (g,) = torch.autograd.grad(loss, m)
                        out[f"{cname}/{tag}/cost/{model}/{cn}/s{sigma}"] = np.array(loss.item())
                        out[f"{cname}/{tag}/grad/{model}/{cn}/s{sigma}"] = g.numpy()
                # omit_boundary False (FWL metric path, src/solver/base.py:608-610)
                w, _ = warper.warp_event(ev, flow, "dense-flow", direction="first")
                iwe = imager.create_iwe(w, "bilinear_vote", 0)
                for cn in ("image_variance", "gradient_magnitude"):
                    fn = R.costs.functions[cn](direction="minimize", precision=precision)
                    out[f"{cname}/{tag}/cost_full/{cn}"] = np.array(fn.calculate({"iwe": iwe, "omit_boundary": False}).item())
    np.savez_compressed(os.path.join(HERE, "reference_random.npz"), **out)

    # ---------------------------------------------------------------- BASELINE config 1 (C1): 30k events, 346x260
    out = {}
    H, W, n = 260, 346, 30000
    np.random.seed(0)
    ev64 = R.utils.generate_events(n, H, W, tmin=0.0, tmax=0.05)   # the reference's own fixture, seeded
    flow64 = R.utils.generate_dense_optical_flow((H, W), max_val=10)
    theta64 = np.array([7.5, -4.25])
    out["events"] = ev64.astype(np.float32)
    out["theta"] = theta64
    ev = torch.from_numpy(ev64).float()
    warper = R.warp.Warp((H, W), normalize_t=True)
    imager = R.eic.EventImageConverter((H, W), outer_padding=0)
    var = R.costs.functions["image_variance"](direction="minimize")
    # 2-dof, variance (the CPU-reference configuration)
    th = torch.from_numpy(theta64).float().requires_grad_(True)
    w, _ = warper.warp_event(ev, th, "2d-translation", direction="first")
    iwe = imager.create_iwe(w, "bilinear_vote", 0)
    loss = var.calculate({"iwe": iwe, "omit_boundary": True})
    (g,) = torch.autograd.grad(loss, th)
    out["c1/2dof/cost"] = np.array(loss.item())
    out["c1/2dof/grad"] = g.numpy()
    out["c1/2dof/iwe"] = iwe.detach().numpy()
    # dense flow, variance (the metric path at C1 size)
    fl = torch.from_numpy(flow64).float().requires_grad_(True)
    out["flow"] = flow64.astype(np.float32)
    w, _ = warper.warp_event(ev, fl, "dense-flow", direction="first")
    iwe = imager.create_iwe(w, "bilinear_vote", 0)
    loss = var.calculate({"iwe": iwe, "omit_boundary": True})
    (g,) = torch.autograd.grad(loss, fl)
    out["c1/dense/cost"] = np.array(loss.item())
    out["c1/dense/grad"] = g.numpy()
    out["c1/dense/iwe"] = iwe.detach().numpy()
    out["c1/dense/warped_xy"] = w.detach().numpy()[:, :2]
    np.savez_compressed(os.path.join(HERE, "reference_c1.npz"), **out)

    # ---------------------------------------------------------------- the reference's own hand-computed vectors
    # (tests/test_warp.py:96-139, tests/test_event_image_converter.py:17-69) re-stated as data and re-verified
    # against the live reference here, so the fixtures can travel to the GPU box.
    out = {}
    ev = np.array([[1, 2, 0], [2, 3, 0.2], [0, 1, 0.6], [1, 0, 1.0]])
    flow = np.array([[[1.0, -0.5, 2, 8], [-2, 0, 2.0, 0], [2, 1, -2, 0]],
                     [[-10, 1.0, 3, 2], [0, 2, -0.9, 0], [0, 10, -3, 0]]])
    expected = np.array([[1.0, 2.0, 0], [2.0, 3.0, 0.2], [0.3, 0.4, 0.6], [3, 0, 1.0]])
    w, _ = R.warp.Warp((3, 4), normalize_t=True).warp_event(torch.from_numpy(ev), torch.from_numpy(flow), "dense-flow")
    assert torch.allclose(w, torch.from_numpy(expected))
    out["warp34/events"], out["warp34/flow"], out["warp34/expected"] = ev, flow, expected
    imager = R.eic.EventImageConverter((3, 4))
    e1 = np.array([[1.0, 2], [0, 1], [1, 0]])
    w1 = np.array([1, 2, 0.8])
    x1 = np.array([[0, 2, 0, 0], [0.8, 0, 1, 0], [0, 0, 0, 0]])
    assert torch.allclose(imager.bilinear_vote_tensor(torch.from_numpy(e1), weight=torch.from_numpy(w1)), torch.from_numpy(x1))
    e2 = np.array([[1.2, 2], [0, 1.9], [0.5, 0.6]])
    w2 = np.array([-1.0, 1.0, 1.5])
    x2 = np.array([[0.3, 0.55, 0.9, 0], [0.3, 0.45, -0.8, 0], [0, 0, -0.2, 0]])
    assert torch.allclose(imager.bilinear_vote_tensor(torch.from_numpy(e2), weight=torch.from_numpy(w2)), torch.from_numpy(x2))
    out["vote34/int/events"], out["vote34/int/weight"], out["vote34/int/expected"] = e1, w1, x1
    out["vote34/frac/events"], out["vote34/frac/weight"], out["vote34/frac/expected"] = e2, w2, x2
    np.savez_compressed(os.path.join(HERE, "reference_handvectors.npz"), **out)
    for f in ("reference_random.npz", "reference_c1.npz", "reference_handvectors.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
