"""Full-size parity on the GPU, through the C ABI, against the CPU oracle:
  * BASELINE config 2 as benched -- 5 M events, 260x346, dense flow, variance cost + gradient, strip kernels;
  * BASELINE config 3's shape -- 480x640, 16x16 tile flow -> Burgers voxel (T=10) -> time-aware warp, gradient magnitude,
    2 M events (what the fp32 + fp64 oracle finishes in well under a minute on the box's host cores).
Bounds = the north-star tolerance: cost within 1e-5 of the fp64 oracle; gradient within 1e-5 norm-wise of the SAME-dtype
(fp32) oracle (the objective's gradient is discontinuous where a floor index flips, so fp32 and fp64 references differ from
each other by far more than that -- the tests print that distance next to the measured errors); IWE within 1e-5 of its peak."""
import numpy as np
import pytest
import torch

import bench
from oracle import cm_oracle as O

pytestmark = pytest.mark.gpu


def _norm_rel(a, b):
    a, b = a.double(), b.double()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


def test_config2_full_size_vs_oracle():
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    cfg = bench.CONFIGS["c2"]
    H, W, n = cfg["H"], cfg["W"], cfg["events"]
    ev = torch.from_numpy(bench.synth_events(n, 0))
    flow = torch.from_numpy(bench.synth_flows(2, 100)[1])
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost="image_variance", motion_model="dense-flow")
    assert obj.plan.n_strips > 0, "config 2 runs the strip kernels"
    val, grad = obj.value_and_grad(flow.to(dev))
    iwe = obj.iwe(flow.to(dev))[0].cpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    v64, g64 = O.objective_value_and_grad(ev.double(), flow.double(), (H, W), motion_model="dense-flow", cost="image_variance")
    v32, g32 = O.objective_value_and_grad(ev, flow, (H, W), motion_model="dense-flow", cost="image_variance")
    _, images = O.objective(ev, flow, (H, W), motion_model="dense-flow", cost="image_variance", return_images=True)
    rel_v = abs(float(val) - float(v64)) / abs(float(v64))
    rel_g = _norm_rel(grad.cpu(), g32)
    rel_i = float((iwe - images["iwe"]).abs().max() / images["iwe"].abs().max())
    print(f"config 2 @ {n} events: cost rel vs fp64 oracle {rel_v:.2e}; grad rel vs fp32 oracle {rel_g:.2e} "
          f"(fp32 oracle vs fp64 oracle: {_norm_rel(g32, g64):.2e}); IWE max err / peak {rel_i:.2e}")
    assert rel_v <= 1e-5, rel_v
    assert rel_g <= 1e-5, rel_g
    assert rel_i <= 1e-5, rel_i
    # the benchmark's own entry point gives the same numbers
    c = torch.zeros(1, dtype=torch.float64, device=dev)
    g = torch.zeros(2, H, W, device=dev)
    obj.step_into(flow.to(dev), c, g)
    assert abs(float(c) - float(val)) <= 1e-6 * abs(float(val))  # (fp32 atomics: the summation order varies run to run)
    assert _norm_rel(g.cpu(), grad.cpu()) <= 1e-6


@pytest.mark.parametrize("cost,sigma", [("gradient_magnitude", 0.0), ("multi_focal_normalized_gradient_magnitude", 1.0)])
def test_config3_shape_time_aware_vs_oracle(cost, sigma):
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    cfg = dict(bench.CONFIGS["c3"], cost=cost, sigma=sigma)
    H, W, T, n = cfg["H"], cfg["W"], cfg["T"], 2_000_000
    ev = torch.from_numpy(bench.synth_events(n, 3, H, W))
    motion = torch.from_numpy(bench.synth_motions(cfg, 1, 7)[0])
    obj = B.ContrastObjective(ev.to(dev), (H, W), cost=cost, motion_model="dense-flow-voxel", n_bins=T, sigma=sigma, orig_events=ev.to(dev))
    assert obj.plan.n_strips > 0, "6.5 events per pixel still qualify for strips"
    tobj = B.TimeAwareObjective(obj, scheme="burgers", t0_location="middle",
                                tile=dict(patch_size=cfg["window"], sliding_window=cfg["window"], patch_shift=(0, 0)))
    val, gm = tobj.value_and_grad(motion.to(dev))
    v64, g64 = bench.oracle_value_and_grad(cfg, ev, motion, torch.float64)
    v32, g32 = bench.oracle_value_and_grad(cfg, ev, motion, torch.float32)
    rel_v = abs(float(val) - float(v64)) / abs(float(v64))
    rel_g = _norm_rel(gm.cpu(), g32)
    print(f"config 3 shape ({cost}) @ {n} events: cost rel vs fp64 oracle {rel_v:.2e}; grad rel vs fp32 oracle {rel_g:.2e} "
          f"(fp32 oracle vs fp64 oracle: {_norm_rel(g32, g64):.2e})")
    assert rel_v <= 1e-5, rel_v
    assert rel_g <= 1e-5, rel_g
    # ... and through the allocation-free entry point the benchmark replays from a CUDA graph
    c = torch.zeros(1, dtype=torch.float64, device=dev)
    g = torch.zeros_like(gm)
    tobj.step_into(motion.to(dev), c, g)
    assert abs(float(c) - float(val)) <= 1e-6 * abs(float(val))
    assert _norm_rel(g.cpu(), gm.cpu()) <= 1e-5


def test_batched_variance_plugin_is_the_pooled_variance():
    """`torch.var` of a cropped batch (src/costs/image_variance.py:52-58) = the pooled unbiased variance, not the mean of the
    per-image variances."""
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    batch = torch.from_numpy((rng.uniform(0, 5, (3, 20, 30)) + np.arange(3)[:, None, None] * 4.0).astype(np.float32)).to(dev)
    for omit in (True, False):
        ref = torch.var(batch[..., 1:-1, 1:-1] if omit else batch)
        x = batch.clone().requires_grad_(True)
        got = B.cost_functions["image_variance"](direction="natural").calculate({"iwe": x, "omit_boundary": omit})
        assert abs(float(got) - float(ref)) <= 1e-5 * abs(float(ref)), (omit, float(got), float(ref))
        (g,) = torch.autograd.grad(got, x)
        xr = batch.clone().requires_grad_(True)
        (gr,) = torch.autograd.grad(torch.var(xr[..., 1:-1, 1:-1] if omit else xr), xr)
        assert _norm_rel(g, gr) <= 1e-5
