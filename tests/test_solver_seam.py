"""The solver seam: `B200CostMixin.calculate_cost` against (a) the reference's own composition of warp -> IWE -> cost
(restated below from src/solver/patch_contrast_base.py:273-352, running on the drop-in CUDA operator classes) and
(b) the CPU oracle; plus an end-to-end scipy L-BFGS-B loop driven by the fused objective."""
import numpy as np
import pytest
import torch

from oracle import cm_oracle as O

pytestmark = pytest.mark.gpu


class _ReferenceSeam:
    """What the reference solver does per objective call (src/solver/patch_contrast_base.py:273-352), verbatim in
    structure; it is the base class the mixin is mixed into in these tests."""

    def __init__(self, B, image_shape, cost_name, sigma, cost_with_weight=None, pad=0):
        self.iwe_config = {"method": "bilinear_vote", "blur_sigma": sigma}
        self.imager = B.EventImageConverter(image_shape, outer_padding=pad)
        self.warper = B.Warp(image_shape, normalize_t=True)
        if cost_name == "hybrid":
            from event_based_optical_flow_b200.costs import HybridCost
            self.cost_func = HybridCost("minimize", cost_with_weight, store_history=True, precision="64")
        else:
            self.cost_func = B.cost_functions[cost_name](direction="minimize", store_history=True, precision="64")

    def calculate_cost(self, events, warp, motion_model, coarse_flow=None, save_intermediate_result=True):
        return self.cost_func.calculate(self.get_arg_for_cost(events, warp, motion_model, coarse_flow))

    def get_arg_for_cost(self, events, warp, motion_model, coarse_flow=None):
        arg = {"omit_boundary": True, "clip": True}
        keys = self.cost_func.required_keys
        m, s = self.iwe_config["method"], self.iwe_config["blur_sigma"]
        if "orig_iwe" in keys:
            arg["orig_iwe"] = self.imager.create_iwe(events, m, s)
        if "iwe" in keys or "backward_iwe" in keys or "backward_warp" in keys:
            ev, _ = self.warper.warp_event(events, warp, motion_model, direction="first")
            iwe = self.imager.create_iwe(ev, m, s)
            arg.update({"iwe": iwe, "backward_iwe": iwe, "backward_warp": ev})
        if "forward_iwe" in keys or "forward_warp" in keys:
            ev, _ = self.warper.warp_event(events, warp, motion_model, direction="last")
            arg.update({"forward_iwe": self.imager.create_iwe(ev, m, s), "forward_warp": ev})
        if "middle_iwe" in keys:
            ev, _ = self.warper.warp_event(events, warp, motion_model, direction="middle")
            arg["middle_iwe"] = self.imager.create_iwe(ev, m, s)
        if "flow" in keys:
            arg["flow"] = coarse_flow
        return arg


def _problem(seed=3, H=48, W=64, n=30000, patch=(4, 4)):
    rng = np.random.default_rng(seed)
    ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
    motion = rng.uniform(-5, 5, (2,) + patch)
    return torch.from_numpy(ev).double(), torch.from_numpy(motion).double(), (H, W)


def _dense(motion, image_shape):
    """A differentiable tile-flow -> dense-flow map standing in for the reference's upsample (torch ops, unchanged caller code)."""
    return torch.nn.functional.interpolate(motion[None], size=image_shape, mode="bilinear", align_corners=False)[0]


@pytest.mark.parametrize("cost_name,sigma", [("image_variance", 0), ("gradient_magnitude", 1),
                                             ("multi_focal_normalized_gradient_magnitude", 1), ("hybrid", 1)])
def test_mixin_matches_reference_composition_and_oracle(cost_name, sigma):
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.solver import B200CostMixin
    dev = torch.device("cuda:0")
    ev, motion, shape = _problem()
    weights = {"multi_focal_normalized_gradient_magnitude": 1.0, "total_variation": 0.01}

    class Fast(B200CostMixin, _ReferenceSeam):
        pass

    ref = _ReferenceSeam(B, shape, cost_name, sigma, weights)
    fast = Fast(B, shape, cost_name, sigma, weights)
    evd = ev.to(dev).requires_grad_()  # the reference hands the solver float64 events with requires_grad (pyramid.py:186)
    out = {}
    for tag, slv in (("ref", ref), ("fast", fast)):
        m = motion.clone().to(dev).requires_grad_(True)
        loss = slv.calculate_cost(evd, _dense(m, shape), "dense-flow", m)
        (g,) = torch.autograd.grad(loss, m)
        out[tag] = (float(loss), g.cpu().numpy())
        assert loss.dtype == torch.float64
    assert abs(out["fast"][0] - out["ref"][0]) <= 1e-5 * abs(out["ref"][0])
    assert np.linalg.norm(out["fast"][1] - out["ref"][1]) <= 1e-5 * np.linalg.norm(out["ref"][1])
    hist = fast.cost_func.get_history()
    assert len(hist["loss"]) == 1
    # the oracle (fp32, same dtype as the kernels) for the contrast part
    if cost_name != "hybrid":
        m = motion.clone().float().requires_grad_(True)
        val = O.objective(ev.float(), _dense(m, shape), shape, motion_model="dense-flow", cost=cost_name, sigma=float(sigma))
        (g,) = torch.autograd.grad(val, m)
        assert abs(out["fast"][0] - float(val)) <= 1e-5 * abs(float(val))
        assert np.linalg.norm(out["fast"][1] - g.numpy()) <= 1e-5 * np.linalg.norm(g.numpy())
    # second call re-uses the resident plan
    (batch,) = fast._b200_cache().values()
    n_plans, n_obj = len(batch.plans), len(batch.objectives)
    fast.calculate_cost(evd, _dense(motion.to(dev), shape), "dense-flow", motion.to(dev))
    assert len(fast._b200_cache()) == 1 and len(batch.plans) == n_plans and len(batch.objectives) == n_obj
    assert batch.events is evd


def test_use_b200_operators_swaps_seam_objects():
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.solver import use_b200_operators
    slv = _ReferenceSeam(B, (20, 30), "image_variance", 0, pad=2)
    slv.cost_func.direction = "minimize"
    use_b200_operators(slv)
    # dual operators: the CUDA classes for CUDA tensors, the solver's original objects for numpy / CPU input
    assert isinstance(slv.imager._cuda, B.EventImageConverter) and tuple(slv.imager.image_size) == (24, 34)
    assert isinstance(slv.warper._cuda, B.Warp) and slv.warper.normalize_t is True
    assert slv.cost_func.name == "image_variance"


def test_scipy_lbfgs_recovers_translation():
    """Events generated by a known 2-dof translation; scipy L-BFGS-B on the fused objective must recover it
    (the outer optimiser stays scipy, as in src/solver/scipy_autograd/scipy_minimize.py:100-115)."""
    import scipy.optimize
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(11)
    H, W, n_edges, per_edge = 64, 80, 12, 1500
    # a few straight "edges" moving with constant velocity: x(t) = x0 - theta * dt  (so warping by +theta*dt realigns them)
    theta_true = np.array([6.0, -9.0])
    t = rng.uniform(0, 1, n_edges * per_edge)
    x0 = np.repeat(rng.uniform(12, H - 12, n_edges), per_edge) + rng.normal(0, 0.15, n_edges * per_edge)
    y0 = np.repeat(rng.uniform(15, W - 15, n_edges), per_edge) + rng.normal(0, 0.15, n_edges * per_edge)
    ev = np.stack([x0 - theta_true[0] * t, y0 - theta_true[1] * t, t, np.ones_like(t)], 1).astype(np.float32)
    ev = ev[np.argsort(ev[:, 2])]
    obj = B.ContrastObjective(torch.from_numpy(ev).to(dev), (H, W), cost="image_variance", motion_model="2d-translation", order="asis")

    def fun(x):
        val, grad = obj.value_and_grad(torch.tensor(x, dtype=torch.float32, device=dev))
        return float(val), grad.double().cpu().numpy()

    def fun_oracle(x):
        val, grad = O.objective_value_and_grad(torch.from_numpy(ev), torch.tensor(x, dtype=torch.float32), (H, W),
                                               motion_model="2d-translation", cost="image_variance")
        return float(val), grad.double().numpy()

    # multi-start as the reference does with its random / grid initialisation (patch_contrast_base.py / pyramid.py)
    starts = ([0.0, 0.0], [4.0, -6.0], [8.0, -12.0])
    best = min((scipy.optimize.minimize(fun, x0_, jac=True, method="L-BFGS-B") for x0_ in starts), key=lambda r: r.fun)
    best_oracle = min((scipy.optimize.minimize(fun_oracle, x0_, jac=True, method="L-BFGS-B") for x0_ in starts), key=lambda r: r.fun)
    # the contrast maximum sits within a pixel of the generating motion (the bilinear vote favours integer alignment) ...
    assert np.allclose(best.x, theta_true, atol=1.0), best.x
    assert best.fun < 0.5 * fun(np.zeros(2))[0]
    # ... and the same loop driven by the CPU oracle ends in the same place
    assert abs(best.fun - best_oracle.fun) <= 1e-3 * abs(best_oracle.fun), (best.fun, best_oracle.fun)
    assert np.allclose(best.x, best_oracle.x, atol=5e-2), (best.x, best_oracle.x)


@pytest.mark.parametrize("pyramid", [False, True])
def test_mixin_time_aware_motion_to_dense_flow(pyramid):
    """The time-aware seam: `motion_to_dense_flow` of the mixin (tile-flow upsample + Burgers voxel, both CUDA) against
    the reference's composition restated with the oracle's torch functions, value and gradient, for both reference
    signatures (time_aware_patch_contrast.py:42-80 and patch_contrast_pyramid.py:464-516)."""
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.solver import B200CostMixin
    dev = torch.device("cuda:0")
    ev, motion, shape = _problem(H=64, W=96, patch=(4, 6))
    window = (16, 16)

    class RefTimeAware(_ReferenceSeam):
        is_time_aware, scale_later, time_bin, flow_interpolation, t0_flow_location = True, True, 10, "burgers", "middle"
        image_shape, patch_size, sliding_window, patch_shift = shape, window, window, (0, 0)
        motion_vector_size, patch_image_size, filter_type, current_scale = 2, (4, 6), "bilinear", 1

        def interpolate_dense_flow_from_patch_tensor(self, m):
            return O.upsample_tile_flow(m, self.image_shape, self.patch_size, self.sliding_window, self.patch_shift)

        def motion_to_dense_flow(self, m, t_scale=1.0):
            if isinstance(m, dict):
                dense = self.interpolate_dense_flow_from_patch_tensor(m[self.current_scale])
                scale = dense.max()
                return O.flow_voxel(dense * t_scale / scale, self.time_bin, self.flow_interpolation, self.t0_flow_location) * scale / t_scale
            dense = self.interpolate_dense_flow_from_patch_tensor(m)
            scale = m.max()
            return O.flow_voxel(dense / scale, self.time_bin, self.flow_interpolation, self.t0_flow_location) * scale

    class Fast(B200CostMixin, RefTimeAware):
        pass

    ref = RefTimeAware(B, shape, "multi_focal_normalized_gradient_magnitude", 1)
    fast = Fast(B, shape, "multi_focal_normalized_gradient_magnitude", 1)
    cot = torch.from_numpy(np.random.default_rng(1).standard_normal((10, 2) + shape))
    out = {}
    for tag, slv, d in (("ref", ref, "cpu"), ("fast", fast, dev)):
        m = motion.clone().to(d).requires_grad_(True)
        vox = slv.motion_to_dense_flow({1: m}, 0.7) if pyramid else slv.motion_to_dense_flow(m)
        assert vox.shape == (10, 2) + shape and vox.dtype == torch.float64
        (g,) = torch.autograd.grad((vox * cot.to(d)).sum(), m)
        out[tag] = (vox.detach().cpu().numpy(), g.cpu().numpy())
    np.testing.assert_allclose(out["fast"][0], out["ref"][0], rtol=1e-5, atol=2e-5)
    assert np.linalg.norm(out["fast"][1] - out["ref"][1]) <= 1e-5 * np.linalg.norm(out["ref"][1])
    # ... and the whole objective through the mixin's calculate_cost on the voxel it produced
    evd = ev.to(dev)
    m = motion.clone().to(dev).requires_grad_(True)
    loss = fast.calculate_cost(evd, fast.motion_to_dense_flow(m), "dense-flow-voxel", m)
    (g,) = torch.autograd.grad(loss, m)
    m32 = motion.clone().float().requires_grad_(True)
    val = O.objective(ev.float(), ref.motion_to_dense_flow(m32), shape, motion_model="dense-flow-voxel",
                      cost="multi_focal_normalized_gradient_magnitude", sigma=1.0)
    (g_ref,) = torch.autograd.grad(val, m32)
    assert abs(float(loss) - float(val)) <= 1e-5 * abs(float(val))
    assert np.linalg.norm(g.cpu().numpy() - g_ref.numpy()) <= 1e-5 * np.linalg.norm(g_ref.numpy())


# ------------------------------------------------------------------------------------------------ cache / second order
def _seam_solver(B, shape, cost_name="image_variance", sigma=0):
    from event_based_optical_flow_b200.solver import B200CostMixin

    class Fast(B200CostMixin, _ReferenceSeam):
        pass

    return Fast(B, shape, cost_name, sigma)


def test_two_frames_of_equal_size_do_not_share_a_plan(dev=None):
    """The reference builds a fresh events tensor per frame with a fixed n_events_per_batch; when the previous frame's tensor
    is freed the caching allocator hands the same address to the next one.  The mixin's cache is keyed on tensor identity
    (and keeps the tensor alive), so frame 2 must be optimised against frame 2's events."""
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    shape, n = (48, 64), 20000
    slv = _seam_solver(B, shape)
    rng = np.random.default_rng(11)
    flow = torch.from_numpy(rng.uniform(-4, 4, (2,) + shape)).to(dev)
    results, ptrs = [], []
    for frame in range(3):
        ev_np = np.stack([rng.integers(0, shape[0], n), rng.integers(0, shape[1], n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
        events = torch.from_numpy(ev_np).double().requires_grad_().to(dev)  # what run_scipy_over_scale does per frame
        ptrs.append(events.data_ptr())
        for _ in range(2):  # several objective calls per frame hit the cache
            loss = slv.calculate_cost(events, flow, "dense-flow")
        ref = O.objective(torch.from_numpy(ev_np), flow.cpu(), shape, motion_model="dense-flow", cost="image_variance", sigma=0.0)
        results.append((float(loss), float(ref)))
        assert len(slv._b200_cache()) <= slv.b200_max_batches
        del events, loss
    for got, ref in results:
        assert abs(got - ref) <= 1e-5 * abs(ref), results
    assert len({r[1] for r in results}) == 3  # (the three frames really differ)


def test_mixin_supports_hessian_vector_products():
    """Newton-CG / trust-* go through torch.autograd.functional.vhp (scipy_autograd/torch_wrapper.py:51-73); the mixin builds
    its objective from an EventPlan, which must not lose the event tensor the second-order path needs."""
    import event_based_optical_flow_b200 as B
    dev = torch.device("cuda:0")
    ev, motion, shape = _problem(seed=5, n=20000)
    slv = _seam_solver(B, shape)
    evd = ev.to(dev)

    def cost_of(m):
        return slv.calculate_cost(evd, _dense(m, shape), "dense-flow", m)

    m0 = motion.to(dev)
    v = torch.from_numpy(np.random.default_rng(2).standard_normal(tuple(motion.shape))).to(dev)
    _, hv = torch.autograd.functional.vhp(cost_of, m0, v)
    f64 = lambda m: O.objective(ev, _dense(m, shape), shape, motion_model="dense-flow", cost="image_variance", sigma=0.0)  # noqa: E731
    _, hv_ref = torch.autograd.functional.vhp(f64, motion, v.cpu())
    f32 = lambda m: O.objective(ev.float(), _dense(m, shape), shape, motion_model="dense-flow", cost="image_variance", sigma=0.0)  # noqa: E731
    _, hv_ref32 = torch.autograd.functional.vhp(f32, motion.float(), v.cpu().float())
    rel = float(torch.linalg.norm(hv.cpu() - hv_ref) / torch.linalg.norm(hv_ref))
    rel32 = float(torch.linalg.norm(hv_ref32.double() - hv_ref) / torch.linalg.norm(hv_ref))
    print(f"H v through the mixin: rel vs fp64 oracle {rel:.2e} (fp32 oracle vs fp64 oracle: {rel32:.2e})")
    # the second-order path composes the modular fp32 kernels; its bound is the fp32 reference's own distance from fp64
    # (measured: 1.4e-5 here against 1e-5 .. 1e-4 for the fp32 oracle), never looser than 5e-5
    assert rel <= max(1e-5, min(5e-5, 3 * rel32)), (rel, rel32)


def test_mixin_objective_scipy_runs_the_fused_tile_flow():
    """`B200CostMixin.objective_scipy` (pyramid signature) evaluates the loss from the patch motion inside the event kernels;
    against the reference's own composition (interpolate -> * t_scale -> calculate_cost, restated below from
    src/solver/patch_contrast_pyramid.py:430-462) running on the CPU oracle."""
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.solver import B200CostMixin
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    H, W, n = 64, 96, 200_000
    ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
    ev = torch.from_numpy(ev).double()
    grid, window = (4, 6), (16, 16)

    class PyramidLike(_ReferenceSeam):
        is_time_aware = False
        filter_type = "bilinear"
        motion_model_for_dense_warp = "dense-flow"
        motion_vector_size = 2
        normalize_t_in_batch = True
        patch_image_size = grid
        patch_size = window
        sliding_window = window
        patch_shift = (0, 0)
        current_scale = 1

        def objective_scipy(self, motion_array, events, coarser_motion, suppress_log=False):
            raise AssertionError("the fused path must not fall through to the reference composition here")

    class Fast(B200CostMixin, PyramidLike):
        b200_fuse_tile_flow = True

    for cost_name, cw in (("image_variance", None), ("hybrid", {"multi_focal_normalized_gradient_magnitude": 1.0, "total_variation": 0.01})):
        slv = Fast(B, (H, W), cost_name, 1 if cost_name == "hybrid" else 0, cost_with_weight=cw)
        motion = torch.from_numpy(rng.uniform(-5, 5, (2,) + grid)).to(dev).requires_grad_(True)
        evd = ev.to(dev)
        loss = slv.objective_scipy(motion.reshape(-1), evd, {}, True)
        (g,) = torch.autograd.grad(loss, motion)
        assert loss.dtype == torch.float64
        (batch,) = slv._b200_cache().values()
        assert all(t.fused for t in batch.tile_objectives.values()) and len(batch.tile_objectives) == 1
        # the reference's composition on the oracle, fp32 like the kernels
        t_scale = float(ev[:, 2].max() - ev[:, 2].min())
        m = motion.detach().cpu().float().requires_grad_(True)
        dense = O.upsample_tile_flow(m, (H, W), window, window, (0, 0)) * t_scale
        sigma = 1.0 if cost_name == "hybrid" else 0.0
        if cost_name == "hybrid":
            from event_based_optical_flow_b200.costs import functions
            ref = O.objective(ev.float(), dense, (H, W), motion_model="dense-flow", cost="multi_focal_normalized_gradient_magnitude", sigma=sigma) \
                + 0.01 * functions["total_variation"]().calculate({"flow": m, "omit_boundary": True})
        else:
            ref = O.objective(ev.float(), dense, (H, W), motion_model="dense-flow", cost=cost_name, sigma=sigma)
        (g_ref,) = torch.autograd.grad(ref, m)
        assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref)), (cost_name, float(loss), float(ref))
        rel = float(torch.linalg.norm(g.cpu().double() - g_ref.double()) / torch.linalg.norm(g_ref.double()))
        assert rel <= 1e-5, (cost_name, rel)
