"""The scenarios of the reference's OWN test-suite for this path, run against the CUDA drop-in classes (same class names,
same calls, CUDA tensors instead of numpy / CPU tensors).  What each test checks is restated here; the expected numbers that
are not self-evident are derived in the comments.  Reference files: tests/test_warp.py, tests/test_event_image_converter.py,
tests/costs/test_image_variance.py, tests/costs/test_gradient_magnitude.py, tests/costs/test_hybrid.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _events(n, H, W, tmin=0.0, tmax=0.5, seed=0):
    rng = np.random.default_rng(seed)
    ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(tmin, tmax, n)), rng.integers(0, 2, n)], 1)
    return torch.from_numpy(ev.astype(np.float64)).to(DEV)


# ---- tests/test_warp.py:12-93
@pytest.mark.parametrize("model,size", [("2d-translation", 2), ("rigid-optical-flow", 2)])
def test_get_motion_vector_size(model, size):
    import event_based_optical_flow_b200 as B
    assert B.Warp((100, 200), normalize_t=True).get_motion_vector_size(model) == size


@pytest.mark.parametrize("normalize", [True, False])
def test_calculate_dt(normalize):
    """dt = t - ref, divided by its own span when normalize_t: [tmin,tmax]=[1,2], ref 1 -> [0,1]; [-1,1], ref 0 -> [-.5,.5]
    normalised, [-1,1] raw; [-1,1], ref -1 -> [0,1] normalised, [0,2] raw."""
    import event_based_optical_flow_b200 as B
    w = B.Warp((100, 200), normalize_t=normalize)
    for (tmin, tmax, ref), (lo_n, hi_n), (lo_r, hi_r) in (((1, 2, 1.0), (0, 1), (0, 1)), ((0, 0.5, 0), (0, 1), (0, 0.5)),
                                                         ((-1, 1, 0), (-0.5, 0.5), (-1, 1)), ((-1, 1, -1), (0, 1), (0, 2))):
        dt = w.calculate_dt(_events(300, 100, 200, tmin, tmax), ref)
        lo, hi = (lo_n, hi_n) if normalize else (lo_r, hi_r)
        np.testing.assert_allclose(float(dt.min()), lo, rtol=1e-2, atol=0.1)
        np.testing.assert_allclose(float(dt.max()), hi, rtol=1e-2, atol=0.1)
    batch = torch.stack([_events(300, 10, 20, 1, i + 2, seed=i) for i in range(4)])
    dt = B.Warp((10, 20), normalize_t=True).calculate_dt(batch, 1.0)
    assert dt.shape == (4, 300)
    np.testing.assert_allclose(dt.max(-1).values.cpu().numpy(), 1.0, rtol=1e-2, atol=0.1)
    np.testing.assert_allclose(dt.min(-1).values.cpu().numpy(), 0.0, rtol=1e-2, atol=0.1)


# ---- tests/test_event_image_converter.py:7-15, 72-122
def test_create_iwe_shape_and_batch_vote():
    import event_based_optical_flow_b200 as B
    imager = B.EventImageConverter((100, 200))
    assert imager.create_iwe(_events(1000, 100, 200)).shape == (100, 200)
    # a batch of two event sets with per-event weights on a 3x4 image.  Set 0 has integer coordinates: weight w lands on
    # pixel (x, y).  Set 1: (1.2, 2) w=-1 -> -0.8 at (1,2), -0.2 at (2,2); (0, 1.9) w=1 -> 0.1 at (0,1), 0.9 at (0,2);
    # (0.5, 0.6) w=1.5 -> 0.5*0.4*1.5 = 0.3 at (0,0) and (1,0), 0.5*0.6*1.5 = 0.45 at (0,1) and (1,1)  => (0,1) = 0.55
    ev = torch.tensor([[[1, 2], [0, 1], [1, 0]], [[1.2, 2], [0, 1.9], [0.5, 0.6]]], dtype=torch.float64, device=DEV)
    wt = torch.tensor([[1.0, 2.0, 0.8], [-1.0, 1.0, 1.5]], dtype=torch.float64, device=DEV)
    img = B.EventImageConverter((3, 4)).bilinear_vote_tensor(ev, weight=wt)
    expected = torch.tensor([[[0, 2, 0, 0], [0.8, 0, 1, 0], [0, 0, 0, 0]],
                             [[0.3, 0.55, 0.9, 0], [0.3, 0.45, -0.8, 0], [0, 0, -0.2, 0]]], dtype=torch.float64)
    assert img.shape == (2, 3, 4)
    torch.testing.assert_close(img.cpu(), expected, rtol=1e-6, atol=1e-6)


# ---- tests/costs/test_image_variance.py:12-77 and test_gradient_magnitude.py:12-76
@pytest.mark.parametrize("cost_name", ["image_variance", "gradient_magnitude"])
def test_history_and_sharpness_ordering(cost_name):
    import event_based_optical_flow_b200 as B
    imager = B.EventImageConverter((260, 346))
    iwe = imager.create_image_from_events_tensor(_events(1000, 260, 346, 0.1, 0.9), "bilinear_vote", weight=1.0)
    for store, expected_len in ((True, 2), (False, 0)):
        cost = B.cost_functions[cost_name](direction="minimize", store_history=store)
        a = cost.calculate({"iwe": iwe, "omit_boundary": True})
        b = cost.calculate({"iwe": iwe, "omit_boundary": True})
        hist = cost.get_history()["loss"]
        assert len(hist) == expected_len and float(a) == float(b)
        if store:
            assert hist[0] == hist[1] == float(a)
    # three events on a 10x40 image, blurred IWE (sigma 1): all apart ("blur") vs two on one pixel ("sharp").  The sharper
    # image has the larger statistic, so the loss is smaller when minimising (the loss is -stat) and larger otherwise.
    small = B.EventImageConverter((10, 40))
    apart = torch.tensor([[5.0, 10.0], [8.0, 3.0], [2.0, 2.0]], device=DEV)
    stacked = torch.tensor([[5.0, 10.0], [5.0, 10.0], [2.0, 2.0]], device=DEV)
    for direction, blur_is_smaller in (("natural", True), ("minimize", False), ("maximize", True)):
        cost = B.cost_functions[cost_name](direction=direction)
        v_blur = float(cost.calculate({"iwe": small.create_iwe(apart), "omit_boundary": False}))
        v_sharp = float(cost.calculate({"iwe": small.create_iwe(stacked), "omit_boundary": False}))
        assert (v_blur < v_sharp) == blur_is_smaller, (direction, v_blur, v_sharp)


# ---- tests/costs/test_hybrid.py:13-67
@pytest.mark.parametrize("store", [True, False])
def test_hybrid_cost_history(store):
    import event_based_optical_flow_b200 as B
    from event_based_optical_flow_b200.costs import HybridCost
    imager = B.EventImageConverter((20, 34))
    cost = HybridCost(direction="minimize", cost_with_weight={"image_variance": 1.0, "gradient_magnitude": 2.4}, store_history=store)
    variance = B.cost_functions["image_variance"](store_history=True)
    for seed in range(3):
        iwe = imager.create_image_from_events_tensor(_events(1000, 20, 34, seed=seed), "bilinear_vote", sigma=0)
        total = cost.calculate({"iwe": iwe, "omit_boundary": True})
        v = variance.calculate({"iwe": iwe, "omit_boundary": True})
        gm = B.cost_functions["gradient_magnitude"]().calculate({"iwe": iwe, "omit_boundary": True})
        assert abs(float(total) - (float(v) + 2.4 * float(gm))) <= 1e-6 * abs(float(total))
    history = cost.get_history()
    assert set(history) == {"loss", "image_variance", "gradient_magnitude"}
    for k in history:
        assert len(history[k]) == (3 if store else 0)
    if store:
        np.testing.assert_allclose(history["image_variance"], variance.get_history()["loss"], rtol=1e-5, atol=1e-5)
