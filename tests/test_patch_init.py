"""The per-patch 2-dof candidate cost of the reference's Optuna initialiser (SURVEY.md section 8f row 4;
src/solver/patch_contrast_pyramid.py:320-415): oracle against goldens made by the unmodified reference solver class (through
the real scipy / cv2), the batched CUDA evaluator against the oracle and the goldens, and the mixin's ask/tell loop."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import patch_init_oracle as PO

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_patch_init.npz")


def _golden():
    g = np.load(GOLD)
    ev = np.concatenate([g["events_xy"].astype(np.float64), g["events_t"][:, None], g["events_p"].astype(np.float64)[:, None]], 1)
    return g, ev


def test_oracle_matches_reference_objective_initial():
    g, ev = _golden()
    sigma, pad = float(g["sigma"]), int(g["padding"])
    worst = 0.0
    for scale in g["scales"]:
        rects, cand, loss = g[f"{scale}/rects"], g[f"{scale}/candidates"], g[f"{scale}/loss"]
        size = tuple(int(v) for v in g[f"{scale}/patch_image_size"])
        for i in range(len(rects)):
            if np.isnan(loss[i]).all():
                continue
            f = PO.crop_to_patch(ev, *rects[i])
            assert len(f) == g[f"{scale}/count"][i]
            for k in range(cand.shape[1]):
                mine = PO.candidate_loss(f, cand[i, k], size, (pad, pad), sigma)
                worst = max(worst, abs(mine - loss[i, k]) / abs(loss[i, k]))
            assert PO.candidate_loss(f, (0.0, 0.0), size, (pad, pad), sigma) == 1.0
    assert worst <= 1e-12, worst


def test_oracle_border_rules_against_scipy_and_cv2():
    """The two third-party stencils the reference calls, restated in the oracle: scipy's 'reflect' Gaussian and OpenCV's
    REFLECT_101 Sobel, on ragged shapes (down to fewer pixels than the Gaussian's radius)."""
    scipy_ndimage = pytest.importorskip("scipy.ndimage")
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for shape in ((16, 21), (3, 9), (2, 2), (7, 5), (33, 40)):
        img = rng.uniform(0, 5, shape)
        for sigma in (1.0, 0.6, 2.0):
            np.testing.assert_allclose(PO.gaussian_blur(img, sigma), scipy_ndimage.gaussian_filter(img, sigma), rtol=1e-12, atol=1e-13)
        gx = cv2.Sobel(img, cv2.CV_64F, 1, 0, ksize=3) / 8.0
        gy = cv2.Sobel(img, cv2.CV_64F, 0, 1, ksize=3) / 8.0
        assert abs(PO.sobel_energy(img) - np.mean(gx ** 2 + gy ** 2)) <= 1e-12 * np.mean(gx ** 2 + gy ** 2) + 1e-15, shape


# ------------------------------------------------------------------------------------------------ host side (CPU)
class FakeTrial:
    def __init__(self, rng, number):
        self.rng, self.number, self.params = rng, number, {}

    def suggest_uniform(self, key, low, high):
        self.params[key] = float(self.rng.uniform(low, high))
        return self.params[key]


class FakeStudy:
    """ask / tell / optimize / best_params of an Optuna study with a seeded uniform sampler."""

    def __init__(self, seed):
        self.rng, self.trials = np.random.default_rng(seed), []

    def ask(self):
        return FakeTrial(self.rng, len(self.trials))

    def tell(self, trial, value):
        self.trials.append((float(value), dict(trial.params)))

    def optimize(self, func, n_trials):
        i = 0
        while i < n_trials:  # Optuna's own loop condition: a float n_trials runs ceil(n_trials) trials
            t = self.ask()
            self.tell(t, func(t))
            i += 1

    @property
    def best_params(self):
        return min(self.trials, key=lambda v: v[0])[1]


def fake_optuna():
    mod = types.ModuleType("optuna")
    mod.studies = []

    def create_study(direction="minimize", sampler=None):
        assert direction == "minimize"
        mod.studies.append(FakeStudy(1000 + len(mod.studies)))
        return mod.studies[-1]

    mod.create_study = create_study
    mod.samplers = types.SimpleNamespace(TPESampler=lambda **kw: None)
    mod.logging = types.SimpleNamespace(set_verbosity=lambda *a, **k: None, WARNING=30)
    return mod


def test_group_patch_events_matches_the_reference_crop_and_time_normalisation():
    from event_based_optical_flow_b200.patch_init import group_patch_events
    g, ev = _golden()
    rects = g["3/rects"]
    rects = np.concatenate([rects, [[10, 60, 20, 90], [0, 1, 0, 1]]])  # + one overlapping the grid, one (nearly) empty
    pe, off, counts, scale = group_patch_events(torch.from_numpy(ev), torch.from_numpy(rects).double(), True)
    assert off[-1] == len(pe) and int(counts.sum()) == len(pe)
    for i in (0, 7, 33, len(rects) - 2, len(rects) - 1):
        f = PO.crop_to_patch(ev, *rects[i])
        assert counts[i] == len(f)
        mine = pe[off[i]:off[i + 1]].numpy()
        if len(f) < 2:
            continue
        w = PO.warp_2dof_middle(f, np.zeros(2))
        np.testing.assert_array_equal(mine[:, :2], f[:, :2].astype(np.float32))
        np.testing.assert_array_equal(mine[:, 2], w[:, 2].astype(np.float32))
        assert float(scale[i]) == f[:, 2].max() - f[:, 2].min()


class OracleEvaluator:
    """The evaluator's interface on the oracle (checks the study loop without a GPU)."""

    def __init__(self, ev, rects, size, sigma):
        self.patches = [PO.crop_to_patch(ev, *r) for r in rects]
        self.size, self.sigma, self.n_patches = size, sigma, len(rects)
        self.valid = np.array([len(f) > 10 for f in self.patches])

    def evaluate(self, cand):
        out = np.full(cand.shape[:2], np.nan)
        for i, f in enumerate(self.patches):
            if self.valid[i]:
                out[i] = [PO.candidate_loss(f, c, self.size, (0, 0), self.sigma) for c in cand[i]]
        return out


def test_run_patch_studies_equals_the_reference_loop_order():
    """Trial-major (batched) == patch-major (the reference's loop, pyramid.py:320-362): every study sees its own history only."""
    from event_based_optical_flow_b200.patch_init import n_trials_at, run_patch_studies, sampling_range
    g, ev = _golden()
    rects, size = g["3/rects"][:12], tuple(int(v) for v in g["3/patch_image_size"])
    ev = ev[::3]
    rng = np.random.default_rng(4)
    motion0 = rng.uniform(-8, 8, (2, len(rects)))
    evaluator = OracleEvaluator(ev, rects, size, 1.0)
    n_trials = n_trials_at(40, 3, 1)
    assert n_trials == 20 and n_trials_at(40, 4, 1) == 14
    mine = run_patch_studies(evaluator, motion0, n_trials, 8, optuna=fake_optuna())
    ref_mod, expect = fake_optuna(), motion0.copy()
    for i, f in enumerate(evaluator.patches):  # the reference's order
        if len(f) <= 10:
            continue
        study = ref_mod.create_study(direction="minimize", sampler=None)

        def objective(trial, f=f, m0=motion0[:, i]):
            (lx, hx), (ly, hy) = sampling_range(m0)
            c = (trial.suggest_uniform("trans_x", lx, hx), trial.suggest_uniform("trans_y", ly, hy))
            return PO.candidate_loss(f, c, size, (0, 0), 1.0)

        study.optimize(objective, n_trials=40 / 2)
        expect[:, i] = (study.best_params["trans_x"], study.best_params["trans_y"])
    np.testing.assert_array_equal(mine, expect)
    assert (mine != motion0).any()


def test_sampling_range_is_the_reference_rule():
    from event_based_optical_flow_b200.patch_init import sampling_range
    for m in ([3.0, -40.0], [0.0, 70.0], [-5.0, 5.0]):
        r = sampling_range(np.array(m))
        for k in range(2):
            c = [0.8 * m[k], m[k] - 10, 1.2 * m[k], m[k] + 10]  # src/solver/patch_contrast_pyramid.py:417-430
            assert r[k, 0] == min(c) and r[k, 1] == max(c)


# ------------------------------------------------------------------------------------------------ CUDA
@pytest.mark.gpu
@pytest.mark.parametrize("global_images", [None, True])  # the one-launch shared-memory kernel / the multi-kernel path for large patches
def test_cuda_candidates_match_reference_goldens(global_images):
    from event_based_optical_flow_b200.patch_init import PatchCandidateEvaluator
    g, ev = _golden()
    sigma, pad = float(g["sigma"]), int(g["padding"])
    for scale in g["scales"]:
        rects, cand, loss = g[f"{scale}/rects"], g[f"{scale}/candidates"], g[f"{scale}/loss"]
        size = tuple(int(v) for v in g[f"{scale}/patch_image_size"])
        evaluator = PatchCandidateEvaluator(torch.from_numpy(ev).cuda(), rects, size, outer_padding=pad, sigma=sigma, global_images=global_images)
        np.testing.assert_array_equal(evaluator.counts[g[f"{scale}/count"] > 0], g[f"{scale}/count"][g[f"{scale}/count"] > 0])
        mine = evaluator.evaluate(cand)
        done = ~np.isnan(loss)
        assert done.any() and not np.isnan(mine[done]).any()
        rel = np.abs(mine[done] - loss[done]) / np.abs(loss[done])
        print(f"scale {scale}: {done.sum()} reference losses, max rel {rel.max():.2e}")
        assert rel.max() <= 1e-5, (scale, rel.max())  # fp32 kernels against the reference's float64 numpy / scipy / cv2 chain
        assert (mine[evaluator.valid, 0] == 1.0).all()  # the zero candidate: integer coordinates vote exact 1.0s, whatever the order
        again = evaluator.evaluate(cand[:, :2])  # fewer candidates, same workspace
        rel2 = np.abs(again[done[:, :2]] - loss[:, :2][done[:, :2]]) / np.abs(loss[:, :2][done[:, :2]])
        assert rel2.max() <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("global_images", [None, True])
@pytest.mark.parametrize("sigma,pad", [(0.0, 0), (1.5, 2), (1.0, (1, 3)), (4.0, 0)])
def test_cuda_candidate_images_against_oracle(sigma, pad, global_images):
    """Blurred images and energies for ragged patches (overlapping, empty, shorter than the Gaussian radius), paddings, sigmas."""
    from event_based_optical_flow_b200.patch_init import PatchCandidateEvaluator
    g, ev = _golden()
    pads = (pad, pad) if isinstance(pad, int) else pad
    size = (9, 14)
    rects = np.array([[0, 9, 0, 14], [5, 14, 7, 21], [100, 109, 200, 214], [251, 260, 332, 346], [40, 49, 40, 54], [300, 309, 0, 14]])
    evaluator = PatchCandidateEvaluator(torch.from_numpy(ev).cuda(), rects, size, outer_padding=pad, sigma=sigma, min_events=3,
                                        global_images=global_images)
    rng = np.random.default_rng(9)
    cand = rng.uniform(-60, 60, (len(rects), 4, 2))
    mine = evaluator.evaluate(cand, keep_images=True)
    imgs = evaluator.images(4).cpu().numpy()
    assert not evaluator.valid[-1] and evaluator.counts[-1] == 0 and np.isnan(mine[-1]).all()
    for i in range(len(rects) - 1):
        f = PO.crop_to_patch(ev, *rects[i])
        assert len(f) == evaluator.counts[i]
        span = f[:, 2].max() - f[:, 2].min()
        for k in range(4):
            ref_img = PO.gaussian_blur(PO.bilinear_vote(PO.warp_2dof_middle(f, cand[i, k] * span), size, pads), sigma)
            np.testing.assert_allclose(imgs[i, k], ref_img, rtol=0, atol=2e-5 * max(1.0, ref_img.max()))
            expect = PO.candidate_loss(f, cand[i, k], size, pads, sigma)
            assert abs(mine[i, k] - expect) <= 1e-5 * abs(expect), (i, k, mine[i, k], expect)
    # raw energies of given thetas (no scaling, no ratio): the other form of the entry point
    e = evaluator.energies(torch.from_numpy(cand[:, :1] * evaluator.theta_scale.cpu().numpy()[:, None, None])).cpu().numpy()
    ok = evaluator.valid
    np.testing.assert_allclose(evaluator.orig_energy.cpu().numpy()[ok] / e[ok, 0], mine[ok, 0], rtol=1e-6)  # (two runs: the votes' fp32 summation order differs)


@pytest.mark.gpu
def test_cuda_candidates_argument_checks():
    from event_based_optical_flow_b200.patch_init import PatchCandidateEvaluator
    g, ev = _golden()
    with pytest.raises(ValueError, match="sigma"):  # CMAX_ERR_ARG -> ValueError (_lib.check)
        PatchCandidateEvaluator(torch.from_numpy(ev).cuda(), [[0, 16, 0, 21]], (16, 21), sigma=9.0).evaluate(np.zeros((1, 1, 2)))
    with pytest.raises(ValueError):
        PatchCandidateEvaluator(torch.from_numpy(ev).cuda(), [[0, 16, 0, 21]], (16, 21)).evaluate(np.zeros((2, 1, 2)))


@pytest.mark.gpu
def test_mixin_initialiser_against_the_reference_method():
    """`B200CostMixin.initialize_guess_from_optuna_sampling` next to the UNMODIFIED reference method
    (src/solver/patch_contrast_pyramid.py:320-362), both driven by the same seeded stand-in for Optuna: same candidates, losses
    to 1e-5, same best guess per patch (up to near-ties of the two best trials)."""
    from unittest import mock

    import yaml
    from oracle import reference_loader as RL
    R = RL.load()
    if R is None:
        pytest.skip("the reference is neither installed under baseline/_ref nor mounted at /root/reference")
    from event_based_optical_flow_b200.solver import B200CostMixin
    cfg = yaml.safe_load(open(os.path.join(R.root, "configs", "mvsec_indoor_no_timeaware.yaml")))
    shape = (cfg["data"]["height"], cfg["data"]["width"])
    base = R.solver.PyramidalPatchContrastMaximization

    class B200Pyramidal(B200CostMixin, base):
        pass

    with mock.patch("torch.cuda.is_available", return_value=False):
        ref = base(shape, {}, cfg["solver"], cfg["optimizer"], cfg["output"], None)
    fast = B200Pyramidal(shape, {}, cfg["solver"], cfg["optimizer"], cfg["output"], None)
    _, ev = _golden()
    rng = np.random.default_rng(21)
    pyramid_module = sys.modules[base.__module__]
    for scale in (2, 3):
        ref.overload_patch_configuration(scale)
        fast.overload_patch_configuration(scale)
        motion0 = rng.uniform(-6, 6, 2 * ref.n_patch)
        ref_optuna, my_optuna = fake_optuna(), fake_optuna()
        with mock.patch.object(pyramid_module, "optuna", ref_optuna):
            expect = ref.initialize_guess_from_optuna_sampling(ev.copy(), motion0.copy())
        with mock.patch.dict(sys.modules, {"optuna": my_optuna}):
            mine = fast.initialize_guess_from_optuna_sampling(ev.copy(), motion0.copy())
        assert expect.shape == mine.shape == (2, ref.n_patch)
        assert len(ref_optuna.studies) == len(my_optuna.studies) > 0
        worst = 0.0
        for a, b in zip(ref_optuna.studies, my_optuna.studies):
            assert len(a.trials) == len(b.trials) == int(np.ceil(40 / (scale - ref.coarest_scale)))
            for (la, pa), (lb, pb) in zip(a.trials, b.trials):
                assert pa == pb  # the same candidate sequence
                worst = max(worst, abs(la - lb) / max(abs(la), 1e-12))
        print(f"scale {scale}: {len(a.trials)} trials x {len(my_optuna.studies)} patches, max rel loss difference {worst:.2e}")
        assert worst <= 1e-5
        same = np.all(mine == expect, axis=0)  # (a near-tie of a study's two best trials may resolve the other way in fp32)
        assert same.mean() >= 0.95, same.mean()
