"""Golden vectors for the per-patch 2-dof candidate cost of the Optuna initialiser: the UNMODIFIED reference solver class
(`PyramidalPatchContrastMaximization` built from configs/mvsec_indoor_no_timeaware.yaml) runs its own `objective_initial`
(src/solver/patch_contrast_pyramid.py:364-415) -- crop, origin shift, numpy 2-dof warp, numpy bilinear vote,
scipy.ndimage.gaussian_filter, cv2.Sobel, normalised gradient magnitude -- with a stand-in `trial` that hands back preset
(trans_x, trans_y) values.  Run in the build container only:

    python tests/golden/make_golden_patch_init.py   ->  tests/golden/reference_patch_init.npz
"""
import os
import sys
from unittest import mock

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import reference_loader as RL  # noqa: E402


class PresetTrial:
    number = 0

    def __init__(self, tx, ty):
        self.values = {"trans_x": float(tx), "trans_y": float(ty)}

    def suggest_uniform(self, key, low, high):
        return self.values[key]


def structured_events(rng, n, shape, velocity=(35.0, -22.0), tmax=0.05):
    """Points of a random texture moving with one velocity (px/s scaled so that the displacement over the batch is a few
    pixels) + uniform noise: the candidate cost then has a real optimum, unlike uniform events."""
    H, W = shape
    k = n * 7 // 10
    anchors = np.stack([rng.uniform(0, H, 600), rng.uniform(0, W, 600)], 1)
    pick = rng.integers(0, len(anchors), k)
    t = rng.uniform(0, tmax, n)
    pos = anchors[pick] + np.stack([velocity[0], velocity[1]])[None] * t[:k, None] / tmax * 0.2 + rng.normal(0, 0.4, (k, 2))
    noise = np.stack([rng.uniform(0, H, n - k), rng.uniform(0, W, n - k)], 1)
    xy = np.floor(np.concatenate([pos, noise]))
    keep = (xy[:, 0] >= 0) & (xy[:, 0] < H) & (xy[:, 1] >= 0) & (xy[:, 1] < W)
    ev = np.stack([xy[:, 0], xy[:, 1], t, rng.integers(0, 2, n).astype(np.float64)], 1)[keep]
    return ev[np.argsort(ev[:, 2], kind="stable")]


def main():
    R = RL.load()
    cfg = yaml.safe_load(open(os.path.join(R.root, "configs", "mvsec_indoor_no_timeaware.yaml")))
    shape = (cfg["data"]["height"], cfg["data"]["width"])
    with mock.patch("torch.cuda.is_available", return_value=False):
        slv = R.solver.PyramidalPatchContrastMaximization(shape, {}, cfg["solver"], cfg["optimizer"], cfg["output"], None)
    rng = np.random.default_rng(17)
    ev = structured_events(rng, 24000, shape)
    out = {"shape": np.array(shape), "events_xy": ev[:, :2].astype(np.int16), "events_t": ev[:, 2], "events_p": ev[:, 3].astype(np.int8),
           "sigma": np.array(float(slv.iwe_config["blur_sigma"])), "padding": np.array(int(slv.padding)),
           "scales": np.array(list(range(1, slv.patch_scales)))}
    K = 5
    for scale in range(1, slv.patch_scales):
        slv.overload_patch_configuration(scale)
        P = slv.n_patch
        rects = np.array([[slv.patches[i].x_min, slv.patches[i].x_max, slv.patches[i].y_min, slv.patches[i].y_max] for i in range(P)])
        cand = rng.uniform(-25, 25, (P, K, 2))
        cand[:, 0] = 0.0  # the zero candidate: warped == original, loss exactly 1
        loss = np.full((P, K), np.nan)
        count = np.zeros(P, dtype=np.int64)
        chosen = list(range(P)) if P <= 16 else sorted(rng.choice(P, 16, replace=False).tolist())
        for i in chosen:
            f = R.utils.crop_event(ev, *rects[i])
            f = R.utils.set_event_origin_to_zero(np.copy(f), rects[i][0], rects[i][2], 0)
            count[i] = len(f)
            if len(f) <= 10:  # the reference skips such patches (pyramid.py:338)
                continue
            for k in range(K):
                loss[i, k] = slv.objective_initial(PresetTrial(*cand[i, k]), f, np.zeros(2))
        out[f"{scale}/patch_image_size"] = np.array(slv.scaled_patch_size[scale])
        out[f"{scale}/rects"] = rects
        out[f"{scale}/candidates"] = cand
        out[f"{scale}/loss"] = loss
        out[f"{scale}/count"] = count
        print(f"scale {scale}: {P} patches of {slv.scaled_patch_size[scale]}, {len(chosen)} evaluated, loss range "
              f"{np.nanmin(loss):.4f} .. {np.nanmax(loss):.4f}")
    np.savez_compressed(os.path.join(HERE, "reference_patch_init.npz"), **out)
    print("wrote reference_patch_init.npz")


if __name__ == "__main__":
    main()
