#!/usr/bin/env python
"""Secondary measurements for profiles/ (NOT the bench.py line): the other single-GPU BASELINE configurations through the
public objectives, CUDA-event timed per stage, L2 flushed before every evaluation.

  config 1: 30 k events, 260x346, 2-dof translation, variance                         (ContrastObjective)
  config 3: 10 M events, 480x640, 16x16 tile flow -> dense -> 10-bin Burgers voxel -> voxel warp, gradient magnitude
            (TimeAwareObjective: every stage a kernel of this library)
  config 3m: the same with the shipped multi-focal normalised gradient-magnitude cost (3 reference times, blur sigma 1)

Prints one JSON line per configuration.  usage: python scripts/bench_configs.py [--events-c3 N] [--iters K]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import event_based_optical_flow_b200 as B  # noqa: E402
from event_based_optical_flow_b200 import ops  # noqa: E402


def events(n, H, W, seed):
    rng = np.random.default_rng(seed)
    ev = np.empty((n, 4), dtype=np.float32)
    ev[:, 0] = rng.integers(0, H, n)
    ev[:, 1] = rng.integers(0, W, n)
    ev[:, 2] = np.sort(rng.uniform(0.0, 0.05, n))
    ev[:, 3] = rng.integers(0, 2, n)
    return torch.from_numpy(ev)


def timed(fn, iters, flush):
    out = []
    for i in range(iters + 3):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(i)
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b))
    return float(np.mean(out[3:])), float(np.min(out[3:]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--events-c3", type=int, default=10_000_000)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    fl = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    fr = torch.zeros((512 << 20) // 4, dtype=torch.int32, device=dev)

    def flush():
        fl.zero_()
        fr.sum()

    rng = np.random.default_rng(0)
    # ---- config 1
    H, W, n = 260, 346, 30_000
    ev = events(n, H, W, 0).to(dev)
    obj = B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="2d-translation")
    thetas = torch.from_numpy(rng.uniform(-20, 20, (8, 2)).astype(np.float32)).to(dev)
    mean, best = timed(lambda i: obj.value_and_grad(thetas[i % 8]), args.iters, flush)
    print(json.dumps({"config": "c1: 30k events, 260x346, 2-dof translation, variance cost+grad", "ms_per_iteration": mean, "best_ms": best,
                      "events_per_s": n / (mean * 1e-3), "strips": bool(obj.plan.lib and obj.plan.set_compact(True))}), flush=True)
    # ---- config 3
    H, W, n, T = 480, 640, args.events_c3, 10
    ev = events(n, H, W, 1).to(dev)
    grid, window = (16, 16), (30, 40)
    motions = torch.from_numpy(rng.uniform(-10, 10, (8, 2) + grid).astype(np.float32)).to(dev)
    for tag, cost, sigma in (("c3", "gradient_magnitude", 0.0), ("c3m", "multi_focal_normalized_gradient_magnitude", 1.0)):
        t0 = time.perf_counter()
        obj = B.ContrastObjective(ev, (H, W), cost=cost, motion_model="dense-flow-voxel", n_bins=T, sigma=sigma)
        tobj = B.TimeAwareObjective(obj, scheme="burgers", t0_location="middle",
                                    tile=dict(patch_size=window, sliding_window=window, patch_shift=(0, 0)))
        torch.cuda.synchronize()
        plan_s = time.perf_counter() - t0
        mean, best = timed(lambda i: tobj.value_and_grad(motions[i % 8]), args.iters, flush)
        # stages, each timed alone
        m = motions[0]
        pad, win = tobj.tile
        dense = ops.tile_flow_upsample(m, (H, W), pad, win)
        vox = ops.flow_voxel(dense, T, "burgers", "middle")
        gv = torch.randn_like(vox)
        st = {
            "tile_upsample": timed(lambda i: ops.tile_flow_upsample(m, (H, W), pad, win), 10, flush)[0],
            "flow_voxel(burgers,T=10)": timed(lambda i: ops.flow_voxel(dense, T, "burgers", "middle"), 10, flush)[0],
            "objective(K1+cost+K3, voxel warp)": timed(lambda i: obj.value_and_grad(vox), 10, flush)[0],
            "flow_voxel_backward": timed(lambda i: ops.flow_voxel_backward(dense, vox, gv, "burgers", "middle"), 10, flush)[0],
            "tile_upsample_backward": timed(lambda i: ops.tile_flow_upsample_backward(dense, grid, pad, win), 10, flush)[0],
        }
        k = len(obj.directions)
        algo_bytes = 32 * n + (16 * T + 16 * k) * H * W  # time-aware, k reference times in one event pass
        print(json.dumps({"config": f"{tag}: {n} events, {H}x{W}, 16x16 tile flow, Burgers voxel T={T}, {cost}, sigma {sigma}",
                          "ms_per_iteration": mean, "best_ms": best, "events_per_s": n / (mean * 1e-3), "plan_create_s": plan_s,
                          "algorithmic_bytes": algo_bytes, "algorithmic_GBps": algo_bytes / (mean * 1e-3) / 1e9, "stages_ms": st}), flush=True)
        del obj, tobj


if __name__ == "__main__":
    main()
