"""Host-side cost of one objective call through the reference's solver class + B200CostMixin at the shipped YAML's regime
(30 000 events, hybrid cost = multi-focal contrast + total variation, finest pyramid level), "eager" (`loss.item()` per term and call like
src/costs/base.py:53-56, the plugin's own total variation through autograd) against the mixin's defaults (history materialised
when read, total variation with its analytic gradient in one pass).  Needs baseline/_ref.
    python scripts/mixin_history_probe.py [calls] [eager,default,...]"""
import os
import sys
import time

import numpy as np
import torch
import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_loader as RL  # noqa: E402  (a probe script, not the product path)
from event_based_optical_flow_b200.solver import B200CostMixin  # noqa: E402


def main():
    R = RL.load()
    cfg = yaml.safe_load(open(os.path.join(R.root, "configs", "mvsec_indoor_no_timeaware.yaml")))
    shape = (cfg["data"]["height"], cfg["data"]["width"])

    class B200Pyramidal(B200CostMixin, R.solver.PyramidalPatchContrastMaximization):
        pass

    slv = B200Pyramidal(shape, {}, cfg["solver"], cfg["optimizer"], cfg["output"], None)
    rng = np.random.default_rng(0)
    n = 30_000
    ev = np.stack([rng.integers(0, shape[0], n), rng.integers(0, shape[1], n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1).astype(np.float64)
    evc = torch.from_numpy(ev).double().requires_grad_().to("cuda")
    scale = slv.patch_scales - 1
    slv.overload_patch_configuration(scale)
    slv.cost_func.enable_history_register()
    xs = [rng.uniform(-10, 10, 2 * slv.n_patch) for _ in range(8)]

    def call(x):
        m = torch.from_numpy(x).double().to("cuda").requires_grad_(True)   # scipy_autograd/torch_wrapper.py:33-36
        loss = slv.objective_scipy(m, evc, {}, True)
        (g,) = torch.autograd.grad(loss, m)
        return loss.cpu().item(), g.cpu().numpy()                           # :46-49

    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    modes = sys.argv[2] if len(sys.argv) > 2 else "eager,default,eager,default"
    for mode in modes.split(","):
        # "eager": the round-2 state before the seam clean-ups (one .item() per cost term and call, the plugin's own total
        # variation through autograd); "default": deferred history + the one-pass total variation
        slv.b200_defer_history = slv.b200_fast_total_variation = mode == "default"
        slv.cost_func.clear_history()
        for k in range(10):
            call(xs[k % 8])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(calls):
            call(xs[k % 8])
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / calls * 1e6
        hist = slv.cost_func.get_history()
        print(f"{mode}: {us:.1f} us per objective call (value + gradient, host in / host out), "
              f"history entries {[len(v) for v in hist.values()]}", flush=True)


if __name__ == "__main__":
    main()
