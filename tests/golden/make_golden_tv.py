"""Golden vectors for the `total_variation` cost: the UNMODIFIED reference class (src/costs/total_variation.py) on seeded
patch-grid flows.  Run in the build container only:  python tests/golden/make_golden_tv.py  -> tests/golden/reference_tv.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _import_reference  # noqa: E402


def main():
    R = _import_reference()
    rng = np.random.default_rng(5)
    out = {}
    shapes = [(1, 1), (2, 2), (3, 3), (4, 4), (8, 8), (16, 16), (5, 9), (16, 21)]
    out["shapes"] = np.array(shapes)
    for k, (h, w) in enumerate(shapes):
        flow = rng.uniform(-20, 20, (2, h, w))
        out[f"{k}/flow"] = flow
        for direction in ("minimize", "maximize"):
            for omit in (True, False):
                for prec, dt in (("32", torch.float32), ("64", torch.float64)):
                    cost = R.costs.TotalVariation(direction=direction, store_history=False, cuda_available=False, precision=prec)
                    f = torch.from_numpy(flow).to(dt).requires_grad_(True)
                    val = cost.calculate({"flow": f, "omit_boundary": omit})
                    (g,) = torch.autograd.grad(val, f)
                    out[f"{k}/{direction}/{int(omit)}/{prec}/value"] = val.detach().numpy()
                    out[f"{k}/{direction}/{int(omit)}/{prec}/grad"] = g.numpy()
        # the batched form [b,2,h,w] the hybrid cost may receive
        fb = torch.from_numpy(np.stack([flow, -0.5 * flow])).double()
        cost = R.costs.TotalVariation(direction="minimize", store_history=False, cuda_available=False, precision="64")
        out[f"{k}/batched"] = cost.calculate({"flow": fb, "omit_boundary": True}).numpy()
    np.savez_compressed(os.path.join(HERE, "reference_tv.npz"), **out)
    print("wrote reference_tv.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
