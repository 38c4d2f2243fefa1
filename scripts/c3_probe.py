import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import event_based_optical_flow_b200 as B
from event_based_optical_flow_b200 import ops
sys.path.insert(0, os.path.join(os.getcwd(), "scripts"))
from bench_configs import events
dev = torch.device("cuda:0")
H, W, n, T = 480, 640, 10_000_000, 10
ev = events(n, H, W, 1).to(dev)
rng = np.random.default_rng(0)
obj = B.ContrastObjective(ev, (H, W), cost="gradient_magnitude", motion_model="dense-flow-voxel", n_bins=T, sigma=0.0)
g = torch.from_numpy(rng.uniform(-10, 10, (1, 2, 16, 16)).astype(np.float32))
dense = torch.nn.functional.interpolate(g, size=(H, W), mode="bilinear", align_corners=False)[0].contiguous().to(dev)
vox = ops.flow_voxel(dense, T, "burgers", "middle")
for _ in range(4):
    obj.value_and_grad(vox)
torch.cuda.synchronize()
