#!/usr/bin/env python
"""Stage timings of the sharded CM iteration (exchange='peer'), eager launches, CUDA events on rank 0.
   torchrun --nproc-per-node N scripts/mgpu_breakdown.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from event_based_optical_flow_b200 import ContrastObjective, _lib  # noqa: E402
from event_based_optical_flow_b200.distributed import global_time_range  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
H, W, n = bench.H, bench.W, bench.EVENTS_PER_GPU
ev = bench.synth_events(n, seed=rank)
ev[:, 2] += 0.05 * rank
ev = torch.from_numpy(ev).to(dev)
flows = torch.from_numpy(bench.synth_flows(4, seed=100)).to(dev)
obj = ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow", process_group=dist.group.WORLD,
                        t_range=global_time_range(ev, dist.group.WORLD), exchange=os.environ.get("EXCHANGE", "push"))
cost = torch.zeros(1, dtype=torch.float64, device=dev)
grad = torch.zeros(2, H, W, device=dev)
stream = torch.cuda.current_stream().cuda_stream
names = ["vote(K1+fold)", "barrier0|push_iwe", "reduce_iwe(+cost)", "cost_after_reduce(gq)", "grad(K3)", "barrier1|push_grad", "reduce_grad"]
MODE = os.environ.get("EXCHANGE", "push")
acc = np.zeros(len(names))
reps = 30
for it in range(reps + 5):
    m = flows[it % 4].contiguous()
    dist.barrier()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    e[0].record()
    obj._vote(m, stream); e[1].record()
    w = obj._push_words.data_ptr() if MODE == "push" else 0
    if MODE == "push":
        _lib.call("cmax_push", obj._iwe_local_ptr, obj._n_iwe, obj._push_iwe_slots, obj._push_iwe_flags, obj._n_peers, w, w + 8, stream)
    else:
        obj._symm.barrier(channel=0)
    e[2].record()
    combined = C.c_int32(0)
    _lib.call("cmax_objective_reduce_iwe", obj.plan.handle, C.byref(obj.spec), obj._peer_iwe, obj._n_peers, None, obj._ws_ptr, cost.data_ptr(),
              C.byref(combined), obj._flags_iwe_ptr if MODE == "push" else None, w if MODE == "push" else None, stream); e[3].record()
    _lib.call("cmax_objective_cost_after_reduce", obj.plan.handle, C.byref(obj.spec), None, obj._ws_ptr, combined.value, 1, cost.data_ptr(), stream); e[4].record()
    _lib.call("cmax_objective_grad", obj.plan.handle, 0, m.data_ptr(), obj._ws_ptr, obj._grad_part.data_ptr(), stream); e[5].record()
    if MODE == "push":
        _lib.call("cmax_push", obj._grad_part.data_ptr(), obj._n_motion, obj._push_grad_slots, obj._push_grad_flags, obj._n_peers, w + 4, w + 12, stream)
    else:
        obj._symm.barrier(channel=1)
    e[6].record()
    _lib.call("cmax_reduce_peers", obj._peer_grad, obj._n_peers, grad.numel(), grad.data_ptr(), obj._flags_grad_ptr if MODE == "push" else None,
              (w + 4) if MODE == "push" else None, stream); e[7].record()
    torch.cuda.synchronize()
    if it >= 5:
        acc += [e[i].elapsed_time(e[i + 1]) * 1e3 for i in range(len(names))]
if rank == 0:
    print({k: round(v / reps, 2) for k, v in zip(names, acc)}, "us; total", round(acc.sum() / reps, 2))
dist.barrier()
dist.destroy_process_group()
