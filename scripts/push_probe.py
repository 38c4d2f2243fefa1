#!/usr/bin/env python
"""Micro-probe of the push exchange primitives on N GPUs (CUDA-event timed, eager launches, rank 0 prints)."""
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
H, W, n = bench.H, bench.W, 200_000
ev = torch.from_numpy(bench.synth_events(n, seed=rank)).to(dev)
obj = ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow", process_group=dist.group.WORLD,
                        t_range=global_time_range(ev, dist.group.WORLD), exchange="push")
stream = torch.cuda.current_stream().cuda_stream
w = obj._push_words.data_ptr()
src = obj._grad_part.data_ptr()
one = lambda arr: (C.c_void_p * 1)(arr[rank])  # noqa: E731


def t(fn, reps=20):
    out = []
    for i in range(reps + 3):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) * 1e3)
    return round(float(np.mean(out[3:])), 2)


res = {}
res["empty_kernel(wait on satisfied flags)"] = t(lambda: _lib.call("cmax_reduce_peers", obj._peer_grad, world, 0, None, obj._flags_grad_ptr, w + 60, stream))
res["push n=0 (fence+flags to all)"] = t(lambda: _lib.call("cmax_push", None, 0, obj._push_grad_slots, obj._push_grad_flags, world, w + 4, w + 12, stream))
res["push 720KB to self only"] = t(lambda: _lib.call("cmax_push", src, obj._n_motion, one(obj._push_grad_slots), one(obj._push_grad_flags), 1, w + 20, w + 24, stream))
res["push 720KB to all"] = t(lambda: _lib.call("cmax_push", src, obj._n_motion, obj._push_grad_slots, obj._push_grad_flags, world, w + 4, w + 12, stream))
res["push 360KB to all"] = t(lambda: _lib.call("cmax_push", obj._iwe_local_ptr, obj._n_iwe, obj._push_iwe_slots, obj._push_iwe_flags, world, w, w + 8, stream))


def push_wait():
    _lib.call("cmax_push", src, obj._n_motion, obj._push_grad_slots, obj._push_grad_flags, world, w + 4, w + 12, stream)
    _lib.call("cmax_reduce_peers", obj._peer_grad, world, 0, None, obj._flags_grad_ptr, w + 4, stream)


res["push 720KB + wait"] = t(push_wait)
res["symm barrier"] = t(lambda: obj._symm.barrier(channel=0))
x = torch.zeros(180_000, device=dev)
res["nccl allreduce 720KB"] = t(lambda: dist.all_reduce(x))
if rank == 0:
    print(res)
dist.barrier()
dist.destroy_process_group()
