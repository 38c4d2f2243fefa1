"""Times torch symmetric-memory barrier, a 360 KB / 720 KB NCCL all-reduce and the peer kernels inside a CUDA graph
(run under torchrun on >= 2 GPUs).  Diagnostic for profiles/."""
import os
import sys

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
buf = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
hdl = symm_mem.rendezvous(buf, dist.group.WORLD)
a = torch.zeros(90_000, device=dev); b = torch.zeros(180_000, device=dev)


def timed(fn, reps=200):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(5):
            fn()
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps // 10):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / reps
    g.reset()
    return us


res = {
    "symm barrier": timed(lambda: hdl.barrier(channel=0)),
    "nccl allreduce 360KB": timed(lambda: dist.all_reduce(a)),
    "nccl allreduce 720KB": timed(lambda: dist.all_reduce(b)),
    "empty kernel (fill 4B)": timed(lambda: a[:1].zero_()),
}
if rank == 0:
    print({k: round(v, 2) for k, v in res.items()})
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
