#!/usr/bin/env python
"""Probe of the sharded CM iteration (lib/libcmax_b200_measure.so: python -m event_based_optical_flow_b200._build --measure): per-CTA globaltimer stamps of the image kernel and of
the gradient-exchange kernel inside the graph-replayed step.  Launch with torchrun, one rank per GPU; rank 0 prints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from event_based_optical_flow_b200 import _lib as _L
_L.use_measure_library()
from event_based_optical_flow_b200 import ContrastObjective
from event_based_optical_flow_b200.distributed import global_time_range
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"])); dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("PROBE_N", 5_000_000))
ev_np = bench.synth_events(n, rank); ev_np[:, 2] += 0.05 * rank
ev = torch.from_numpy(ev_np).to(dev)
if os.environ.get("PROBE_SHARD", "time") == "pixel":
    from event_based_optical_flow_b200.distributed import reshard_events_by_pixel
    ev = reshard_events_by_pixel(ev, (bench.H, bench.W), dist.group.WORLD)
flows = torch.from_numpy(bench.synth_flows(4, 100)).to(dev)
obj = ContrastObjective(ev, (bench.H, bench.W), cost="image_variance", process_group=dist.group.WORLD, t_range=global_time_range(ev, dist.group.WORLD), exchange="peer")
cost = torch.zeros(1, dtype=torch.float64, device=dev); grad = torch.zeros(2, bench.H, bench.W, device=dev); fb = flows[0].clone()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev); flush_rd = torch.zeros(128 << 20, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): obj.step_into(fb, cost, grad)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize(); dist.barrier()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): obj.step_into(fb, cost, grad)
OFF = int(_L.load().cmax_objective_probe_offset(obj.plan.handle))
ws = obj._ws; base = obj._ws_ptr - ws.data_ptr()
out = []
for it in range(8):
    fb.copy_(flows[it % 4]); flush.zero_(); flush_rd.sum(); torch.cuda.synchronize(); dist.barrier()
    flush.zero_(); flush_rd.sum(); obj._symm.barrier(channel=0)  # as bench.py: flush, then align the ranks in-stream
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    st = ws[base + OFF: base + OFF + (148 + 64) * 64].view(torch.int64).cpu().numpy().reshape(148 + 64, 8).astype(np.float64)
    out.append((a.elapsed_time(b) * 1e3, st))
gathered = [None] * world
dist.all_gather_object(gathered, [(ms, st.tolist()) for ms, st in out[3:]])
if rank == 0:
    for r in range(world):  # per-rank summary of the last step: K1 end (pdl), flags seen, image end, exchange start/flags/end
        ms, st = gathered[r][-1]
        st = np.array(st); img, ex = st[:148], st[148:]; ex = ex[ex[:, 0] > 0]; t0 = img[:, 0].min()
        f = lambda a: (np.median(a) - t0) / 1e3
        print(f"rank {r}: step {ms:.1f} us | pdl {f(img[:,1]):.1f} fold {f(img[:,3]):.1f} flags {f(img[:,7]):.1f} barrier {f(img[:,4]):.1f} end {f(img[:,6]):.1f} | "
              f"xchg start {f(ex[:,0]):.1f} flags {f(ex[:,1]):.1f} end {f(ex[:,2]):.1f}")
    names = ["start", "after pdl_wait (K1 done)", "after fold loop", "after fold (cta)", "after grid barrier", "after slot-reduce+combine", "end", "after wait_flags (peers' IWEs ready)"]
    order = [0, 1, 2, 3, 7, 4, 5, 6]
    for ms, st in out[3:]:
        img, ex = st[:148], st[148:]
        t0 = img[:, 0].min()
        print(f"step {ms:.1f} us | image kernel stamps (us since its first CTA start; min/median/max over CTAs):")
        for i in order:
            v = (img[:, i] - t0) / 1e3
            print(f"   {names[i]:40s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}")
        ex = ex[ex[:, 0] > 0]
        for i, nm in enumerate(["exchange kernel start", "after wait_flags (peers' gradients ready)", "end"]):
            v = (ex[:, i] - t0) / 1e3
            print(f"   {nm:40s} {v.min():7.2f} {np.median(v):7.2f} {v.max():7.2f}")
g.reset(); del g
torch.cuda.synchronize(); dist.barrier(); dist.destroy_process_group()
