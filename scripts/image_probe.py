#!/usr/bin/env python
"""One-off probe (library lib/libcmax_b200_measure.so: python -m event_based_optical_flow_b200._build --measure): per-CTA globaltimer stamps of the image kernel inside the graph-replayed CM iteration."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from event_based_optical_flow_b200 import _lib as _L
_L.use_measure_library()
from event_based_optical_flow_b200 import ContrastObjective, _lib
dev = torch.device("cuda:0")
n = int(os.environ.get("PROBE_N", 5_000_000))
ev = torch.from_numpy(bench.synth_events(n, 0)).to(dev)
flows = torch.from_numpy(bench.synth_flows(4, 100)).to(dev)
obj = ContrastObjective(ev, (bench.H, bench.W), cost="image_variance")
cost = torch.zeros(1, dtype=torch.float64, device=dev); grad = torch.zeros(2, bench.H, bench.W, device=dev); fb = flows[0].clone()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev); flush_rd = torch.zeros(128 << 20, dtype=torch.int32, device=dev)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): obj.step_into(fb, cost, grad)
torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g): obj.step_into(fb, cost, grad)
lib = _lib.load()
OFF_SLOTS = int(_L.load().cmax_objective_probe_offset(obj.plan.handle)) - 2048 * 8
# slots offset inside the workspace: find via layout knowledge (off_slots) -> expose through a tiny search: stamps are the only non-zero u64 > 1e15 there
ws = obj._ws
res = []
for it in range(8):
    fb.copy_(flows[it % 4]); flush.zero_(); flush_rd.sum()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    base = obj._ws_ptr - ws.data_ptr()
    st = ws[base + OFF_SLOTS + 2048 * 8: base + OFF_SLOTS + 2048 * 8 + 148 * 64].view(torch.int64).cpu().numpy().reshape(148, 8)[:, :7].astype(np.float64)
    t0 = st[:, 0].min()
    st = (st - t0) / 1e3
    res.append((a.elapsed_time(b) * 1e3, st))
for ms, st in res[3:]:
    print(f"step {ms:.1f} us | stamps (us since first CTA start; min/median/max over CTAs):")
    for i, name in enumerate(["start", "after pdl_wait", "after fold loop", "after commit", "after grid barrier", "after slot-reduce+combine", "end"]):
        print(f"   {name:28s} {st[:, i].min():7.2f} {np.median(st[:, i]):7.2f} {st[:, i].max():7.2f}")
ms, st = res[-1]
print("per-CTA (blockIdx: start, pdl, fold, commit, barrier, combine, end)")
for b in list(range(0, 40)) + list(range(100, 148, 6)):
    print(b, " ".join(f"{v:6.2f}" for v in st[b]))
