#!/usr/bin/env python
"""Benchmark of the contrast-maximization inner loop (BASELINE.json metric): events/s per CM iteration
(warp + IWE + variance cost + gradient w.r.t. the dense flow) at 346x260.

  python bench.py [--gpus N] [--steps K] [--warmup W]            the B200 path (one rank per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]       the reference algorithm on the host cores

A step = one CM iteration over the resident event batch with a fresh flow field.  Workload = BASELINE config 2
(5 M synthetic events per GPU, 260x346 dense flow, variance cost + gradient); for N > 1 every rank holds its own
5 M-event contiguous shard (weak scaling) and the partial IWE / gradient are summed across the ranks each step
(--exchange: NVLink peer-memory kernels behind in-stream barriers by default, NCCL all-reduce or the push exchange on request).
`value` is timed on the device (CUDA events, L2 flushed before every step, the step replayed from a CUDA graph); `e2e` is the
same step through the public API with the flow coming from pinned host memory and gradient + cost going back every step;
`roofline` times each event kernel alone against MEASURED_PEAKS.json; `cpu_baseline` is the oracle port on the host cores.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W = 260, 346
EVENTS_PER_GPU = 5_000_000
MAX_FLOW = 10.0          # px of displacement over the normalised window
N_FLOWS = 8              # distinct pre-generated flow fields cycled through the steps
L2_FLUSH_BYTES = 512 << 20
METRIC = "events/sec per CM iteration (warp+IWE+cost+grad) @346x260"
UNIT = "events/s"


def synth_events(n: int, seed: int) -> np.ndarray:
    """The reference's own fixture (src/utils/event_utils.py:18-47), seeded: integer pixel coordinates, sorted
    uniform timestamps in [0, 0.05), random polarity; fp32 [n,4] = (x=row, y=col, t, p)."""
    rng = np.random.default_rng(seed)
    ev = np.empty((n, 4), dtype=np.float32)
    ev[:, 0] = rng.integers(0, H, n)
    ev[:, 1] = rng.integers(0, W, n)
    ev[:, 2] = np.sort(rng.uniform(0.0, 0.05, n))
    ev[:, 3] = rng.integers(0, 2, n)
    return ev


def synth_flows(k: int, seed: int) -> np.ndarray:
    """Smooth flows as the pyramid produces at its finest scale: a 16x16 patch grid, bilinearly up-sampled
    (SURVEY.md section 8d), |flow| <= MAX_FLOW."""
    import torch
    rng = np.random.default_rng(seed)
    grid = torch.from_numpy(rng.uniform(-MAX_FLOW, MAX_FLOW, (k, 2, 16, 16)).astype(np.float32))
    return torch.nn.functional.interpolate(grid, size=(H, W), mode="bilinear", align_corners=False).numpy()


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two event kernels at this workload, from the
# committed `ncu --set full` capture profiles/r01_ncu_r1o.txt (strip kernels) / r01_ncu_r1k.txt (run kernels)
TRAFFIC_SOURCE = "profiles/r01_ncu_r1o.txt (strips) / r01_ncu_r1k.txt (runs): dram__bytes_read.sum + dram__bytes_write.sum per launch"
_TRAFFIC = {"K1 vote (vote_strips_kernel)": 26.09e6, "K3 grad (grad_strips_kernel)": 26.81e6,
            "K1 vote (vote_runs_kernel)": 42.18e6, "K3 grad (grad_runs_kernel)": 42.91e6}


def traffic_of(kernel: str):
    return _TRAFFIC.get(kernel)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Polls SM clock + throttle reasons through NVML while the timed regions run."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            names = {
                nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
                nv.nvmlClocksThrottleReasonApplicationsClocksSetting: "applications_clocks_setting",
            }
            while not self._stop_evt.is_set():
                util = nv.nvmlDeviceGetUtilizationRates(h).gpu
                mhz = int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                if util > 0:
                    self.samples.append(mhz)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(self.period)
        except Exception as e:  # NVML missing -> report, do not fake
            self.error = repr(e)

    def finish(self) -> dict:
        self._stop_evt.set()
        self.join(timeout=2.0)
        out = {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples)}
        if self.error:
            out["error"] = self.error
        return out


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_reference_steps(ev_np: np.ndarray, flows_np: np.ndarray, steps: int, warmup: int):
    """The reference's algorithm for this path (torch CPU branch: warp -> bilinear vote -> variance -> autograd
    gradient), restated in oracle/cm_oracle.py and pinned to the reference's outputs by tests/test_oracle_golden.py.
    fp32, all host threads.  Returns (seconds per step list, threads)."""
    import torch
    from oracle import cm_oracle as O
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ev = torch.from_numpy(ev_np)
    times = []
    for i in range(warmup + steps):
        flow = torch.from_numpy(flows_np[i % len(flows_np)])
        t0 = time.perf_counter()
        val, grad = O.objective_value_and_grad(ev, flow, (H, W), motion_model="dense-flow", cost="image_variance")
        float(val)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, threads


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = EVENTS_PER_GPU
    ev = synth_events(n, seed=0)
    flows = synth_flows(N_FLOWS, seed=100)
    times, threads = cpu_reference_steps(ev, flows, args.steps, args.warmup)
    sec = float(np.mean(times))
    value = n / sec
    sample = f"{n} events (one full config-2 batch) per step, fp32, torch CPU ops, {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: 5M events, 260x346 dense flow, variance cost+grad", "events": n, "image": [H, W]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the B200 path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from event_based_optical_flow_b200 import ContrastObjective, _lib
    from event_based_optical_flow_b200.distributed import global_time_range
    _lib.load()

    n = EVENTS_PER_GPU
    ev_np = synth_events(n, seed=rank)            # rank r's contiguous shard of the N*5M-event stream
    ev_np[:, 2] = (ev_np[:, 2] + 0.05 * rank)     # shards are consecutive in time
    flows_np = synth_flows(N_FLOWS, seed=100)     # identical on every rank (the flow is replicated)
    ev = torch.from_numpy(ev_np).to(dev)
    flows = torch.from_numpy(flows_np).to(dev)
    group = dist.group.WORLD if world > 1 else None
    t_range = global_time_range(ev, group)
    obj = ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow", sigma=0.0, order=args.order,
                            process_group=group, t_range=t_range, exchange=args.exchange)
    if args.vote_variant >= 0 or args.grad_variant >= 0:  # default: what the plan chose (strip kernels when the batch qualifies)
        obj.plan.set_variant(args.vote_variant if args.vote_variant >= 0 else 5, args.grad_variant if args.grad_variant >= 0 else 5)
    compact = obj.plan.set_compact(not args.no_compact)
    strips = args.vote_variant in (-1, 5) and args.grad_variant in (-1, 5) and args.order == "pixel"  # (dense synthetic batch: the plan builds strips)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=dev)

    def flush_l2():
        """Write a buffer 4x the L2 (evicts everything), then read another one of the same size so that the L2 is left
        full of CLEAN lines: a write-only flush leaves ~126 MB of dirty lines whose write-back would be charged to the
        step that follows."""
        flush.zero_()
        flush_rd.sum()
    cost_buf = torch.zeros(1, dtype=torch.float64, device=dev)
    grad_buf = torch.zeros(2, H, W, dtype=torch.float32, device=dev)
    flow_buf = torch.zeros(2, H, W, dtype=torch.float32, device=dev)

    # one CM iteration, captured once in a CUDA graph: single GPU = 4 kernel nodes; sharded = 5 kernels + 2 NCCL
    # all-reduces (NCCL is capturable; every rank captures the same sequence)
    graph = None
    if not args.no_graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                obj.step_into(flow_buf, cost_buf, grad_buf)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            obj.step_into(flow_buf, cost_buf, grad_buf)

    def step(i: int):
        if graph is not None:
            flow_buf.copy_(flows[i % N_FLOWS])  # outside the timed bracket: the flow is "already resident"
            return None
        return flows[i % N_FLOWS]

    def run_steps(count: int, first: int):
        """-> per-step device milliseconds (CUDA events on the launching stream), L2 flushed before every step."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        for k in range(count):
            f = step(first + k)
            flush_l2()
            evs[k][0].record()
            if graph is not None:
                graph.replay()
            else:
                c, g = obj.value_and_grad(f)
            evs[k][1].record()
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    run_steps(args.warmup, 0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = run_steps(args.steps, args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = float(np.sum(ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # ---- end to end through the public API: per step the motion comes from pinned host memory (what scipy hands
    # over, scipy_autograd/torch_wrapper.py:33-36) and cost + gradient go back to the host (:46-49).  Events stay
    # resident, as in the reference (patch_contrast_pyramid.py:186 moves them once per optimize()).
    host_flows = [torch.from_numpy(flows_np[i]).pin_memory() for i in range(N_FLOWS)]
    host_grad = torch.empty(2, H, W, dtype=torch.float32).pin_memory()
    host_cost = torch.empty(1, dtype=torch.float64).pin_memory()

    # one device block [gradient | cost] so that the result goes back to the host in ONE copy
    out_dev = torch.zeros(2 * H * W + 2, dtype=torch.float32, device=dev)
    out_host = torch.empty(2 * H * W + 2, dtype=torch.float32).pin_memory()
    e2e_grad = out_dev[:2 * H * W].view(2, H, W)
    e2e_cost = out_dev[2 * H * W:].view(torch.float64)
    e2e_flow = torch.zeros(2, H, W, dtype=torch.float32, device=dev)

    def e2e_steps(count: int):
        """Host wall-clock of `count` end-to-end steps.  The L2 flush (hygiene, not part of a step) is enqueued and waited
        for BEFORE the clock of each step starts, so no flush time has to be estimated and subtracted."""
        total = 0.0
        for k in range(count):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_flow.copy_(host_flows[k % N_FLOWS], non_blocking=True)   # H2D from pinned host memory
            obj.step_into(e2e_flow, e2e_cost, e2e_grad)                  # the allocation-free public entry point
            out_host.copy_(out_dev, non_blocking=True)                   # D2H: gradient + cost
            torch.cuda.synchronize()
            total += time.perf_counter() - t0
        return total

    e2e_steps(max(3, args.warmup))
    if world > 1:
        dist.barrier()
    e2e_s = e2e_steps(args.steps)
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * n * args.steps / e2e_s
    h2d = int(host_flows[0].numel() * 4)
    d2h = int(out_host.numel() * 4)

    line = None
    if rank == 0:
        # ---- roofline of the dominant kernels, each timed alone with CUDA events (stage mask = event kernel only)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        HWp = H * W
        kernels = {}
        if world == 1:
            from event_based_optical_flow_b200 import _lib as L
            import ctypes as C
            obj.value_and_grad(flows[0])  # leave a consistent workspace behind
            stream = torch.cuda.current_stream().cuda_stream
            m = flows[1].contiguous()
            spec_p = C.byref(obj.spec)
            ng = grad_buf.numel()
            # the staged entry points enqueue exactly the kernels of the fused call: vote = K1 alone, grad(pre_zeroed) = K3 alone

            def time_kernel(fn, reps=20):
                out = []
                for _ in range(reps + 3):
                    flush_l2()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    torch.cuda.synchronize()
                    out.append(a.elapsed_time(b))
                return float(np.mean(out[3:]))

            def fold_and_cost():
                L.call("cmax_objective_fold", obj.plan.handle, obj._ws_ptr, None, stream)
                L.call("cmax_objective_cost", obj.plan.handle, spec_p, None, obj._ws_ptr, 1, cost_buf.data_ptr(), grad_buf.data_ptr(), ng, stream)

            k1 = time_kernel(lambda: L.call("cmax_objective_vote", obj.plan.handle, 0, m.data_ptr(), obj._ws_ptr, stream))
            fold_and_cost()
            k3 = time_kernel(lambda: L.call("cmax_objective_grad", obj.plan.handle, 0, m.data_ptr(), obj._ws_ptr, grad_buf.data_ptr(), 1, stream))
            # in-situ stage times of one eager iteration (L2 flushed before the iteration only): K1 | fold + cost | K3
            stage_ms = np.zeros(3)
            for rep in range(13):
                flush_l2()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e[0].record()
                L.call("cmax_objective_vote", obj.plan.handle, 0, m.data_ptr(), obj._ws_ptr, stream)
                e[1].record()
                fold_and_cost()
                e[2].record()
                L.call("cmax_objective_grad", obj.plan.handle, 0, m.data_ptr(), obj._ws_ptr, grad_buf.data_ptr(), 1, stream)
                e[3].record()
                torch.cuda.synchronize()
                if rep >= 3:
                    stage_ms += [e[i].elapsed_time(e[i + 1]) for i in range(3)]
            stage_ms /= 10
            # algorithmic bytes per launch (DESIGN.md "Kernels"): K1 = 16 B/event + flow read 8 HW + IWE write 4 HW;
            # K3 = 16 B/event + flow read 8 HW + dL/dIWE read 4 HW + gradient write 8 HW
            kernels = {"K1 vote (vote_strips_kernel)" if strips else "K1 vote (vote_runs_kernel)": {"ms": k1, "bytes": 16 * n + 12 * HWp},
                       "K3 grad (grad_strips_kernel)" if strips else "K3 grad (grad_runs_kernel)": {"ms": k3, "bytes": 16 * n + 20 * HWp}}
            for v in kernels.values():
                v["GBps"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        step_bytes = 32 * n + 32 * HWp  # per CM iteration, single reference time, dense flow
        step_gbps = step_bytes / (ms_per_step * 1e-3) / 1e9
        if kernels:
            dom = max(kernels, key=lambda k: kernels[k]["ms"])
            roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["GBps"] / peak, "traffic": traffic_of(dom), "traffic_source": TRAFFIC_SOURCE, "peak_source": peak_src,
                    "kernels": kernels, "stages_in_situ_ms": {"vote(K1)": stage_ms[0], "fold+cost(2 image-kernel launches, staged API)": stage_ms[1],
                                                              "grad(K3)": stage_ms[2]}, "step": {"bytes": step_bytes, "achieved": step_gbps, "frac": step_gbps / peak}}
        else:
            roof = {"bound": "hbm", "kernel": "whole CM iteration (per GPU)", "achieved": step_gbps, "peak": peak, "unit": "GB/s",
                    "frac": step_gbps / peak, "traffic": None, "peak_source": peak_src}

        # ---- CPU baseline: the oracle port of the reference algorithm on this box's host cores, bounded sample
        cpu = None
        if world == 1 and not args.skip_cpu:
            times, threads = cpu_reference_steps(synth_events(n, seed=0), flows_np, steps=5, warmup=1)
            cpu = {"value": n / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"5 timed + 1 warm-up CM iterations over the full {n}-event config-2 batch, fp32 torch CPU ops"}

        clocks = sampler.finish() if sampler else None
        # K1 vote, image kernel (fold + variance + cost + gradient quads), K3 grad; sharded "peer": + the gradient-exchange
        # kernel (the IWE exchange lives inside the image kernel); "nccl": K1, fold, cost, K3 (+ 2 NCCL all-reduces)
        per_step_kernels = 3 if world == 1 else {'peer': 4, 'nccl': 4}[args.exchange]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "config2: 5M events per GPU, 260x346 dense flow, variance cost+grad", "events_per_gpu": n,
                       "image": [H, W], "flow": "smooth (16x16 grid upsampled), |f|<=10px, fresh per step",
                       "event_order": args.order, "packed_event_bytes": 4.5 if strips else (8 if compact else 16), "vote_variant": args.vote_variant, "grad_variant": args.grad_variant,
                       "cuda_graph": graph is not None, "l2": f"flushed before every timed step ({L2_FLUSH_BYTES >> 20} MiB written, then {L2_FLUSH_BYTES >> 20} MiB read so no dirty lines remain)",
                       "parallelism": (f"events sharded x{world}, sum(IWE)+sum(grad) per step via " +
                                       {"nccl": "NCCL all-reduce",
                                        "peer": "NVLink peer-memory reads behind in-kernel flags (no collective, no barrier kernel)"}[args.exchange])
                       if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "host pinned flow -> device, ContrastObjective.step_into (public API), gradient+cost -> pinned host in one copy, sync; events resident"},
            "gpu_launches": per_step_kernels * args.steps,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        # a live CUDA graph that holds NCCL kernels keeps the communicator busy: release it before tearing NCCL down
        if graph is not None:
            graph.reset()
            del graph
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    if os.environ.get("BENCH_DEBUG_DUMP_AFTER"):
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BENCH_DEBUG_DUMP_AFTER"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=("b200", "reference"), default="b200")
    ap.add_argument("--order", choices=("asis", "tile", "pixel"), default="pixel")
    ap.add_argument("--vote-variant", type=int, default=-1, help="-1 = the plan's choice (5 = strip kernels when the batch qualifies, else 2)")
    ap.add_argument("--grad-variant", type=int, default=-1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="force the 16-byte packed-event format")
    ap.add_argument("--exchange", choices=("nccl", "peer"), default="peer",
                    help="multi-GPU: NCCL all-reduces between the stages, or NVLink peer reads behind flags inside the kernels")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg (used under ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
