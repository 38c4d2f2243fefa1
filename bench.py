#!/usr/bin/env python
"""Benchmark of the contrast-maximization inner loop (BASELINE.json metric): events/s per CM iteration
(warp + IWE + cost + gradient w.r.t. the motion).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2]      the B200 path (one rank per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]               the reference algorithm on the host cores

A step = one CM iteration over the resident event batch with a fresh motion.  Workloads (BASELINE.json `configs`):
  c2 (default, the configuration the metric is quoted on): 5 M synthetic events per GPU, 260x346 dense flow, variance cost
     + gradient; for N > 1 every rank holds its own 5 M-event contiguous shard (weak scaling) and the partial IWE / gradient
     are summed across the ranks each step (--exchange: inside the kernels over NVLink peer memory behind flags [default],
     or NCCL all-reduces between the stages);
  c1: 30 k events, 260x346, 2-dof translation, variance (the shipped YAML's batch size; launch-bound regime);
  c3: 10 M events, 480x640, 16x16 tile flow -> Burgers flow voxel (T=10) -> time-aware warp, gradient-magnitude cost;
  c4: 10 M events per GPU, 260x346, 16x16 tile flow -> Burgers voxel (T=10), the shipped multi-focal normalised
      gradient-magnitude cost with blur (configs/mvsec_indoor_burgers.yaml), sharded like c2.
`value` is timed on the device (CUDA events, L2 flushed before every step, the step replayed from a CUDA graph); `e2e` is the
same step through the public API with the motion coming from pinned host memory and gradient + cost going back every step;
`roofline` times each event kernel alone, live, against MEASURED_PEAKS.json; `parity` compares the benched evaluation with the
CPU oracle; `plan_ms` is the one-time cost of making the batch resident (sort + packing), `value_amortised_50_iters` charges
it to 50 CM iterations (config 2's "50 CM iters"); `sharded_vs_single` (N > 1) re-evaluates the gathered batch on one GPU.
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

H, W = 260, 346          # config 2 (kept as module constants: scripts/ import them)
EVENTS_PER_GPU = 5_000_000
MAX_FLOW = 10.0          # px of displacement over the normalised window
N_FLOWS = 8              # distinct pre-generated motions cycled through the steps
L2_FLUSH_BYTES = 512 << 20
UNIT = "events/s"

CONFIGS = {
    "c1": dict(H=260, W=346, events=30_000, model="2d-translation", cost="image_variance", sigma=0.0,
               workload="config1: 30k events, 260x346, 2-dof translation, variance cost+grad"),
    "c2": dict(H=260, W=346, events=5_000_000, model="dense-flow", cost="image_variance", sigma=0.0,
               workload="config2: 5M events per GPU, 260x346 dense flow, variance cost+grad"),
    "c3": dict(H=480, W=640, events=10_000_000, model="time-aware", cost="gradient_magnitude", sigma=0.0, T=10, window=(30, 40),
               workload="config3: 10M events, 480x640, 16x16 tile flow -> Burgers voxel T=10 -> time-aware warp, gradient-magnitude cost+grad"),
    "c4": dict(H=260, W=346, events=10_000_000, model="time-aware", cost="multi_focal_normalized_gradient_magnitude", sigma=1.0, T=10,
               window=(16, 21),
               workload="config4: 10M events per GPU, 260x346, 16x16 tile flow -> Burgers voxel T=10, multi-focal normalised "
                        "gradient magnitude (3 reference times, blur sigma 1: configs/mvsec_indoor_burgers.yaml), sharded"),
}


def metric_name(cfg) -> str:
    return f"events/sec per CM iteration (warp+IWE+cost+grad) @{cfg['W']}x{cfg['H']}"


def synth_events(n: int, seed: int, h: int = H, w: int = W) -> np.ndarray:
    """The reference's own fixture (src/utils/event_utils.py:18-47), seeded: integer pixel coordinates, sorted
    uniform timestamps in [0, 0.05), random polarity; fp32 [n,4] = (x=row, y=col, t, p)."""
    rng = np.random.default_rng(seed)
    ev = np.empty((n, 4), dtype=np.float32)
    ev[:, 0] = rng.integers(0, h, n)
    ev[:, 1] = rng.integers(0, w, n)
    ev[:, 2] = np.sort(rng.uniform(0.0, 0.05, n))
    ev[:, 3] = rng.integers(0, 2, n)
    return ev


def synth_flows(k: int, seed: int, h: int = H, w: int = W) -> np.ndarray:
    """Smooth flows as the pyramid produces at its finest scale: a 16x16 patch grid, bilinearly up-sampled
    (SURVEY.md section 8d), |flow| <= MAX_FLOW."""
    import torch
    rng = np.random.default_rng(seed)
    grid = torch.from_numpy(rng.uniform(-MAX_FLOW, MAX_FLOW, (k, 2, 16, 16)).astype(np.float32))
    return torch.nn.functional.interpolate(grid, size=(h, w), mode="bilinear", align_corners=False).numpy()


def synth_motions(cfg, k: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if cfg["model"] == "dense-flow":
        return synth_flows(k, seed, cfg["H"], cfg["W"])
    if cfg["model"] == "2d-translation":
        return rng.uniform(-20, 20, (k, 2)).astype(np.float32)
    return rng.uniform(-MAX_FLOW, MAX_FLOW, (k, 2, 16, 16)).astype(np.float32)  # tile motion of the time-aware configurations


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the event kernels, from the committed `ncu --set full`
# captures under profiles/ (config 2 only; other configurations report traffic null)
TRAFFIC_SOURCE = "profiles/r02_ncu_r2k.txt: dram__bytes_read.sum + dram__bytes_write.sum per launch"
_TRAFFIC = {"K1 vote (vote_strips_kernel)": 26.09e6, "K3 grad (grad_strips_kernel)": 26.81e6}


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


# ------------------------------------------------------------------------------------------------ the CPU oracle leg
def oracle_value_and_grad(cfg, ev, motion, dtype):
    """One CM evaluation of configuration `cfg` with the CPU oracle (oracle/cm_oracle.py = the reference's torch branch,
    pinned to the reference's own outputs by tests/test_oracle_golden.py).  -> (cost, gradient w.r.t. `motion`)."""
    import torch
    from oracle import cm_oracle as O
    ev = ev.to(dtype)
    m = motion.to(dtype)
    size = (cfg["H"], cfg["W"])
    if cfg["model"] != "time-aware":
        return O.objective_value_and_grad(ev, m, size, motion_model=cfg["model"], cost=cfg["cost"], sigma=cfg["sigma"])
    m = m.detach().clone().requires_grad_(True)
    dense = O.upsample_tile_flow(m, size, cfg["window"], cfg["window"], (0, 0))
    voxel = O.flow_voxel(dense, cfg["T"], "burgers", "middle")
    value = O.objective(ev, voxel, size, motion_model="dense-flow-voxel", cost=cfg["cost"], sigma=cfg["sigma"])
    (grad,) = torch.autograd.grad(value, m)
    return value.detach(), grad


def reference_value_and_grad(R, cfg, ev, motion):
    """One CM evaluation with the UNMODIFIED reference classes (baseline/_ref, see oracle/reference_loader.py), composed the
    way its solver composes them (src/solver/patch_contrast_base.py:289-352): Warp.warp_event -> EventImageConverter.create_iwe ->
    cost.calculate -> torch.autograd.grad.  Configurations 1 and 2 (one warp, variance)."""
    import torch
    size = (cfg["H"], cfg["W"])
    warper = R.warp.Warp(size, calculate_feature=False, normalize_t=True)
    imager = R.event_image_converter.EventImageConverter(size)
    cost = R.costs.functions[cfg["cost"]](direction="minimize", store_history=False)
    m = motion.detach().clone().requires_grad_(True)
    warped, _ = warper.warp_event(ev, m, cfg["model"], direction="first")
    iwe = imager.create_iwe(warped, "bilinear_vote", cfg["sigma"])
    loss = cost.calculate({"iwe": iwe, "omit_boundary": True})
    (grad,) = torch.autograd.grad(loss, m)
    return loss.detach(), grad


def patch_init_leg(dev):
    """The pyramid's per-patch initialiser at the shipped YAML's regime (30 000 events on 260x346, finest level: 16x16 patches of
    16x21 pixels, one TPE trial of every patch per call = 256 candidate costs): device time and end-to-end time of one batched
    call, next to the reference's own chain (src/solver/patch_contrast_pyramid.py:379-415: numpy warp / vote, scipy Gaussian,
    cv2 Sobel) timed on a sample of the same evaluations.  An auxiliary leg: reported, never part of `value`."""
    import torch
    from event_based_optical_flow_b200.patch_init import PatchCandidateEvaluator
    from oracle import patch_init_oracle as PO
    rng = np.random.default_rng(3)
    n, size, grid = 30_000, (16, 21), (16, 16)
    ev = synth_events(n, 77).astype(np.float64)
    rects = np.array([[i * size[0], (i + 1) * size[0], j * size[1], (j + 1) * size[1]] for i in range(grid[0]) for j in range(grid[1])])
    ev_dev = torch.from_numpy(ev).to(dev)
    prepare = []
    for _ in range(2):  # the first construction also pays torch's lazy loading of the operators it uses
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        evaluator = PatchCandidateEvaluator(ev_dev, rects, size, sigma=1.0)
        torch.cuda.synchronize()
        prepare.append((time.perf_counter() - t0) * 1e3)
    prepare_first_ms, prepare_ms = prepare
    cand = rng.uniform(-20, 20, (len(rects), 1, 2))
    loss = evaluator.evaluate(cand)
    cand_dev = torch.from_numpy(cand).to(dev)
    evaluator.losses(cand_dev)
    torch.cuda.synchronize()
    dev_us = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        evaluator.losses(cand_dev)
        b.record()
        torch.cuda.synchronize()
        dev_us.append(a.elapsed_time(b) * 1e3)
    t0 = time.perf_counter()
    for _ in range(20):
        evaluator.evaluate(cand)
    e2e_us = (time.perf_counter() - t0) / 20 * 1e6
    # the CPU side: the unmodified reference operators composed as calculate_cost_for_small_patch composes them, else the port
    sample = [int(i) for i in np.nonzero(evaluator.valid)[0][:48]]
    kind, R = "port", None
    try:
        from oracle import reference_loader
        R = reference_loader.load()
    except Exception:
        R = None
    if R is not None:
        kind = "reference"
        warper = R.warp.Warp(size, calculate_feature=False, normalize_t=True)
        imager = R.event_image_converter.EventImageConverter(size, outer_padding=0)
        cost = R.costs.NormalizedGradientMagnitude(direction="minimize", store_history=False, precision="64", cuda_available=False)

        def cpu_loss(f, c):
            theta_c = np.array(c) * (f[:, 2].max() - f[:, 2].min())
            warped, _ = warper.warp_event(f, theta_c, "2d-translation", direction="middle")
            return cost.calculate({"omit_boundary": False, "clip": True, "orig_iwe": imager.create_iwe(f, "bilinear_vote", 1),
                                   "iwe": imager.create_iwe(warped, "bilinear_vote", 1)})
    else:
        def cpu_loss(f, c):
            return PO.candidate_loss(f, c, size, (0, 0), 1.0)
    crops = {i: PO.crop_to_patch(ev, *rects[i]) for i in sample}
    cpu_loss(crops[sample[0]], cand[sample[0], 0])
    t0 = time.perf_counter()
    ref = np.array([float(cpu_loss(crops[i], cand[i, 0])) for i in sample])
    cpu_us = (time.perf_counter() - t0) / len(sample) * 1e6
    rel = float(np.max(np.abs(loss[sample, 0] - ref) / np.abs(ref)))
    return {"workload": f"{n} events, 260x346, {len(rects)} patches of {size[0]}x{size[1]}, 1 candidate per patch and call, sigma 1",
            "evaluations_per_call": int(len(rects)), "prepare_first_ms": prepare_first_ms, "prepare_ms": prepare_ms, "device_us_per_call": float(np.median(dev_us)),
            "e2e_us_per_call": e2e_us, "cpu_us_per_evaluation": cpu_us, "cpu_kind": kind, "cpu_sample": len(sample),
            "speedup_e2e_vs_cpu": cpu_us * len(rects) / e2e_us, "max_rel_vs_cpu": rel, "gpu_launches_per_call": 1}


def load_reference_for(cfg):
    """The unmodified reference, when it is installed and this configuration is one it is composed for here; else None (the
    oracle port times instead)."""
    if cfg["model"] == "time-aware":
        return None
    try:
        from oracle import reference_loader
        return reference_loader.load()
    except Exception:
        return None


def cpu_reference_steps(cfg, ev_np: np.ndarray, motions_np: np.ndarray, steps: int, warmup: int):
    """fp32, all host threads.  Returns (seconds per step list, threads, kind): kind "reference" = the unmodified reference's own
    classes, "port" = the oracle restatement (time-aware configurations, or no baseline/_ref)."""
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    ev = torch.from_numpy(ev_np)
    R = load_reference_for(cfg)
    times = []
    for i in range(warmup + steps):
        m = torch.from_numpy(motions_np[i % len(motions_np)])
        t0 = time.perf_counter()
        val, _ = reference_value_and_grad(R, cfg, ev, m) if R is not None else oracle_value_and_grad(cfg, ev, m, torch.float32)
        float(val)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, threads, ("reference" if R is not None else "port")


def reference_sample(cfg) -> int:
    """Events of one CPU step: the whole batch where the oracle finishes in seconds, else a bounded sample."""
    return min(cfg["events"], 5_000_000 if cfg["model"] != "time-aware" else 1_000_000)


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    n = reference_sample(cfg)
    ev = synth_events(n, 0, cfg["H"], cfg["W"])
    motions = synth_motions(cfg, N_FLOWS, seed=100)
    times, threads, kind = cpu_reference_steps(cfg, ev, motions, args.steps, args.warmup)
    sec = float(np.mean(times))
    value = n / sec
    whole = "the full batch" if n == cfg["events"] else f"a bounded sample of the {cfg['events']}-event batch"
    sample = (f"{n} events per step ({whole}), fp32, {threads} threads; " +
              ("the unmodified reference classes (Warp -> EventImageConverter -> cost -> autograd) from baseline/_ref" if kind == "reference"
               else "the oracle restatement of the reference (oracle/cm_oracle.py)"))
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["workload"], "events": n, "image": [cfg["H"], cfg["W"]]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def build_objective(cfg, ev, group, t_range, args):
    """-> (callable step_into(motion, cost, grad), the ContrastObjective, shape of the motion)."""
    from event_based_optical_flow_b200 import ContrastObjective, TimeAwareObjective
    size = (cfg["H"], cfg["W"])
    if cfg["model"] != "time-aware":
        obj = ContrastObjective(ev, size, cost=cfg["cost"], motion_model=cfg["model"], sigma=cfg["sigma"], order=args.order,
                                process_group=group, t_range=t_range, exchange=args.exchange)
        return obj.step_into, obj, obj.motion_shape
    obj = ContrastObjective(ev, size, cost=cfg["cost"], motion_model="dense-flow-voxel", n_bins=cfg["T"], sigma=cfg["sigma"],
                            order=args.order, process_group=group, t_range=t_range, exchange=args.exchange, orig_events=ev)
    tobj = TimeAwareObjective(obj, scheme="burgers", t0_location="middle",
                              tile=dict(patch_size=cfg["window"], sliding_window=cfg["window"], patch_shift=(0, 0)))
    return tobj.step_into, obj, (2, 16, 16)


def algorithmic_bytes(cfg, n, k_ref):
    """per CM iteration, and per launch of K1 / K3."""
    HW = cfg["H"] * cfg["W"]
    if cfg["model"] == "time-aware":
        T = cfg["T"]
        return {"step": 32 * n + (16 * T + 16 * k_ref) * HW, "K1": 16 * n + (8 * T + 4 * k_ref) * HW, "K3": 16 * n + (16 * T + 4 * k_ref) * HW}
    if cfg["model"] == "2d-translation":
        return {"step": 32 * n + 16 * k_ref * HW, "K1": 16 * n + 4 * k_ref * HW, "K3": 16 * n + 4 * k_ref * HW}
    return {"step": 32 * n + (16 + 16 * k_ref) * HW, "K1": 16 * n + (8 + 4 * k_ref) * HW, "K3": 16 * n + (16 + 4 * k_ref) * HW}


def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    cfg = CONFIGS[args.config]
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

    from event_based_optical_flow_b200 import EventPlan, _lib
    from event_based_optical_flow_b200.distributed import global_time_range
    _lib.load()

    Hc, Wc_ = cfg["H"], cfg["W"]
    n = cfg["events"]
    ev_np = synth_events(n, seed=rank, h=Hc, w=Wc_)   # rank r's contiguous shard of the N*n-event stream
    ev_np[:, 2] = (ev_np[:, 2] + 0.05 * rank)        # shards are consecutive in time
    motions_np = synth_motions(cfg, N_FLOWS, seed=100)  # identical on every rank (the motion is replicated)
    ev = torch.from_numpy(ev_np).to(dev)
    motions = torch.from_numpy(motions_np).to(dev)
    group = dist.group.WORLD if world > 1 else None
    t_range = global_time_range(ev, group)
    reshard_ms = None
    if world > 1 and args.shard == "pixel":
        # every rank was handed a time slice (what a streaming source delivers); one all-to-all turns the slices into
        # contiguous slices of the pixel-ordered stream, which makes both per-iteration exchanges local (distributed.py)
        from event_based_optical_flow_b200.distributed import reshard_events_by_pixel
        reshard_events_by_pixel(ev, (Hc, Wc_), group)  # warm-up (NCCL channel set-up)
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ev = reshard_events_by_pixel(ev, (Hc, Wc_), group)
        torch.cuda.synchronize()
        reshard_ms = (time.perf_counter() - t0) * 1e3
        n_local = int(ev.shape[0])
    else:
        n_local = n

    # ---- one-time cost of making the batch resident (validation, sort, strips): host wall clock around plan creation,
    # which ends with the plan's single stream synchronisation
    plan_times = []
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        probe_plan = EventPlan(ev, (Hc, Wc_), 0, args.order, t_range)
        torch.cuda.synchronize()
        plan_times.append((time.perf_counter() - t0) * 1e3)
        probe_plan.close()
        del probe_plan
    plan_ms = float(np.min(plan_times[1:]))

    step_into, obj, motion_shape = build_objective(cfg, ev, group, t_range, args)
    if args.vote_variant >= 0 or args.grad_variant >= 0:  # default: what the plan chose (strip kernels when the batch qualifies)
        obj.plan.set_variant(args.vote_variant if args.vote_variant >= 0 else 5, args.grad_variant if args.grad_variant >= 0 else 5)
    compact = obj.plan.set_compact(not args.no_compact)
    strips = obj.plan.n_strips > 0 and args.vote_variant in (-1, 5) and args.grad_variant in (-1, 5)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    flush_rd = torch.zeros(L2_FLUSH_BYTES // 4, dtype=torch.int32, device=dev)

    def flush_l2():
        """Write a buffer 4x the L2 (evicts everything), then read another one of the same size so that the L2 is left
        full of CLEAN lines: a write-only flush leaves ~126 MB of dirty lines whose write-back would be charged to the
        step that follows."""
        flush.zero_()
        flush_rd.sum()
    n_motion = int(np.prod(motion_shape))
    cost_buf = torch.zeros(1, dtype=torch.float64, device=dev)
    grad_buf = torch.zeros(motion_shape, dtype=torch.float32, device=dev)
    motion_buf = torch.zeros(motion_shape, dtype=torch.float32, device=dev)

    # one CM iteration, captured once in a CUDA graph (sharded: the flag exchange lives inside the kernels, NCCL is capturable)
    graph = None
    if not args.no_graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                step_into(motion_buf, cost_buf, grad_buf)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step_into(motion_buf, cost_buf, grad_buf)

    def align():
        """Sharded runs: an in-stream cross-GPU barrier between the L2 flush and the start event.  The flush (1 GiB of memory
        traffic per rank, ~250 us) is not part of a step, but the ranks drift apart by several microseconds while each runs
        its own; without re-aligning them the first exchange of the step would be charged that drift."""
        if world > 1 and getattr(obj, "_symm", None) is not None:
            obj._symm.barrier(channel=0)

    def run_steps(count: int, first: int):
        """-> per-step device milliseconds (CUDA events on the launching stream), L2 flushed before every step."""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(count)]
        for k in range(count):
            motion_buf.copy_(motions[(first + k) % N_FLOWS])  # outside the timed bracket: the motion is "already resident"
            flush_l2()
            align()
            evs[k][0].record()
            if graph is not None:
                graph.replay()
            else:
                step_into(motion_buf, cost_buf, grad_buf)
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
    host_motions = [torch.from_numpy(motions_np[i]).pin_memory() for i in range(N_FLOWS)]
    # one device block [gradient | cost] so that the result goes back to the host in ONE copy
    n_pad = (n_motion + 1) // 2 * 2
    out_dev = torch.zeros(n_pad + 2, dtype=torch.float32, device=dev)
    out_host = torch.empty(n_pad + 2, dtype=torch.float32).pin_memory()
    e2e_grad = out_dev[:n_motion].view(motion_shape)
    e2e_cost = out_dev[n_pad:].view(torch.float64)
    e2e_motion = torch.zeros(motion_shape, dtype=torch.float32, device=dev)

    def e2e_steps(count: int):
        """Host wall-clock of `count` end-to-end steps.  The L2 flush (hygiene, not part of a step) is enqueued and waited
        for BEFORE the clock of each step starts, so no flush time has to be estimated and subtracted."""
        total = 0.0
        for k in range(count):
            flush_l2()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_motion.copy_(host_motions[k % N_FLOWS], non_blocking=True)   # H2D from pinned host memory
            step_into(e2e_motion, e2e_cost, e2e_grad)                        # the allocation-free public entry point
            out_host.copy_(out_dev, non_blocking=True)                       # D2H: gradient + cost
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
    h2d = int(host_motions[0].numel() * 4)
    d2h = int(out_host.numel() * 4)

    # ---- Hessian-vector product (what Newton-CG / trust-* of the shipped YAMLs ask for per inner CG step): device time per call
    hvp_ms = None
    if world == 1 and cfg["model"] != "time-aware" and not args.skip_cpu:
        try:
            vec = torch.randn(motion_shape, dtype=torch.float32, device=dev)
            obj.hvp(motions[0], vec)
            torch.cuda.synchronize()
            t_h = []
            for k in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                obj.hvp(motions[k % N_FLOWS], vec)
                b.record()
                torch.cuda.synchronize()
                t_h.append(a.elapsed_time(b))
            hvp_ms = float(np.mean(t_h))
        except Exception as e:  # report, never fail the benchmark line over the second-order leg
            hvp_ms = f"failed: {e!r}"

    # ---- the pyramid's per-patch initialiser, batched: one call = one TPE trial of all 256 patches
    patch_init = None
    if world == 1 and args.config == "c2" and not args.skip_cpu:
        try:
            patch_init = patch_init_leg(dev)
        except Exception as e:
            patch_init = f"failed: {e!r}"

    # ---- the same batch driven by the patch grid the pyramid optimises (16x16 nodes, the shipped YAML's geometry): the tile-flow
    # model evaluates the dense flow inside the event kernels, so the per-step host traffic is 2 KB each way instead of 720 KB
    tile_leg = None
    if world == 1 and args.config == "c2" and obj.plan.n_strips > 0:
        from event_based_optical_flow_b200 import TileFlowObjective
        tm_np = np.random.default_rng(7).uniform(-MAX_FLOW, MAX_FLOW, (N_FLOWS, 2, 16, 16)).astype(np.float32)
        tm_host = [torch.from_numpy(tm_np[i]).pin_memory() for i in range(N_FLOWS)]
        tm_dev = torch.zeros(2, 16, 16, dtype=torch.float32, device=dev)
        t_out = torch.zeros(514, dtype=torch.float32, device=dev)
        t_out_host = torch.empty(514, dtype=torch.float32).pin_memory()
        t_grad, t_cost = t_out[:512].view(2, 16, 16), t_out[512:].view(torch.float64)
        tile_leg = {"what": "same batch, motion = 16x16 patch grid (the shipped YAML's geometry); 'fused' = the tile-flow model (the event kernels "
                            "evaluate interpolate(motion) themselves, no dense flow / gradient), 'composed' = up-sampling kernel + dense model + adjoint kernel"}
        for tag, fused in (("fused", True), ("composed", False)):
            tile = TileFlowObjective(obj, patch_size=(16, 21), sliding_window=(16, 21), patch_shift=(2, 5), t_scale=1.0, fused=fused)
            for _ in range(3):
                tile.step_into(tm_dev, t_cost, t_grad)
            tg = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(tg):
                tile.step_into(tm_dev, t_cost, t_grad)
            dev_ms = []
            for k in range(args.warmup + args.steps):
                tm_dev.copy_(tm_host[k % N_FLOWS])
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                tg.replay()
                b.record()
                torch.cuda.synchronize()
                if k >= args.warmup:
                    dev_ms.append(a.elapsed_time(b))
            wall = 0.0
            for k in range(args.warmup + args.steps):
                flush_l2()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tm_dev.copy_(tm_host[k % N_FLOWS], non_blocking=True)
                tg.replay()  # (the public step_into, captured: what a solver with a per-shape graph cache replays)
                t_out_host.copy_(t_out, non_blocking=True)
                torch.cuda.synchronize()
                if k >= args.warmup:
                    wall += time.perf_counter() - t0
            tile_leg[tag] = {"ms_per_step": float(np.mean(dev_ms)), "value": n / (float(np.mean(dev_ms)) * 1e-3),
                             "e2e": {"value": n * args.steps / wall, "unit": UNIT, "h2d_bytes_per_step": 2048, "d2h_bytes_per_step": 2056}}
            tg.reset()

    # ---- sharded vs single GPU: rank 0 gathers every shard, evaluates the whole batch on its own GPU and compares
    sharded_check = None
    if world > 1:
        motion_buf.copy_(motions[0])
        step_into(motion_buf, cost_buf, grad_buf)
        torch.cuda.synchronize()
        grads = [torch.empty_like(grad_buf) for _ in range(world)]
        costs = [torch.empty_like(cost_buf) for _ in range(world)]
        dist.all_gather(grads, grad_buf)
        dist.all_gather(costs, cost_buf)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([ev.shape[0]], dtype=torch.int64, device=dev))
        n_max = int(max(int(c) for c in counts))
        padded = torch.zeros((n_max, 4), dtype=ev.dtype, device=dev)
        padded[:ev.shape[0]] = ev
        shards = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, shards, dst=0)
        if rank == 0:
            shards = [sh[:int(c)] for sh, c in zip(shards, counts)]
        if rank == 0:
            args_single = argparse.Namespace(**{**vars(args), "exchange": "nccl"})
            full_ev = torch.cat(shards)
            del shards
            single_step, single_obj, _ = build_objective(cfg, full_ev, None, None, args_single)
            c1 = torch.zeros_like(cost_buf)
            g1 = torch.zeros_like(grad_buf)
            single_step(motion_buf, c1, g1)
            torch.cuda.synchronize()
            sharded_check = {
                "cost_rel": abs(float(cost_buf) - float(c1)) / abs(float(c1)),
                "grad_rel": float(torch.linalg.norm(grad_buf.double() - g1.double()) / torch.linalg.norm(g1.double())),
                "all_ranks_bit_equal": bool(all(torch.equal(grads[0], g) for g in grads) and all(torch.equal(costs[0], c) for c in costs)),
                "tolerance": 1e-5,
                "what": f"rank 0 gathered the {world} shards ({world * n} events) and evaluated them on one GPU with the same motion",
            }
            del single_obj, full_ev
        dist.barrier()

    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy, burst)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        k_ref = len(obj.directions)
        ab = algorithmic_bytes(cfg, n, k_ref)
        kernels = {}
        stage_ms = None
        if world == 1:
            # ---- roofline of the two event kernels, each timed alone, live, with CUDA events: the staged entry points enqueue
            # exactly the kernels of the fused call (vote = K1 alone, grad(pre_zeroed) = K3 alone)
            import ctypes as C
            L = _lib
            model = L.MOTION[obj.motion_model]
            if cfg["model"] == "time-aware":
                from event_based_optical_flow_b200 import ops
                pad = ops.tile_flow_geometry((Hc, Wc_), cfg["window"], cfg["window"], (0, 0))
                m = ops.flow_voxel(ops.tile_flow_upsample(motions[1], (Hc, Wc_), pad, cfg["window"]), cfg["T"], "burgers", "middle")
            else:
                m = motions[1].contiguous()
            gk = torch.zeros(obj.motion_shape, dtype=torch.float32, device=dev)
            obj.value_and_grad(m)  # leave a consistent workspace behind
            stream = torch.cuda.current_stream().cuda_stream
            spec_p = C.byref(obj.spec)
            orig = obj._orig_stat.data_ptr() if obj._orig_stat is not None else None

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
                L.call("cmax_objective_cost", obj.plan.handle, spec_p, orig, obj._ws_ptr, 1, cost_buf.data_ptr(), gk.data_ptr(), gk.numel(), stream)

            def vote():
                L.call("cmax_objective_vote", obj.plan.handle, model, m.data_ptr(), obj._ws_ptr, stream)

            def grad():
                L.call("cmax_objective_grad", obj.plan.handle, model, m.data_ptr(), obj._ws_ptr, gk.data_ptr(), 1, stream)

            k1 = time_kernel(vote)
            fold_and_cost()
            k3 = time_kernel(grad)
            # in-situ stage times of one eager iteration (L2 flushed before the iteration only): K1 | fold + cost | K3
            stage = np.zeros(3)
            for rep in range(13):
                flush_l2()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                e[0].record()
                vote()
                e[1].record()
                fold_and_cost()
                e[2].record()
                grad()
                e[3].record()
                torch.cuda.synchronize()
                if rep >= 3:
                    stage += [e[i].elapsed_time(e[i + 1]) for i in range(3)]
            stage /= 10
            stage_ms = {"vote(K1)": stage[0], "fold+cost(image kernel launches, staged API)": stage[1], "grad(K3)": stage[2]}
            kind = "strips" if strips else "runs"
            kernels = {f"K1 vote (vote_{kind}_kernel)": {"ms": k1, "bytes": ab["K1"]}, f"K3 grad (grad_{kind}_kernel)": {"ms": k3, "bytes": ab["K3"]}}
            for v in kernels.values():
                v["GBps"] = v["bytes"] / (v["ms"] * 1e-3) / 1e9
        step_gbps = ab["step"] / (ms_per_step * 1e-3) / 1e9
        if kernels:
            dom = max(kernels, key=lambda k: kernels[k]["ms"])
            roof = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["GBps"] / peak, "traffic": _TRAFFIC.get(dom) if args.config == "c2" else None,
                    "traffic_source": TRAFFIC_SOURCE if args.config == "c2" else None, "peak_source": peak_src,
                    "kernels": kernels, "stages_in_situ_ms": stage_ms,
                    "step": {"bytes": ab["step"], "achieved": step_gbps, "frac": step_gbps / peak}}
        else:
            roof = {"bound": "hbm", "kernel": "whole CM iteration (per GPU)", "achieved": step_gbps, "peak": peak, "unit": "GB/s",
                    "frac": step_gbps / peak, "traffic": None, "peak_source": peak_src}

        # ---- parity with the CPU oracle on the benched batch (time-aware configurations: a bounded sample) and the CPU baseline
        cpu = parity = None
        if world == 1 and not args.skip_cpu:
            n_cpu = reference_sample(cfg)
            ev_cpu = synth_events(n_cpu, 0, Hc, Wc_) if n_cpu != n else ev_np
            m0 = torch.from_numpy(motions_np[0])
            pobj = None
            if n_cpu != n:
                pstep, pobj, _ = build_objective(cfg, torch.from_numpy(ev_cpu).to(dev), None, None, args)
            else:
                pstep = step_into
            c_gpu = torch.zeros(1, dtype=torch.float64, device=dev)
            g_gpu = torch.zeros(motion_shape, dtype=torch.float32, device=dev)
            pstep(m0.to(dev), c_gpu, g_gpu)
            torch.cuda.synchronize()
            del pobj
            torch.set_num_threads(os.cpu_count() or 1)
            v64, g64 = oracle_value_and_grad(cfg, torch.from_numpy(ev_cpu), m0, torch.float64)
            v32, g32 = oracle_value_and_grad(cfg, torch.from_numpy(ev_cpu), m0, torch.float32)
            parity = {
                "cost_rel_vs_fp64_oracle": abs(float(c_gpu) - float(v64)) / abs(float(v64)),
                "grad_rel_vs_fp32_oracle": float(torch.linalg.norm(g_gpu.cpu().double() - g32.double()) / torch.linalg.norm(g32.double())),
                "grad_rel_vs_fp64_oracle": float(torch.linalg.norm(g_gpu.cpu().double() - g64) / torch.linalg.norm(g64)),
                "fp32_oracle_grad_rel_vs_fp64_oracle": float(torch.linalg.norm(g32.double() - g64) / torch.linalg.norm(g64)),
                "tolerance": 1e-5, "events": n_cpu,
                "what": "cost vs the fp64 oracle, gradient norm-wise vs the same-dtype (fp32) oracle; the fp64 rows show how far "
                        "fp32 itself is from fp64 (the gradient is discontinuous where a floor index flips)",
            }
            times, threads, kind = cpu_reference_steps(cfg, ev_cpu, motions_np, steps=3 if cfg["model"] == "time-aware" else 5, warmup=1)
            cpu = {"value": n_cpu / float(np.mean(times)), "unit": UNIT, "cores": threads, "kind": kind,
                   "sample": f"{len(times)} timed + 1 warm-up CM iterations over {n_cpu} events "
                             f"({'the full batch' if n_cpu == n else 'a bounded sample of the batch'}), fp32, "
                             + ("the unmodified reference classes from baseline/_ref" if kind == "reference" else "the oracle restatement")}
            if kind == "reference":  # the parity object names what it was checked against; add the live reference next to the oracle
                vr, gr = reference_value_and_grad(load_reference_for(cfg), cfg, torch.from_numpy(ev_cpu), m0)
                parity["cost_rel_vs_fp32_reference"] = abs(float(c_gpu) - float(vr)) / abs(float(vr))
                parity["grad_rel_vs_fp32_reference"] = float(torch.linalg.norm(g_gpu.cpu().double() - gr.double()) / torch.linalg.norm(gr.double()))

        clocks = sampler.finish() if sampler else None
        if cfg["model"] == "time-aware":
            per_step_kernels = None  # tile upsample, voxel propagation, K1, image-side chain, K3, the two adjoints: see profiles/
        else:
            # K1 vote, image kernel (fold + variance + cost + gradient quads), K3 grad; sharded "peer": + the gradient-exchange
            # kernel (the IWE exchange lives inside the image kernel); "nccl": K1, fold, cost, K3 (+ 2 NCCL all-reduces)
            per_step_kernels = (3 if world == 1 else 4) + (1 if cfg["model"] == "2d-translation" else 0)
        if reshard_ms is not None:
            plan_ms += reshard_ms  # the one-time all-to-all belongs to making the batch resident
        amortised_ms = ms_per_step + plan_ms / 50.0
        line = {
            "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "events_per_gpu": n, "image": [Hc, Wc_],
                       "motion": {"dense-flow": "smooth dense flow (16x16 grid upsampled), |f|<=10px, fresh per step",
                                  "2d-translation": "2-dof translation, |theta|<=20px, fresh per step",
                                  "time-aware": "16x16 tile motion, |f|<=10px, fresh per step"}[cfg["model"]],
                       "event_order": args.order, "packed_event_bytes": 4.5 if strips else (8 if compact else 16),
                       "vote_variant": args.vote_variant, "grad_variant": args.grad_variant, "cuda_graph": graph is not None,
                       "l2": f"flushed before every timed step ({L2_FLUSH_BYTES >> 20} MiB written, then {L2_FLUSH_BYTES >> 20} MiB read so no dirty lines remain)"
                             + ("; ranks re-aligned by an in-stream barrier between the flush and the start event" if world > 1 and args.exchange == "peer" else ""),
                       "parallelism": (f"events sharded x{world} (" + ("time slices re-distributed once into contiguous slices of the pixel-ordered stream"
                                                                      if args.shard == "pixel" else "contiguous time slices") + "), sum(IWE)+sum(grad) per step via " +
                                       {"nccl": "NCCL all-reduce",
                                        "peer": "NVLink peer-memory reads behind in-kernel flags (no collective, no barrier kernel)"}[args.exchange])
                       if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "host pinned motion -> device, step_into (public API), gradient+cost -> pinned host in one copy, sync; events resident"},
            "gpu_launches": per_step_kernels * args.steps if per_step_kernels else None,
            "plan_ms": plan_ms, "reshard_ms": reshard_ms, "events_this_rank": n_local, "value_amortised_50_iters": world * n / (amortised_ms * 1e-3),
            "hvp_ms": hvp_ms, "patch_init": patch_init, "tile_flow": tile_leg, "roofline": roof, "parity": parity, "sharded_vs_single": sharded_check, "cpu_baseline": cpu, "clocks": clocks,
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
    ap.add_argument("--config", choices=sorted(CONFIGS), default="c2", help="BASELINE.json configuration (default: c2, the one the metric is quoted on)")
    ap.add_argument("--order", choices=("asis", "tile", "pixel"), default="pixel")
    ap.add_argument("--vote-variant", type=int, default=-1, help="-1 = the plan's choice (5 = strip kernels when the batch qualifies, else 2)")
    ap.add_argument("--grad-variant", type=int, default=-1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-compact", action="store_true", help="force the 16-byte packed-event format")
    ap.add_argument("--exchange", choices=("nccl", "peer"), default="peer",
                    help="multi-GPU: NCCL all-reduces between the stages, or NVLink peer reads behind flags inside the kernels")
    ap.add_argument("--shard", choices=("pixel", "time"), default="pixel",
                    help="multi-GPU: 'time' keeps the contiguous time slices every rank is handed; 'pixel' (default) re-distributes them once "
                         "(one all-to-all at plan time) into contiguous slices of the pixel-ordered stream, which keeps both exchanges local")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the parity / cpu_baseline legs (used under ncu)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
