"""2-GPU parity of the sharded objective (needs >= 2 CUDA devices; skipped otherwise): contiguous event shards + the
global time range + the two per-iteration exchanges (NCCL all-reduces between the stages; NVLink peer-memory reads behind
flags inside the image / gradient-exchange kernels) must reproduce the single-GPU cost and gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import event_based_optical_flow_b200 as B
        from event_based_optical_flow_b200.distributed import make_sharded_objective, reshard_events_by_pixel, shard_events
        rng = np.random.default_rng(7)
        H, W, n = 96, 128, 400_001
        ev = np.stack([rng.integers(0, H, n), rng.integers(0, W, n), np.sort(rng.uniform(0, 0.05, n)), rng.integers(0, 2, n)], 1)
        ev = torch.from_numpy(ev.astype(np.float32)).to(dev)
        flow = torch.from_numpy(rng.uniform(-6, 6, (2, H, W)).astype(np.float32)).to(dev)
        results = {}
        for cost, sigma in (("image_variance", 0.0), ("multi_focal_normalized_gradient_magnitude", 1.0)):
            full = B.ContrastObjective(ev, (H, W), cost=cost, motion_model="dense-flow", sigma=sigma)
            v_ref, g_ref = full.value_and_grad(flow)
            for exchange, reshard in (("nccl", False), ("peer", False), ("peer", True)):
                mine = shard_events(ev, world, rank)
                if reshard:  # contiguous slices of the pixel-ordered stream: the exchanges skip the rows a peer cannot have touched
                    mine = reshard_events_by_pixel(mine, (H, W))
                obj = make_sharded_objective(mine, (H, W), cost=cost, motion_model="dense-flow", sigma=sigma,
                                             exchange=exchange, orig_events=mine)
                for _ in range(3):  # repeated evaluations exercise the buffer-reuse hazards of the peer exchange
                    v, g = obj.value_and_grad(flow)
                v_only = obj.value(flow)
                torch.cuda.synchronize()
                rel_v = abs(float(v) - float(v_ref)) / abs(float(v_ref))
                rel_g = float(torch.linalg.norm(g - g_ref) / torch.linalg.norm(g_ref))
                results[(cost, exchange, reshard)] = (rel_v, rel_g, abs(float(v_only) - float(v)) / abs(float(v)))
                # every rank must hold the identical result (SPMD optimisers stay in lock-step)
                both = [torch.zeros_like(g) for _ in range(world)]
                dist.all_gather(both, g)
                assert all(torch.equal(both[0], b) for b in both), (cost, exchange, reshard)
        # the tile-flow model on a sharded objective: the gradient that crosses NVLink is 2 * hp * wp floats
        grid, window = (6, 8), (16, 16)
        motion = torch.from_numpy(rng.uniform(-5, 5, (2,) + grid).astype(np.float32)).to(dev)
        full = B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow")
        v_ref, g_ref = B.TileFlowObjective(full, window, window, (0, 0), 0.9, fused=False).value_and_grad(motion)
        for exchange in ("nccl", "peer"):
            mine = reshard_events_by_pixel(shard_events(ev, world, rank), (H, W))
            obj = make_sharded_objective(mine, (H, W), cost="image_variance", motion_model="dense-flow", exchange=exchange)
            tile = B.TileFlowObjective(obj, window, window, (0, 0), 0.9)
            assert tile.fused, "sharded + strips -> the fused tile-flow model is the default"
            for _ in range(3):
                v, g = tile.value_and_grad(motion)
            torch.cuda.synchronize()
            rel_v = abs(float(v) - float(v_ref)) / abs(float(v_ref))
            rel_g = float(torch.linalg.norm(g - g_ref) / torch.linalg.norm(g_ref))
            results[("tile-flow", exchange, True)] = (rel_v, rel_g, 0.0)
            both = [torch.zeros_like(g) for _ in range(world)]
            dist.all_gather(both, g)
            assert all(torch.equal(both[0], b) for b in both), ("tile-flow", exchange)
        # second order on a sharded objective (a collective call): H v of the single-GPU objective, through `hvp` and through
        # the path scipy_autograd takes (functional.vhp over obj(m)); ops.SumPartials / ops.UseReplicated carry the sums
        vec = torch.from_numpy(rng.standard_normal((2, H, W)).astype(np.float32)).to(dev)
        for cost, sigma in (("image_variance", 0.0), ("multi_focal_normalized_gradient_magnitude", 1.0)):
            full = B.ContrastObjective(ev, (H, W), cost=cost, motion_model="dense-flow", sigma=sigma)
            hv_ref = full.hvp(flow, vec)
            mine = reshard_events_by_pixel(shard_events(ev, world, rank), (H, W))
            obj = make_sharded_objective(mine, (H, W), cost=cost, motion_model="dense-flow", sigma=sigma, exchange="peer")
            hv = obj.hvp(flow, vec)
            _, hv_vhp = torch.autograd.functional.vhp(lambda m: obj(m), flow.double(), vec.double())
            rel = lambda a, b: float(torch.linalg.norm(a.double() - b.double()) / torch.linalg.norm(b.double()))  # noqa: E731
            results[("hvp", cost, "hvp")] = (0.0, rel(hv, hv_ref), 0.0)
            results[("hvp", cost, "vhp")] = (0.0, rel(hv_vhp, hv_ref), 0.0)
            both = [torch.zeros_like(hv) for _ in range(world)]
            dist.all_gather(both, hv)
            assert all(torch.equal(both[0], b) for b in both), ("hvp", cost)
        pm = torch.from_numpy(rng.uniform(-5, 5, (2,) + grid)).to(dev)
        pv = torch.from_numpy(rng.standard_normal((2,) + grid)).to(dev)
        full = B.ContrastObjective(ev, (H, W), cost="image_variance", motion_model="dense-flow")
        _, hv_ref = torch.autograd.functional.vhp(lambda m: B.TileFlowObjective(full, window, window, (0, 0), 0.9, fused=False)(m), pm, pv)
        obj = make_sharded_objective(reshard_events_by_pixel(shard_events(ev, world, rank), (H, W)), (H, W), cost="image_variance",
                                     motion_model="dense-flow", exchange="peer")
        _, hv = torch.autograd.functional.vhp(lambda m: B.TileFlowObjective(obj, window, window, (0, 0), 0.9)(m), pm, pv)
        results[("hvp", "tile-flow", "vhp")] = (0.0, float(torch.linalg.norm(hv - hv_ref) / torch.linalg.norm(hv_ref)), 0.0)
        out[rank] = results
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_objective_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world
        for rank in range(world):
            for key, (rel_v, rel_g, rel_vo) in out[rank].items():
                assert rel_v <= 1e-5, (rank, key, rel_v)
                print(rank, key, rel_v, rel_g)
                assert rel_g <= (2e-5 if key[0] == "hvp" else 1e-5), (rank, key, rel_g)  # (H v: fp32 second differences of two summation orders)
                assert rel_vo <= 1e-6, (rank, key, rel_vo)
