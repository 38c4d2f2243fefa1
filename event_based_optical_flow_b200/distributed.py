"""Event sharding for the multi-GPU contrast-maximization path.

One process per GPU (torchrun).  The time-sorted event array is cut into contiguous slices, one per rank, resident
for the whole optimize(); the flow is replicated.  Per CM iteration there are exactly two exchanges, both a
sum-all-reduce over NVLink (NCCL): the partial IWE stack after K1 and the partial motion gradient after K3; the cost
kernels run redundantly on every rank on the identical reduced IWE, so all ranks see bit-identical cost/gradient and
an SPMD scipy loop stays in lock-step.  The only one-time exchange is the global (t_min, t_max): reference time,
normalisation period and voxel bin edges must come from the WHOLE batch (src/warp.py:217-224, 254-258, 342-345).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of rank's contiguous slice; slices differ in length by at most one event."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} of {world_size}")
    base, extra = divmod(int(n), world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_events(events: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    b, e = shard_bounds(events.shape[0], world_size, rank)
    return events[b:e]


def global_time_range(events_shard: torch.Tensor, group=None) -> Tuple[float, float]:
    """(t_min, t_max) over all ranks' shards, as fp32-representable python floats.  One host sync, once per optimize()."""
    t = events_shard[:, 2].detach().to(torch.float32)
    if t.numel() > 0:
        mm = torch.stack([t.min(), -t.max()])
    else:
        mm = torch.tensor([float("inf"), float("inf")], dtype=torch.float32, device=t.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(mm, op=dist.ReduceOp.MIN, group=group)
    lo, hi = float(mm[0]), -float(mm[1])
    return lo, hi


def make_sharded_objective(events_shard: torch.Tensor, image_size, group=None, **kw):
    """ContrastObjective over this rank's shard with the global time range and the two exchanges wired in
    (`exchange="nccl"` or `"peer"`, see ContrastObjective)."""
    from .objective import ContrastObjective
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    t_range = global_time_range(events_shard, group)
    pg = None
    if world > 1:
        pg = group if group is not None else dist.group.WORLD
    return ContrastObjective(events_shard, image_size, t_range=t_range, process_group=pg, **kw)
