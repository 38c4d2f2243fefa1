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


def pixel_sort_key(events: torch.Tensor, image_size) -> torch.Tensor:
    """The plan's sort key of every event (tile-major: 32x32 tiles of un-warped pixels, row-major inside a tile; cmax_events.cu
    keys_kernel), int64 [n]."""
    H, W = int(image_size[0]), int(image_size[1])
    r = events[:, 0].detach().long().clamp_(0, H - 1)
    c = events[:, 1].detach().long().clamp_(0, W - 1)
    tiles_x = (W + 31) // 32
    return ((r // 32) * tiles_x + c // 32) * 1024 + (r % 32) * 32 + (c % 32)


def reshard_events_by_pixel(events_shard: torch.Tensor, image_size, group=None) -> torch.Tensor:
    """Re-distribute time-sliced shards so that every rank holds a CONTIGUOUS SLICE OF THE PIXEL-ORDERED STREAM (equal event
    counts up to one pixel's worth): one all-to-all, once per optimize().

    Why: a rank's partial IWE is then non-zero only in the band of rows its source pixels can be warped into, and its partial
    gradient only in the rows of its own source pixels.  The sharded kernels publish those row ranges with their flags and
    every reader skips the peers whose rows miss its pixels (cmax_objective_sharded), so the two per-iteration exchanges move
    the overlaps between neighbouring bands instead of N whole images per rank -- their cost stops growing with the number
    of GPUs.  Results do not depend on how the events are distributed (both sums are exact over any partition).
    Events of one pixel keep their time order (source shards are consecutive in time and arrive in rank order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return events_shard
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    H, W = int(image_size[0]), int(image_size[1])
    n_keys = ((H + 31) // 32) * ((W + 31) // 32) * 1024
    key = pixel_sort_key(events_shard, image_size)
    hist = torch.bincount(key, minlength=n_keys)
    dist.all_reduce(hist, group=group)
    cum = torch.cumsum(hist, 0)
    total = int(cum[-1])
    # rank q takes the keys whose cumulative count ends in (q * total / world, (q + 1) * total / world]
    targets = torch.tensor([(q * total) // world for q in range(1, world)], dtype=cum.dtype, device=cum.device)
    bounds = torch.searchsorted(cum, targets, right=False)  # first key index of rank q+1 = bounds[q] + 1
    dest = torch.searchsorted(bounds, key, right=False)      # key <= bounds[q] -> rank <= q
    sorted_dest, order = torch.sort(dest.to(torch.uint8), stable=True)  # (8-bit keys: one radix pass instead of eight)
    send = events_shard.detach()[order].contiguous()
    # counts per destination from the SORTED destinations (a bincount of 5 M values into 8 bins is 5 M atomics onto 8 addresses)
    edges = torch.searchsorted(sorted_dest.to(torch.int32), torch.arange(world, device=dest.device, dtype=torch.int32), right=False)
    edges = torch.cat([edges, edges.new_tensor([sorted_dest.numel()])])
    send_counts = edges[1:] - edges[:-1]
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_list, recv_list = send_counts.tolist(), recv_counts.tolist()
    recv = events_shard.new_empty((sum(recv_list), events_shard.shape[1]))
    dist.all_to_all_single(recv, send, output_split_sizes=recv_list, input_split_sizes=send_list, group=group)
    return recv


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
