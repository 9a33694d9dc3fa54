"""Multi-GPU plumbing: campaigns are independent units, so they are sharded statically over the
ranks (one process per GPU) with NO data-path collective; the only collective of the path is
the final gather of the traces on rank 0.  Works with the `nccl`
backend on GPUs and with `gloo` on CPU tensors (tests/test_dist_gloo.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_campaigns: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of campaign ids owned by `rank`: campaign c -> c // ceil(B / world).
    Callers order campaigns by (model, pool, n) first, so a rank's campaigns share pool tiles in L2."""
    per = -(-num_campaigns // world_size)
    lo = min(num_campaigns, rank * per)
    return lo, min(num_campaigns, lo + per)


def shard(items: Sequence, world_size: int, rank: int) -> List:
    lo, hi = shard_bounds(len(items), world_size, rank)
    return list(items[lo:hi])


def gather_traces(trace: torch.Tensor, values: torch.Tensor, num_campaigns: int,
                  group: Optional[dist.ProcessGroup] = None) -> Optional[Tuple[torch.Tensor, torch.Tensor]]:
    """Gather per-rank traces (int32 [b_rank, T]) and values (fp64 [b_rank, T]) on rank 0 in campaign
    order.  Shards may be ragged (last rank shorter): they are padded to ceil(B / world) rows for
    the collective and trimmed afterwards.  Returns (trace [B, T], values [B, T]) on rank 0, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return trace, values
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-num_campaigns // world)
    T = trace.shape[1]
    pt = torch.full((per, T), -1, dtype=trace.dtype, device=trace.device)
    pv = torch.full((per, T), float("nan"), dtype=values.dtype, device=values.device)
    pt[: trace.shape[0]] = trace
    pv[: values.shape[0]] = values
    if rank == 0:
        gt = [torch.empty_like(pt) for _ in range(world)]
        gv = [torch.empty_like(pv) for _ in range(world)]
    else:
        gt = gv = None
    dist.gather(pt, gt, dst=0, group=group)
    dist.gather(pv, gv, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat(gt)[:num_campaigns], torch.cat(gv)[:num_campaigns]
