"""Image sharding and the detection all-gather.

The path shards by image with no data-path collective: every rank decodes / NMSes / assigns its own
contiguous slice of the batch.  The only exchange is the evaluator's: fixed-size padded detections
`[B_loc, 300, 6]` + `[B_loc]` counts, all-gathered once (per step or per accumulated chunk) over
NCCL / NVLink.  The reference has no multi-GPU path at all (train.py:33, devices=1).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous image range [lo, hi) of `rank`; earlier ranks take the remainder."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank: Optional[int] = None, world: Optional[int] = None):
    """Slices every tensor (or list of tensors) along dim 0 to this rank's images."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    def cut(t):
        lo, hi = shard_range(t.shape[0], rank, world)
        return t[lo:hi]
    if isinstance(tensors, (list, tuple)):
        return type(tensors)(cut(t) for t in tensors)
    return cut(tensors)


def gather_detections(dets: torch.Tensor, counts: torch.Tensor, batch: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gathers per-rank padded detections into the global batch order.
    dets [B_loc, max_det, 6], counts [B_loc] -> ([B, max_det, 6], [B]).  Ranks may hold unequal
    shards (batch not divisible by world): shards are padded to the largest and trimmed after."""
    world = dist.get_world_size(group)
    if world == 1:
        return dets, counts
    b_loc = dets.shape[0]
    if batch is None:
        sizes = [torch.zeros(1, dtype=torch.int64, device=dets.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([b_loc], dtype=torch.int64, device=dets.device), group=group)
        sizes = [int(s) for s in sizes]
    else:
        sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    m = max(sizes)
    if b_loc < m:
        dets = torch.cat([dets, dets.new_zeros((m - b_loc,) + dets.shape[1:])])
        counts = torch.cat([counts, counts.new_zeros(m - b_loc)])
    all_d = dets.new_empty((world * m,) + dets.shape[1:])
    all_c = counts.new_empty(world * m)
    dist.all_gather_into_tensor(all_d, dets.contiguous(), group=group)
    dist.all_gather_into_tensor(all_c, counts.contiguous(), group=group)
    if all(s == m for s in sizes):
        return all_d, all_c
    keep: List[int] = []
    for r, s in enumerate(sizes):
        keep += list(range(r * m, r * m + s))
    idx = torch.tensor(keep, device=dets.device)
    return all_d[idx], all_c[idx]
