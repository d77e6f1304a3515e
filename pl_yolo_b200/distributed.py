"""Image sharding and the detection all-gather.

The path shards by image with no data-path collective: every rank decodes / NMSes / assigns its own
contiguous slice of the batch.  The only exchange is the evaluator's: fixed-size padded detections
`[B_loc, 300, 6]` + `[B_loc]` counts, all-gathered once per step over NCCL / NVLink — as ONE collective on a
fused buffer (`fused_det_buffer`), issued on a side stream so that it overlaps the next step's score kernel
(`DetectionExchange`).  The reference has no multi-GPU path at all (train.py:33, devices=1).
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


def fused_det_buffer(b_loc: int, max_det: int, device) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """One allocation holding a rank's padded detections and counts back to back, so that ONE collective moves both:
    -> (buf float32 [b_loc * max_det * 6 + b_loc], dets view [b_loc, max_det, 6], counts int32 view [b_loc])."""
    n = b_loc * max_det * 6
    buf = torch.empty(n + b_loc, dtype=torch.float32, device=device)
    return buf, buf[:n].view(b_loc, max_det, 6), buf[n:].view(torch.int32)


def split_gathered(gbuf: torch.Tensor, world: int, b_loc: int, max_det: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Views of an all-gathered fused buffer [world * (b_loc * max_det * 6 + b_loc)] in global image order:
    -> (dets [world * b_loc, max_det, 6], counts [world * b_loc] int32); copies (ranks are interleaved with counts)."""
    n = b_loc * max_det * 6
    g = gbuf.view(world, n + b_loc)
    return g[:, :n].reshape(world * b_loc, max_det, 6), g[:, n:].contiguous().view(torch.int32).reshape(world * b_loc)


class DetectionExchange:
    """Per-step all-gather of the fused detection buffers, overlapped with the following steps.

    `submit(buf, gbuf)` is called after the step's kernels were enqueued on the current stream: the collective runs on
    a side stream behind an event, so the next step's score kernel does not wait for it.  A buffer pair may be reused
    once `wait_for(slot)` was called for it (the compute stream then waits for that slot's collective); `finish()` joins
    everything back into the current stream."""

    def __init__(self, device, slots: int, group=None):
        self.side = torch.cuda.Stream(device)
        self.group = group
        self.done = [None] * slots

    def wait_for(self, slot: int) -> None:
        if self.done[slot] is not None:
            torch.cuda.current_stream().wait_event(self.done[slot])

    def submit(self, slot: int, buf: torch.Tensor, gbuf: torch.Tensor) -> None:
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            dist.all_gather_into_tensor(gbuf, buf, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev

    def finish(self) -> None:
        torch.cuda.current_stream().wait_stream(self.side)
