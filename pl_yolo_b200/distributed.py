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

from . import _lib


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

    def submit(self, slot: int, buf: torch.Tensor = None, gbuf: torch.Tensor = None, push=None) -> None:
        """Default: NCCL all-gather of `buf` into `gbuf`; `push`: any callable that enqueues the exchange on the
        current (side) stream instead, e.g. PeerDetections.push (copy-engine peer copies, no SM at all)."""
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream())
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            if push is not None:
                push()
            else:
                dist.all_gather_into_tensor(gbuf, buf, group=self.group)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[slot] = ev

    def finish(self) -> None:
        torch.cuda.current_stream().wait_stream(self.side)


class _DevMem:
    """A raw device allocation seen through __cuda_array_interface__ (zero-copy view for torch.as_tensor)."""

    def __init__(self, ptr: int, n_float: int):
        self.__cuda_array_interface__ = {"shape": (n_float,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _as_tensor(ptr: int, n_float: int, device) -> torch.Tensor:
    return torch.as_tensor(_DevMem(ptr, n_float), device=device)


class PeerDetections:
    """Gathered detection buffers that the NMS kernels fill directly on every rank (no collective).

    Every rank owns `slots` gathered buffers laid out like `fused_det_buffer` per rank block,
    `[world][b_loc * max_det * 6 floats | b_loc int32]`, and maps the other ranks' buffers into its address space through
    CUDA IPC (plyolo_peer_alloc / plyolo_peer_open; the handles travel once with all_gather_object).  `outputs(slot)` gives the local views the C ABI writes
    (`dets`, `counts`: this rank's own block of its own buffer) and `peers(slot)` the addresses of this rank's block
    inside every OTHER rank's buffer: pass both to `ops.decode_postprocess_raw(..., out=..., peers=...)` and the
    all-gather of SURVEY §8e happens inside `nms_fast_kernel` / `nms_general_kernel` as plain stores over NVLink.
    `fence()` orders consumption across ranks (barrier after the stream's work)."""

    def __init__(self, b_loc: int, max_det: int, device, slots: int = 1, group=None):
        import ctypes
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b_loc, self.max_det, self.group = b_loc, max_det, group
        self.block = b_loc * max_det * 6 + b_loc          # floats per rank block
        n_float = self.world * self.block
        L = _lib.lib()
        self._owned, self._opened = [], []
        self.local, handles = [], []
        with torch.cuda.device(device):
            for _ in range(slots):
                ptr, h = ctypes.c_void_p(), ctypes.create_string_buffer(64)
                _lib.check(L.plyolo_peer_alloc(n_float * 4, ctypes.byref(ptr), h), "plyolo_peer_alloc")
                self._owned.append(ptr.value)
                handles.append(h.raw)
                self.local.append(_as_tensor(ptr.value, n_float, device))
            every = [None] * self.world
            dist.all_gather_object(every, handles, group=group)
            self.remote_ptr = []   # [slot][rank] -> device address of that rank's gathered buffer (None for this rank)
            for s in range(slots):
                row = []
                for r in range(self.world):
                    if r == self.rank:
                        row.append(None)
                        continue
                    ptr = ctypes.c_void_p()
                    _lib.check(L.plyolo_peer_open(every[r][s], ctypes.byref(ptr)), "plyolo_peer_open")
                    self._opened.append(ptr.value)
                    row.append(ptr.value)
                self.remote_ptr.append(row)
            self._remote_views = [[_as_tensor(p, n_float, device) for p in row if p is not None] for row in self.remote_ptr]
        self.keep = [torch.empty((b_loc, max_det), dtype=torch.int32, device=device) for _ in range(slots)]

    def close(self) -> None:
        L = _lib.lib()
        torch.cuda.synchronize()
        for p in self._opened:
            L.plyolo_peer_close(p, 1)
        dist.barrier(group=self.group)   # nobody maps our buffers any more
        for p in self._owned:
            L.plyolo_peer_close(p, 0)
        self._opened, self._owned, self.local, self._remote_views = [], [], [], []

    def outputs(self, slot: int):
        base = self.local[slot][self.rank * self.block:(self.rank + 1) * self.block]
        n = self.b_loc * self.max_det * 6
        return base[:n].view(self.b_loc, self.max_det, 6), base[n:].view(torch.int32), self.keep[slot]

    def peers(self, slot: int):
        n = self.b_loc * self.max_det * 6
        off = self.rank * self.block * 4
        pd = [p + off for p in self.remote_ptr[slot] if p is not None]
        return pd, [p + n * 4 for p in pd]

    def push(self, slot: int) -> None:
        """Copies this rank's finished block into every other rank's gathered buffer with plain device-to-device copies
        on the current stream: peer-mapped addresses, so the copy engines move the 230 KB over NVLink and no SM is
        involved (use it on a side stream: DetectionExchange.submit(slot, push=...))."""
        lo, hi = self.rank * self.block, (self.rank + 1) * self.block
        src = self.local[slot][lo:hi]
        for t in self._remote_views[slot]:
            t[lo:hi].copy_(src, non_blocking=True)

    def gathered(self, slot: int):
        return split_gathered(self.local[slot], self.world, self.b_loc, self.max_det)

    def fence(self) -> None:
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)
