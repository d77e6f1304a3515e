"""world_size-2 gloo tests of the N>1 host logic (sharding + detection all-gather), CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pl_yolo_b200.distributed import fused_det_buffer, gather_detections, shard_batch, shard_range, split_gathered


def test_shard_range_covers_batch():
    for batch in (1, 7, 32, 33, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        full_d = torch.rand(batch, 300, 6, generator=g)
        full_c = torch.randint(0, 301, (batch,), generator=g, dtype=torch.int32)
        d, c = shard_batch([full_d, full_c])
        lo, hi = shard_range(batch, rank, world)
        assert d.shape[0] == hi - lo
        gd, gc = gather_detections(d, c, batch=batch)
        gd2, gc2 = gather_detections(d, c)  # sizes discovered with a collective
        ok = torch.equal(gd, full_d) and torch.equal(gc, full_c) and torch.equal(gd2, full_d) and torch.equal(gc2, full_c)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_gather_detections_gloo_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def _fused_worker(rank, world, port, b_loc, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        full_d = torch.rand(world * b_loc, 300, 6, generator=g)
        full_c = torch.randint(0, 301, (world * b_loc,), generator=g, dtype=torch.int32)
        buf, dets, counts = fused_det_buffer(b_loc, 300, "cpu")
        dets.copy_(full_d[rank * b_loc:(rank + 1) * b_loc])      # what the C ABI writes through the two views
        counts.copy_(full_c[rank * b_loc:(rank + 1) * b_loc])
        gbuf = torch.empty(world * buf.numel())
        dist.all_gather_into_tensor(gbuf, buf)                    # ONE collective for detections + counts
        gd, gc = split_gathered(gbuf, world, b_loc, 300)
        q.put((rank, bool(torch.equal(gd, full_d) and torch.equal(gc, full_c))))
    finally:
        dist.destroy_process_group()


def test_fused_detection_exchange_gloo_world2():
    """The per-step exchange of bench.py / DESIGN section 7: one all-gather of the fused [dets | counts] buffer."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
