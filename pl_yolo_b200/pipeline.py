"""Batches in flight: the detection tail of an evaluation loop, several batches deep.

The reference's evaluation step (models/lit_yolox.py validation_step -> models/utils/postprocess / the evaluators) handles
one batch at a time: decode, filter, NMS, then the next batch.  On the GPU one batch's step has a serial tail — the NMS of
the images whose tiles were scored last (~15 us) — and a start-up (~10 us) during which most SMs idle.  Batches are
independent, so the tail of batch i can run under the score kernel of batch i+1: every batch in flight gets its own CUDA
stream (and with it its own scratch: pl_yolo_b200.ops caches workspaces per stream) and its own output buffers; the SMs
co-schedule the kernels (the score CTA and the NMS CTA are sized to share an SM, csrc/postprocess.cu).  Results are
bit-identical to the one-stream path (tests/test_gpu_pipeline.py); measured on B200 (cfg2, 32 x 640 x 640, 80 classes):
45.6 us/batch with one batch in flight, 28.5 with two, 24.7 with three, 23.2-24.4 with four.

`Lanes` is the mechanism (fork / issue / join on streams; it also works under CUDA-graph capture — bench.py captures
rounds of steps this way); `PostprocessPipeline` is the user-facing loop helper.
"""
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import ops


def default_depth(batch: int) -> int:
    """Batches in flight that paid off on B200: 4 for batches up to 64 images, 2 above (the NMS clusters of more
    concurrent large batches only compete for the same SMs)."""
    return 4 if batch <= 64 else 2


class Lanes:
    """`depth` CUDA streams used round-robin.  fork(): every lane waits for the work already queued on the current
    stream; issue(i, fn): fn() with lane i % depth current; join(): the current stream waits for every lane."""

    def __init__(self, depth: int, device=None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        if not torch.cuda.is_available():
            raise RuntimeError("pl_yolo_b200.pipeline needs a CUDA device (there is no CPU path)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(depth)]

    @property
    def depth(self) -> int:
        return len(self.streams)

    def fork(self) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        for s in self.streams:
            s.wait_event(ev)

    def issue(self, i: int, fn: Callable[[], object]):
        with torch.cuda.stream(self.streams[i % len(self.streams)]):
            return fn()

    def join(self) -> None:
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            cur.wait_event(ev)


class PostprocessPipeline:
    """Fused decode + postprocess (YOLOXDecoder + postprocess of the reference, models/heads/yolox/yolox_decoder.py and
    models/utils/postprocess.py:4-44) with up to `depth` batches in flight.

        pipe = PostprocessPipeline(strides=[8, 16, 32], conf_thre=0.01, nms_thre=0.65)
        for heads in batches:                # heads: the per-level head maps [B, 5+C, H, W] of one batch
            done = pipe.submit(heads)        # -> results of the batch submitted `depth` batches ago, or None
            if done is not None: consume(done)
        for done in pipe.drain(): consume(done)

    Each result is (dets [B, max_det, 6], counts [B] int32, keep_idx [B, max_det] int32) like ops.decode_postprocess_raw;
    it is handed back once the CURRENT stream has been made to wait for it, so the caller uses it like any other tensor.
    The head maps of a submitted batch must stay alive and unmodified until its result was returned."""

    def __init__(self, strides: Sequence[int], conf_thre: float = 0.7, nms_thre: float = 0.45, class_agnostic: bool = False,
                 max_nms: int = 10000, max_det: int = 300, flavor: int = 0, depth: Optional[int] = None, device=None):
        self.strides = [int(s) for s in strides]
        self.args = (float(conf_thre), float(nms_thre), bool(class_agnostic), int(max_nms), int(max_det), int(flavor))
        self._depth_arg, self._device = depth, device
        self.lanes: Optional[Lanes] = None
        self._pending: List[Tuple[torch.cuda.Event, tuple, list]] = []
        self._n = 0

    def _ensure(self, batch: int, device) -> Lanes:
        if self.lanes is None:
            self.lanes = Lanes(self._depth_arg or default_depth(batch), self._device or device)
        return self.lanes

    def _pop(self):
        ev, res, _keepalive = self._pending.pop(0)
        cur = torch.cuda.current_stream(self.lanes.device)
        cur.wait_event(ev)
        for t in res:
            t.record_stream(cur)  # allocated on the lane's stream, consumed on the caller's
        return res

    def submit(self, heads: Sequence[torch.Tensor]):
        heads = list(heads)
        lanes = self._ensure(int(heads[0].shape[0]), heads[0].device)
        done = self._pop() if len(self._pending) >= lanes.depth else None
        lane = lanes.streams[self._n % lanes.depth]
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(lanes.device))  # the head maps were produced on the caller's stream
        lane.wait_event(ready)
        with torch.cuda.stream(lane):
            res = ops.decode_postprocess_raw(heads, self.strides, *self.args)
            ev = torch.cuda.Event()
            ev.record(lane)
        for t in heads:
            t.record_stream(lane)  # produced on the caller's stream, read on the lane's
        self._pending.append((ev, res, heads))
        self._n += 1
        return done

    def drain(self):
        while self._pending:
            yield self._pop()
