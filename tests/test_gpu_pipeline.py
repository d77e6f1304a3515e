"""Batches in flight (pl_yolo_b200.pipeline): the same detections as the one-stream path, whatever the depth."""
import pytest
import torch

from pl_yolo_b200 import PostprocessPipeline, Lanes, ops, synth

pytestmark = pytest.mark.gpu
STRIDES = [8, 16, 32]


def _batches(n, B, size, mode="clustered"):
    return [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, size, 80, seed=700 + i, mode=mode)] for i in range(n)]


@pytest.mark.parametrize("depth,B,size", [(1, 3, 320), (2, 4, 640), (3, 2, 320), (4, 8, 640), (None, 5, 320)])
def test_pipeline_matches_serial(depth, B, size):
    batches = _batches(7, B, size)
    want = [ops.decode_postprocess_raw(h, STRIDES, 0.01, 0.65, False, 10000, 300, 0) for h in batches]
    pipe = PostprocessPipeline(STRIDES, conf_thre=0.01, nms_thre=0.65, depth=depth)
    got = []
    for _ in range(3):  # several passes: lanes and their scratch are reused
        got.clear()
        for h in batches:
            done = pipe.submit(h)
            if done is not None:
                got.append(done)
        got.extend(pipe.drain())
        assert len(got) == len(batches)
        for g, w in zip(got, want):
            assert torch.equal(g[1], w[1]) and torch.equal(g[2], w[2]) and torch.equal(g[0], w[0])
    assert pipe.lanes.depth == (depth or 4)


def test_pipeline_general_path_and_agnostic():
    # a low threshold (thousands of candidates per image: groups overflow into the general NMS path, launched from the
    # device) and the class-agnostic call, several batches in flight
    batches = _batches(4, 2, 640)
    for agnostic in (False, True):
        want = [ops.decode_postprocess_raw(h, STRIDES, 0.001, 0.65, agnostic, 10000, 300, 0) for h in batches]
        pipe = PostprocessPipeline(STRIDES, conf_thre=0.001, nms_thre=0.65, class_agnostic=agnostic, depth=3)
        got = [d for h in batches for d in [pipe.submit(h)] if d is not None] + list(pipe.drain())
        for g, w in zip(got, want):
            assert torch.equal(g[1], w[1]) and torch.equal(g[2], w[2]) and torch.equal(g[0], w[0])


def test_lanes_under_graph_capture():
    batches = _batches(4, 4, 320)
    want = [ops.decode_postprocess_raw(h, STRIDES, 0.01, 0.65, False, 10000, 300, 0) for h in batches]
    lanes = Lanes(2)
    outs = [None] * 8
    cap = torch.cuda.Stream()

    def issue():
        lanes.fork()
        for i in range(8):
            outs[i] = lanes.issue(i, lambda i=i: ops.decode_postprocess_raw(batches[i % 4], STRIDES, 0.01, 0.65, False, 10000, 300, 0))
        lanes.join()

    with torch.cuda.stream(cap):
        issue()
        cap.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            issue()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    for i in range(8):
        for k in range(3):
            assert torch.equal(outs[i][k], want[i % 4][k])


@pytest.mark.parametrize("max_nms,conf", [(600, 0.01), (10000, 0.01), (10000, 0.0005)])
def test_general_path_under_graph_capture(max_nms, conf):
    # max_nms below the candidate count / a very low threshold send images to the general NMS path.  Outside capture the
    # class-split kernel launches it from the device; under capture the host launches it (children of a graph node are
    # not ordered before the rest of the graph).  The outputs are zeroed before every replay and read right behind an
    # event, so rows that arrive late (or never) cannot hide behind an earlier result.
    heads = _batches(1, 3, 640)[0]
    want = ops.decode_postprocess_raw(heads, STRIDES, conf, 0.65, False, max_nms, 300, 0)
    out = (torch.zeros(3, 300, 6, device="cuda"), torch.zeros(3, dtype=torch.int32, device="cuda"),
           torch.zeros(3, 300, dtype=torch.int32, device="cuda"))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ops.decode_postprocess_raw(heads, STRIDES, conf, 0.65, False, max_nms, 300, 0, out=out)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            ops.decode_postprocess_raw(heads, STRIDES, conf, 0.65, False, max_nms, 300, 0, out=out)
        for _ in range(3):
            for t in out:
                t.zero_()
            g.replay()
            ev = torch.cuda.Event()
            ev.record(st)
            ev.synchronize()
            got = [t.clone() for t in out]
            st.synchronize()
            for a, b in zip(got, want):
                assert torch.equal(a, b)
