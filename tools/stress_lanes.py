#!/usr/bin/env python
"""Stress for batches in flight: many graph replays of 4-lane rounds over rotating input sets (two batch sizes, one of
them with images on the general NMS path), every step's outputs compared with the one-stream results."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth
from pl_yolo_b200.pipeline import Lanes

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for B, size, conf, max_nms in ((32, 640, 0.01, 10000), (6, 640, 0.01, 900), (1, 640, 0.01, 10000)):
    sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, size, 80, seed=300 + 2 * s)] for s in range(4)]
    run = lambda i: ops.decode_postprocess_raw(sets[i % 4], [8, 16, 32], conf, 0.65, False, max_nms, 300, 0)
    ref = [run(i) for i in range(4)]
    torch.cuda.synchronize()
    lanes, STEPS = Lanes(4), 16
    outs = [None] * STEPS
    cap = torch.cuda.Stream()

    def issue():
        lanes.fork()
        for i in range(STEPS):
            outs[i] = lanes.issue(i, lambda i=i: run(i))
        lanes.join()

    with torch.cuda.stream(cap):
        issue()
        cap.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            issue()
    bad = 0
    for it in range(iters):
        for o in outs:
            for t in o:
                t.zero_()
        g.replay()
        torch.cuda.synchronize()
        for i in range(STEPS):
            for k in range(3):
                if not torch.equal(outs[i][k], ref[i % 4][k]):
                    bad += 1
    print("B=%d max_nms=%d: %d replays x %d steps, mismatching outputs: %d" % (B, max_nms, iters, STEPS, bad))
    assert bad == 0
print("stress lanes ok")
