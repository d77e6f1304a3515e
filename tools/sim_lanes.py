#!/usr/bin/env python
"""cfg3 SimOTA: (a) steps in flight on NS streams, (b) one step split into NS sub-batches on NS streams (joined per step)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth
from pl_yolo_b200.pipeline import Lanes

B, LMAX = 32, 120
sets = []
for s in range(4):
    heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=2 * s)]
    labels = torch.from_numpy(synth.make_labels(B, 640, LMAX, 80, seed=2 * s + 1)).cuda()
    preds, _ = ops.decode_raw(heads, [8, 16, 32], False)
    sets.append((preds, labels))
hw = [80, 80, 40, 40, 20, 20]
STEPS = 16


def timed(tag, issue):
    cap = torch.cuda.Stream()
    with torch.cuda.stream(cap):
        issue()
        cap.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            issue()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-40s %.1f us/step" % (tag, e0.elapsed_time(e1) * 1000 / (10 * STEPS)))


for ns in (1, 2, 4):
    lanes = Lanes(ns)
    def in_flight():
        lanes.fork()
        for i in range(STEPS):
            lanes.issue(i, lambda i=i: ops.simota_assign_raw(*sets[i % 4], hw, [8, 16, 32]))
        lanes.join()
    timed("steps in flight: %d" % ns, in_flight)
for ns in (2, 4):
    lanes = Lanes(ns)
    def split():
        for i in range(STEPS):
            p, l = sets[i % 4]
            lanes.fork()
            for j in range(ns):
                lo, hi = j * B // ns, (j + 1) * B // ns
                lanes.issue(j, lambda lo=lo, hi=hi: ops.simota_assign_raw(p[lo:hi], l[lo:hi], hw, [8, 16, 32]))
            lanes.join()
    timed("one step split into %d sub-batches" % ns, split)
