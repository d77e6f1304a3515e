#!/usr/bin/env python
"""Times the fused decode+postprocess (cfg2, 4 rotating input sets, CUDA-graph replay) and its two kernels (stage events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

B = 32
sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=2 * s)] for s in range(4)]
run = lambda i: ops.decode_postprocess_raw(sets[i % 4], [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
for i in range(8):
    run(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(8):
        run(i)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print("%s decode+postprocess us/step (graph): %.2f" % (sys.argv[1] if len(sys.argv) > 1 else "", e0.elapsed_time(e1) * 1000 / 80))
