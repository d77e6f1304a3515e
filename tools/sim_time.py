#!/usr/bin/env python
"""Times the cfg3 SimOTA assignment (4 rotating input sets, CUDA events over 40 calls after warm-up)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

B, LMAX = 32, 120
sets = []
for s in range(4):
    heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=2 * s)]
    labels = torch.from_numpy(synth.make_labels(B, 640, LMAX, 80, seed=2 * s + 1)).cuda()
    preds, _ = ops.decode_raw(heads, [8, 16, 32], False)
    sets.append((preds, labels))
hw = [80, 80, 40, 40, 20, 20]
for i in range(8):
    ops.simota_assign_raw(*sets[i % 4], hw, [8, 16, 32])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 40
e0.record()
for i in range(N):
    ops.simota_assign_raw(*sets[i % 4], hw, [8, 16, 32])
e1.record()
torch.cuda.synchronize()
print("%s simota us/call (eager, incl. launch gaps): %.1f" % (sys.argv[1] if len(sys.argv) > 1 else "", e0.elapsed_time(e1) * 1000 / N))
