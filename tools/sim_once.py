#!/usr/bin/env python
"""Runs the cfg3 SimOTA assignment a few times (for `ncu --metrics gpu__time_duration.sum ...` launch lists)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

B, LMAX = 32, 120
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=0)]
labels = torch.from_numpy(synth.make_labels(B, 640, LMAX, 80, seed=1)).cuda()
preds, _ = ops.decode_raw(heads, [8, 16, 32], False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.simota_assign_raw(preds, labels, [80, 80, 40, 40, 20, 20], [8, 16, 32])
torch.cuda.synchronize()
