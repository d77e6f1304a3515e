#!/usr/bin/env python
"""Three warm calls of every hot-path entry on the bench workloads (for one `ncu --set full` capture of all kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth
B = 32
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=0)]
labels = torch.from_numpy(synth.make_labels(B, 640, 120, 80, seed=1)).cuda()
hw = [80, 80, 40, 40, 20, 20]
for _ in range(3):
    ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
    preds, _ = ops.decode_raw(heads, [8, 16, 32], False)
    ops.simota_assign_raw(preds, labels, hw, [8, 16, 32])
torch.cuda.synchronize()
