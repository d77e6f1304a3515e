#!/usr/bin/env python
"""A few eager fused decode+postprocess calls on the cfg2 workload (for ncu -k ... -s ... -c ...)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth
B = int(os.environ.get("NMS_B", "32"))
sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=2 * s)] for s in range(2)]
for i in range(6):
    ops.decode_postprocess_raw(sets[i % 2], [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
torch.cuda.synchronize()
