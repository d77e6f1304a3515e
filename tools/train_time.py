#!/usr/bin/env python
"""Training-step time of the YOLOX loss (decode + SimOTA + loss tail, forward + backward to the head maps) at cfg3
size: fused loss tail (N2) vs the batched torch tail of the same shim.  4 rotating input sets, CUDA events."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import YOLOXLoss, synth

B, LMAX = 32, 120
sets = []
for s in range(4):
    heads = [torch.from_numpy(h).cuda().requires_grad_(True) for h in synth.make_heads(B, 640, 80, seed=2 * s)]
    labels = torch.from_numpy(synth.make_labels(B, 640, LMAX, 80, seed=2 * s + 1)).cuda()
    sets.append((heads, labels))


def step(mod, i):
    heads, labels = sets[i % 4]
    for h in heads:
        h.grad = None
    out = mod(heads, labels)
    out["loss"].backward()
    return out


for name, mod in (("fused loss tail", YOLOXLoss(80, [8, 16, 32])), ("torch loss tail", YOLOXLoss(80, [8, 16, 32], fused_loss=False))):
    for i in range(6):
        step(mod, i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    N = 20
    e0.record()
    for i in range(N):
        step(mod, i)
    e1.record()
    torch.cuda.synchronize()
    print("%-16s forward+backward %.1f us/step (B=32, 640^2, G~U{1..120}; eager, one host sync per step)" % (name, e0.elapsed_time(e1) * 1000 / N))
