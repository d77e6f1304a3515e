#!/usr/bin/env python
"""Times the incumbent on the GPU box: the reference's own op chain (torch / torchvision CUDA kernels, replayed by
oracle/torch_ops_replay.py) on CUDA tensors, synchronised on both sides — SURVEY §2.1 / §8d "the true incumbent"
(models/evaluators/postprocess.py:36-41 -> torchvision's sm_100 nms_kernel; yolox_loss.py:43-118).
Usage: incumbent_time.py [B] [size]"""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_ops_replay as R
from pl_yolo_b200 import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
size = int(sys.argv[2]) if len(sys.argv) > 2 else 640
lmax = 120 if size <= 640 else 500
strides = [8, 16, 32]
dev = "cuda"
heads = [torch.from_numpy(h).to(dev) for h in synth.make_heads(B, size, 80, seed=0, objects_per_image=12 if size <= 640 else 40)]
labels = torch.from_numpy(synth.make_labels(B, size, lmax, 80, seed=1, min_gt=1 if size <= 640 else 250)).to(dev)
hw = synth.level_shapes(size)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts), sorted(ts)[len(ts) // 2]


def dn():
    preds, _ = R.decode(heads, strides, True)
    return R.postprocess(preds, 0.01, 0.65)


preds_t, _ = R.decode(heads, strides, False)


def sim():
    return R.simota(preds_t, labels, hw, strides, stable=False)


out = {"B": B, "size": size}
best, med = timed(dn, 10)
out["decode_nms"] = {"best_ms": best * 1e3, "median_ms": med * 1e3, "img_per_s": B / best}
best, med = timed(sim, 3)
out["simota"] = {"best_ms": best * 1e3, "median_ms": med * 1e3, "img_per_s": B / best}
print(json.dumps(out))
