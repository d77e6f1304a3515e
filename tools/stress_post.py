#!/usr/bin/env python
"""Stress: repeats the fused and the preds-based postprocess on cfg2-sized inputs and checks every result
against the first one (flushes out launch failures and nondeterminism)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

which = sys.argv[1] if len(sys.argv) > 1 else "both"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
B = 32
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, 0)]
d, c, k = ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
p2 = torch.zeros(B, 8400, 85, device="cuda")
p2[:, :300, :4] = d[..., :4]
p2[:, :300, 4] = d[..., 4]
p2[:, :300, 5:].scatter_(2, d[..., 5].long().unsqueeze(-1), 1.0)
preds, _ = ops.decode_raw(heads, [8, 16, 32], True)
torch.cuda.synchronize()
ref = {}
for it in range(iters):
    for name in (["fused", "p2", "preds"] if which == "both" else [which]):
        if name == "fused":
            out = ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
        elif name == "p2":
            out = ops.postprocess_raw(p2, 0.01, 0.65, False, 10000, 300, 0)
        else:
            out = ops.postprocess_raw(preds, 0.01, 0.65, False, 10000, 300, 0)
        torch.cuda.synchronize()
        if name not in ref:
            ref[name] = [t.clone() for t in out]
        else:
            for a, b in zip(out, ref[name]):
                assert torch.equal(a, b), (name, it)
print("stress ok", which, iters)
