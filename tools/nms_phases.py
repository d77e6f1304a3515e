#!/usr/bin/env python
"""Per-phase cycle breakdown of nms_group_kernel on the bench workload (debug hook plyolo_debug_nms_profile)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth

B, G = 32, 4
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 0)]
L = _lib.lib()
L.plyolo_debug_nms_profile.argtypes = [ctypes.c_void_p]
prof = torch.zeros((B * G, 16), dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
L.plyolo_debug_nms_profile(prof.data_ptr())
ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
torch.cuda.synchronize()
L.plyolo_debug_nms_profile(None)
p = prof.cpu().numpy().astype(np.int64)
names = ["stage+hist", "prefix+scatter", "sort+gather", "sweep", "compact", "cluster sync", "gather lists", "rank+write"]
t = np.stack([p[:, 0], p[:, 1], p[:, 2], p[:, 8], p[:, 3], p[:, 4], p[:, 5], p[:, 6], p[:, 7]], 1)
us = lambda a: a / 1965.0
d = np.diff(t[:, :8], axis=1)
print("phase us over (image, group) CTAs: mean / max")
for i, n in enumerate(names[:7]):
    print("  %-15s %7.2f %7.2f" % (n, us(d[:, i]).mean(), us(d[:, i]).max()))
last = p[:, 13] == 1
print("  %-15s %7.2f %7.2f" % (names[7], us(t[last, 8] - t[last, 7]).mean(), us(t[last, 8] - t[last, 7]).max()))
print("  CTA total (to publish)  %7.2f %7.2f" % (us(t[:, 7] - t[:, 0]).mean(), us(t[:, 7] - t[:, 0]).max()))
print("  last CTA total          %7.2f %7.2f" % (us(t[last, 8] - t[last, 0]).mean(), us(t[last, 8] - t[last, 0]).max()))
print("n per group", p[:, 10].reshape(B, G).tolist())
print("kept per group", p[:, 11].reshape(B, G).tolist())
print("ncross", p[::G, 12].tolist(), "fallback images", int(p[:, 14].sum()))
