#!/usr/bin/env python
"""Per-phase cycle totals of the score kernel's consumer groups (debug hook plyolo_debug_score_profile)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth
B = 32
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=0)]
L = _lib.lib()
L.plyolo_debug_score_profile.argtypes = [ctypes.c_void_p]
KC = 4  # kConsumers of postprocess.cu (the kernel indexes [blockIdx.x][kConsumers][8])
prof = torch.zeros((148, KC, 8), dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
L.plyolo_debug_score_profile(prof.data_ptr())
ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
torch.cuda.synchronize()
L.plyolo_debug_score_profile(None)
p = prof.cpu().numpy().astype(np.float64)
tiles = p[:, :, 7].sum()
names = {1: "wait for data", 2: "class sweep + box", 4: "release, ballots, atomics issued, barrier", 5: "slot records", 6: "bucket records"}
tot = 0
for k, n in names.items():
    us = p[:, :, k].sum() / tiles / 1965.0
    tot += us
    print("  %-26s %6.2f us per tile" % (n, us))
print("  total per tile per group   %6.2f us   (tiles %d)" % (tot, tiles))
