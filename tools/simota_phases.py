#!/usr/bin/env python
"""Per-phase breakdown of simota_match_kernel on the cfg3 workload (debug hook plyolo_debug_simota_profile)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth

B, LMAX = 32, 120
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=0)]
labels = torch.from_numpy(synth.make_labels(B, 640, LMAX, 80, seed=1)).cuda()
preds, _ = ops.decode_raw(heads, [8, 16, 32], False)
hw = [80, 80, 40, 40, 20, 20]
L = _lib.lib()
L.plyolo_debug_simota_profile.argtypes = [ctypes.c_void_p]
NC = (LMAX + 7) // 8
prof = torch.zeros((B, NC, 16), dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.simota_assign_raw(preds, labels, hw, [8, 16, 32])
L.plyolo_debug_simota_profile(prof.data_ptr())
ops.simota_assign_raw(preds, labels, hw, [8, 16, 32])
torch.cuda.synchronize()
L.plyolo_debug_simota_profile(None)
p = prof.cpu().numpy().reshape(-1, 16)
p = p[p[:, 0] != 0]
names = ["in-both lists", "(none)", "dynamic k", "cheap bounds", "queue k", "eval 1", "U, tight bounds", "eval 2", "select+claim", "conflicts", "finalize", "-"]
d = np.diff(p[:, :12], axis=1) / 1965.0
print("active CTAs", len(p), " phase us: mean / max")
for i, n in enumerate(names[:11]):
    print("  %-14s %7.2f %7.2f" % (n, d[:, i].mean(), d[:, i].max()))
print("  CTA total      %7.2f %7.2f" % ((p[:, 11] - p[:, 0]).mean() / 1965.0, (p[:, 11] - p[:, 0]).max() / 1965.0))
print("  span first start -> last end (us): %.1f" % ((p[:, 11].max() - p[:, 0].min()) / 1965.0))
print("  pairs/CTA mean %.0f, eval1 %.1f, eval2 %.1f" % (p[:, 14].mean(), p[:, 12].mean(), p[:, 13].mean()))
