#!/usr/bin/env python
"""Per-phase cycle breakdown of nms_kernel on the bench workload (debug hook plyolo_debug_nms_profile)."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth

B = 32
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 0)]
L = _lib.lib()
L.plyolo_debug_nms_profile.argtypes = [ctypes.c_void_p]
prof = torch.zeros((B, 16), dtype=torch.int64, device="cuda")
for _ in range(3):
    ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
L.plyolo_debug_nms_profile(prof.data_ptr())
ops.decode_postprocess_raw(heads, [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
torch.cuda.synchronize()
L.plyolo_debug_nms_profile(None)
p = prof.cpu().numpy().astype(np.int64)
p[:, 4] = p[:, 13]  # slot 4 is read before the barrier by the compiler's schedule: use the last warp's exit time
names = ["prefix", "pass1", "scatter+x", "classes", "compact", "sort2", "output"]
d = np.diff(p[:, :8], axis=1) / 1965.0  # us at 1965 MHz
print("phase us (mean / max over images):")
for i, n in enumerate(names):
    print("  %-9s %7.2f %7.2f" % (n, d[:, i].mean(), d[:, i].max()))
print("  total     %7.2f %7.2f" % (d.sum(1).mean(), d.sum(1).max()))
print("Nk", p[:, 10].tolist())
print("Kt", p[:, 11].tolist())
print("ncross", p[:, 12].tolist())
print("slowest class sweep: us", [round((int(v) >> 32) / 1965.0, 1) for v in p[:, 14]], "n", [int(v) & 0xffffffff for v in p[:, 14]])
print("slowest class sort : us", [round((int(v) >> 32) / 1965.0, 1) for v in p[:, 15]], "n", [int(v) & 0xffffffff for v in p[:, 15]])
