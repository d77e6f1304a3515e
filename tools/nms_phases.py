#!/usr/bin/env python
"""Per-phase breakdown of nms_fast_kernel on the bench workload (debug hook plyolo_debug_nms_profile) and the
general-path flags of the images.  PLYOLO_NO_PDL=1 gives the phases without the score kernel running beside it."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth

B = int(os.environ.get("NMS_B", "32"))
G = 4  # class groups per image (kGroups)
size = int(os.environ.get("NMS_SIZE", "640"))
heads = [torch.from_numpy(h).cuda() for h in synth.make_heads(B, size, 80, seed=int(sys.argv[1]) if len(sys.argv) > 1 else 0)]
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
ws = [w for (k, w) in ops._WS.items() if k[2] == "post"][0]
ctr = ws[: B * 16 * 4].view(torch.int32).cpu().numpy().reshape(B, 16)
names = ["wait for tiles", "load + max-first rounds", "survivor histogram", "prefix + scatter", "sort + sweep",
         "cluster sync 1", "gather lists + sync 2", "rank + write"]
us = lambda a: a / 1965.0
done = p[:, 8] > 0
d = np.diff(p[:, :9], axis=1)[done]
print("phase us over (image, group) CTAs that finished on the fast path (%d of %d): mean / max" % (done.sum(), len(p)))
for i, n in enumerate(names):
    print("  %-22s %7.2f %7.2f" % (n, us(d[:, i]).mean(), us(d[:, i]).max()))
print("  %-22s %7.2f %7.2f" % ("  of which load", us(p[done, 9] - p[done, 1]).mean(), us(p[done, 9] - p[done, 1]).max()))
print("  %-22s %7.2f %7.2f" % ("  warp 0 in warp_sort", us(p[done, 15]).mean(), us(p[done, 15]).max()))
work = p[done, 8] - p[done, 1]
print("  %-22s %7.2f %7.2f" % ("after the tiles", us(work).mean(), us(work).max()))
print("n per group (max %d)" % p[:, 10].max(), p[:, 10].reshape(B, G)[:8].tolist())
print("kept per group (max %d), survivors of the rounds per group: mean %.0f max %d (of n mean %.0f)" % (p[:, 11].max(), p[:, 14].mean(), p[:, 14].max(), p[:, 10].mean()))
# absolute timeline (globaltimer, ns): when each image's tiles were complete and when its NMS was done
t_flag = p[:, 13].reshape(B, G).min(1).astype(np.float64)
t_end = p[:, 12].reshape(B, G).max(1).astype(np.float64)
t0 = t_flag.min()
print("image: tiles complete at / NMS done at (us after the first image's tiles were complete)")
print("  " + "  ".join("%d: %.1f/%.1f" % (b, (t_flag[b] - t0) / 1e3, (t_end[b] - t0) / 1e3) for b in range(0, B, max(1, B // 16))))
print("  last image's tiles complete at %.1f us; last NMS done at %.1f us -> exposed tail %.1f us" % ((t_flag.max() - t0) / 1e3, (t_end.max() - t0) / 1e3, (t_end.max() - t_flag.max()) / 1e3))
print("ncross", ctr[:, 9].tolist())
print("general-path flags", ctr[:, 10].tolist(), "tiles done", ctr[:, 11].tolist()[:4])
