#!/usr/bin/env python
"""Times the score stage alone (memset + score_kernel<fused>; debug hook plyolo_debug_skip_nms) and the whole fused
decode+postprocess on the cfg2 workload (4 rotating input sets, CUDA-graph replay).  PLYOLO_LIB selects a variant."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import _lib, ops, synth

B = int(os.environ.get("NMS_B", "32"))
sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, 640, 80, seed=2 * s)] for s in range(4)]
run = lambda i: ops.decode_postprocess_raw(sets[i % 4], [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
L = _lib.lib()
L.plyolo_debug_skip_nms.argtypes = [ctypes.c_int]


def timed(tag):
    for i in range(8):
        run(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(8):
            run(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("%s %s: %.2f us/step" % (sys.argv[1] if len(sys.argv) > 1 else "", tag, e0.elapsed_time(e1) * 1000 / 80))


timed("decode+postprocess")
L.plyolo_debug_skip_nms(1)
timed("score stage alone")
L.plyolo_debug_skip_nms(0)
