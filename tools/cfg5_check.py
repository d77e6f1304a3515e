#!/usr/bin/env python
"""cfg5 (every image on the general NMS path, launched from the device): per-step graphs vs eager, outputs re-zeroed
before every replay so that a launch that did not happen cannot hide behind an earlier result."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

B, size = 8, 1280
sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, size, 80, seed=2 * s, objects_per_image=40)] for s in range(4)]
outs = [(torch.zeros(B, 300, 6, device="cuda"), torch.zeros(B, dtype=torch.int32, device="cuda"), torch.zeros(B, 300, dtype=torch.int32, device="cuda")) for _ in range(4)]
run = lambda i: ops.decode_postprocess_raw(sets[i], [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0, out=outs[i])
ref = []
for i in range(4):
    run(i)
    torch.cuda.synchronize()
    ref.append([t.clone() for t in outs[i]])
st = torch.cuda.Stream()
graphs = []
with torch.cuda.stream(st):
    for i in range(4):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            run(i)
        graphs.append(g)
torch.cuda.synchronize()


def check(tag):
    ok = all(torch.equal(outs[i][k], ref[i][k]) for i in range(4) for k in range(3))
    print(tag, "outputs identical:", ok, "counts", [int(outs[i][1].sum()) for i in range(4)])


def zero():
    for o in outs:
        for t in o:
            t.zero_()
    torch.cuda.synchronize()


for mode in ("eager", "graph"):
    zero()
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for r in range(8):
            for i in range(4):
                if mode == "eager":
                    run(i)
                else:
                    graphs[i].replay()
        e1.record(st)
    torch.cuda.synchronize()
    print("%s: %.1f us/step" % (mode, e0.elapsed_time(e1) * 1000 / 32))
    check(mode)
zero()
with torch.cuda.stream(st):
    graphs[0].replay()
    e = torch.cuda.Event(); e.record(st); e.synchronize()   # event completion right behind the graph
    print("right after the event behind one replay: counts", int(outs[0][1].sum()), "(want %d)" % int(ref[0][1].sum()))
torch.cuda.synchronize()
check("after sync")
