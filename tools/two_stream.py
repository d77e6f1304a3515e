#!/usr/bin/env python
"""Batches in flight: the cfg2 step (fused decode+postprocess of one batch) issued round-robin on NS streams, each with
its own workspace and outputs (ops caches scratch per stream), captured into one CUDA graph.  NS=1 is the bench's
single-stream step; NS=2 lets batch i+1's score kernel start under batch i's NMS tail."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pl_yolo_b200 import ops, synth

B = int(os.environ.get("NMS_B", "32"))
size = int(os.environ.get("NMS_SIZE", "640"))
STEPS = 24
NSETS = int(os.environ.get('NMS_SETS', '4'))
sets = [[torch.from_numpy(h).cuda() for h in synth.make_heads(B, size, 80, seed=2 * s)] for s in range(NSETS)]
run = lambda i: ops.decode_postprocess_raw(sets[i % NSETS], [8, 16, 32], 0.01, 0.65, False, 10000, 300, 0)
ref = [run(i) for i in range(NSETS)]
torch.cuda.synchronize()


def timed(ns):
    side = [torch.cuda.Stream() for _ in range(ns)]
    outs = [None] * STEPS
    def issue():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for s in side:
            s.wait_event(ev)
        for i in range(STEPS):
            with torch.cuda.stream(side[i % ns]):
                outs[i] = run(i)
        for s in side:
            e = torch.cuda.Event()
            e.record(s)
            main.wait_event(e)
    cap = torch.cuda.Stream()
    with torch.cuda.stream(cap):
        issue()
        cap.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            issue()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    ok = all(torch.equal(outs[i][k], ref[i % NSETS][k]) for i in range(STEPS) for k in range(3))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("streams %d: %.2f us/step  (results identical to the single-stream run: %s)" % (ns, e0.elapsed_time(e1) * 1000 / (10 * STEPS), ok))


for ns in [int(v) for v in os.environ.get('NS', '1,2,3,1,2').split(',')]:
    timed(ns)
