#!/usr/bin/env python
"""Day-0 probe: how do the ATen / torchvision CUDA kernels the
reference path calls behave bit-wise on this B200?

Only `torch` / `torchvision` library ops are executed here (no reference code, no product
code).  Raw inputs/outputs are dumped to gpurun_out/probe/ so hypotheses can be re-tested
offline; a JSON summary is printed and saved.

Questions (SURVEY.md §8a traps):
  T9  reduction order of `x.sum(-1)` over 80 contiguous floats and of the top-10 slice sum
  T10 stability of `sort(stable=False)` by length
  T7  sigmoid == 1/(1+exp(-x)) bit-wise; BCE == log1p/log formula
  T5  FMA shape inside torchvision::nms's IoU
  T1  first-index rule of max(dim) / min(dim) on ties
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torchvision

OUT = os.path.join(os.environ.get("GRAFT_REPO_ROOT", "."), "gpurun_out", "probe")
os.makedirs(OUT, exist_ok=True)
dev = torch.device(os.environ.get("PROBE_DEV", "cuda:0"))
summary = {"torch": torch.__version__, "torchvision": torchvision.__version__,
           "gpu": (torch.cuda.get_device_name(0) if torch.cuda.is_available() else "none"), "cpu_count": os.cpu_count()}


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa
        return "ERR %r" % (e,)


summary["lscpu"] = sh("lscpu | grep -E 'Model name|Socket|Core|Thread|^CPU\\(s\\)'")
summary["nvidia_smi"] = sh("nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit,memory.total --format=csv")
summary["ref_visible"] = os.path.exists("/root/reference")

f32 = np.float32


# ----------------------------------------------------------------------------- T9 sum order
def tree_adjacent(v):
    """balanced tree, adjacent pairing (shfl_down offsets 1,2,4,...) over axis -1 (pow2 length)."""
    v = v.copy()
    n = v.shape[-1]
    off = 1
    while off < n:
        nxt = np.zeros_like(v)
        nxt[..., : n - off] = v[..., off:]
        v = (v + nxt).astype(f32)
        off <<= 1
    return v[..., 0]


def hyp_lanes(x, bw):
    """ATen Reduce.cuh hypothesis: bw lanes, vt0=4 strided accumulators, then (smem fold to 32) + shfl tree."""
    n = x.shape[-1]
    acc = [np.zeros(x.shape[:-1] + (bw,), f32) for _ in range(4)]
    nchunk = (n + bw - 1) // bw
    # full unrolled groups of 4*bw then tail; accumulator i gets chunks i, i+4, ...
    for ch in range(nchunk):
        lo = ch * bw
        hi = min(n, lo + bw)
        acc[ch % 4][..., : hi - lo] = (acc[ch % 4][..., : hi - lo] + x[..., lo:hi]).astype(f32)
    v = acc[0]
    for i in range(1, 4):
        v = (v + acc[i]).astype(f32)
    while v.shape[-1] > 32:
        half = v.shape[-1] // 2
        v = (v[..., :half] + v[..., half:]).astype(f32)
    return tree_adjacent(v)


def hyp_seq(x):
    s = np.zeros(x.shape[:-1], f32)
    for i in range(x.shape[-1]):
        s = (s + x[..., i]).astype(f32)
    return s


res = {}
g = torch.Generator().manual_seed(1234)
dump = {}
for (G, N) in [(1, 1), (1, 4), (1, 8), (1, 15), (1, 16), (2, 16), (3, 100), (60, 3000), (120, 6000)]:
    x = (5 * torch.rand(G, N, 80, generator=g)).float()
    y = x.to(dev).sum(-1).cpu().numpy()
    xn = x.numpy()
    r = {}
    for name, fn in [("lanes32", lambda a: hyp_lanes(a, 32)), ("lanes64", lambda a: hyp_lanes(a, 64)),
                     ("lanes16", lambda a: hyp_lanes(a, 16)), ("seq", hyp_seq)]:
        r[name] = float((fn(xn) == y).mean())
    res["sum80_G%d_N%d" % (G, N)] = r
    if G * N <= 300:
        dump["sum80_x_%d_%d" % (G, N)] = xn
        dump["sum80_y_%d_%d" % (G, N)] = y
# other class counts (C != 80) for generality
for C in [1, 2, 3, 20, 31, 32, 33, 64, 65, 91, 128, 129, 200]:
    x = (5 * torch.rand(7, 333, C, generator=g)).float()
    y = x.to(dev).sum(-1).cpu().numpy()
    xn = x.numpy()
    bw = 1
    while bw * 2 <= min(C, 32):
        bw *= 2
    r = {"lanes_pow2": float((hyp_lanes(xn, bw) == y).mean()), "seq": float((hyp_seq(xn) == y).mean()), "bw": bw}
    res["sumC%d" % C] = r
    dump["sumC_x_%d" % C] = xn[:, :16]
    dump["sumC_y_%d" % C] = y[:, :16]

# top-10 slice sum, exactly as the reference forms it: sorted desc [G,Nc] -> [:, :10] -> sum(1)
for (G, N) in [(1, 10), (1, 50), (3, 7), (5, 9), (17, 400), (58, 5000), (120, 8400), (500, 20000)]:
    x = torch.rand(G, N, generator=g).float()
    xd = x.to(dev)
    srt, _ = xd.sort(descending=True)
    k = min(10, N)
    top = srt[:, :k]
    y = top.sum(1).cpu().numpy()
    tn = top.cpu().numpy()
    r = {}
    for bw in (8, 16, 32, 4):
        r["lanes%d" % bw] = float((hyp_lanes(tn, bw) == y).mean())
    r["seq"] = float((hyp_seq(tn) == y).mean())
    # contiguous copy for comparison
    y2 = top.contiguous().sum(1).cpu().numpy()
    r["contig_same"] = float((y2 == y).mean())
    res["top10_G%d_N%d" % (G, N)] = r
    if G <= 20:
        dump["top10_x_%d_%d" % (G, N)] = tn
        dump["top10_y_%d_%d" % (G, N)] = y
# labels.sum(2) over 5 elements
x = torch.rand(32, 120, 5, generator=g).float() * 300
y = x.to(dev).sum(2).cpu().numpy()
res["sum5"] = {"seq": float((hyp_seq(x.numpy()) == y).mean()), "lanes4": float((hyp_lanes(x.numpy(), 4) == y).mean())}
dump["sum5_x"] = x.numpy()[:2]
dump["sum5_y"] = y[:2]
summary["T9_sum_order"] = res

# ----------------------------------------------------------------------------- T10 sort stability
res = {}
for n in [2, 8, 16, 31, 32, 33, 64, 100, 128, 129, 500, 1000, 2048, 4096, 4097, 5000, 20000]:
    for desc in (False, True):
        stable_cnt = 0
        for trial in range(20):
            keys = torch.randint(0, 3, (n,), generator=g).float().to(dev)
            v, idx = keys.sort(descending=desc)
            idx = idx.cpu().numpy()
            v = v.cpu().numpy()
            ok = True
            for kv in (0.0, 1.0, 2.0):
                ii = idx[v == kv]
                if len(ii) > 1 and not (np.diff(ii) > 0).all():
                    ok = False
            stable_cnt += ok
        res["n%d_desc%d" % (n, int(desc))] = stable_cnt
# 2-D sort (the reference sorts ious [G,Nc] along dim 1)
for (G, n) in [(4, 20), (4, 33), (60, 5000), (120, 8400)]:
    keys = torch.randint(0, 3, (G, n), generator=g).float().to(dev)
    v, idx = keys.sort(descending=True)
    idx = idx.cpu().numpy(); v = v.cpu().numpy()
    ok = 0
    for r_ in range(G):
        good = True
        for kv in (0.0, 1.0, 2.0):
            ii = idx[r_][v[r_] == kv]
            if len(ii) > 1 and not (np.diff(ii) > 0).all():
                good = False
        ok += good
    res["2d_G%d_n%d" % (G, n)] = "%d/%d" % (ok, G)
summary["T10_sort_stable_of20"] = res

# ----------------------------------------------------------------------------- T7 elementwise
x = (torch.randn(1 << 20, generator=g) * 6).float()
xd = x.to(dev)
s = torch.sigmoid(xd)
s2 = 1.0 / (1.0 + torch.exp(-xd))
s3 = torch.reciprocal(1.0 + torch.exp(-xd))
xs = xd.clone(); xs.sigmoid_()
res = {"sigmoid==1/(1+exp(-x))": float((s == s2).float().mean()), "sigmoid==recip": float((s == s3).float().mean()),
       "sigmoid_==sigmoid": float((s == xs).float().mean())}
p = torch.rand(1 << 20, generator=g).float().to(dev)
p[:8] = torch.tensor([0.0, 1.0, 1e-30, 1e-45, 0.5, 0.99999994, 1e-10, 5e-324], device=dev)
for tval in (0.0, 1.0):
    t = torch.full_like(p, tval)
    b = torch.nn.functional.binary_cross_entropy(p, t, reduction="none")
    l1 = torch.clamp(torch.log1p(-p), min=-100.0)
    l0 = torch.clamp(torch.log(p), min=-100.0)
    f1 = (t - 1) * l1 - t * l0
    f2 = (t - 1) * torch.clamp(torch.log(1 - p), min=-100.0) - t * l0
    res["bce_t%d==log1p" % int(tval)] = float((b == f1).float().mean())
    res["bce_t%d==log(1-p)" % int(tval)] = float((b == f2).float().mean())
# save samples of libdevice-backed ops for later comparison with product kernels
xs_ = (torch.randn(4096, generator=g) * 6).float()
ps_ = torch.rand(4096, generator=g).float()
dump["ew_x"] = xs_.numpy()
dump["ew_exp"] = torch.exp(xs_.to(dev)).cpu().numpy()
dump["ew_sigmoid"] = torch.sigmoid(xs_.to(dev)).cpu().numpy()
dump["ew_p"] = ps_.numpy()
dump["ew_log"] = torch.log(ps_.to(dev)).cpu().numpy()
dump["ew_log1p_neg"] = torch.log1p(-ps_.to(dev)).cpu().numpy()
dump["ew_sqrt"] = torch.sqrt(ps_.to(dev)).cpu().numpy()
# CPU vs CUDA bit-equality of the same ops
for name, fn, arg in [("exp", torch.exp, xs_), ("sigmoid", torch.sigmoid, xs_), ("log", torch.log, ps_),
                      ("log1p", lambda a: torch.log1p(-a), ps_), ("sqrt", torch.sqrt, ps_)]:
    res["cpu==cuda_" + name] = float((fn(arg) == fn(arg.to(dev)).cpu()).float().mean())
summary["T7_elementwise"] = res

# ----------------------------------------------------------------------------- T1 max/min ties
x = torch.zeros(64, 80)
x[:, 7] = 1.0; x[:, 30] = 1.0; x[:, 79] = 1.0
mx = torch.max(x.to(dev), 1)[1].cpu()
mn = torch.min((-x).to(dev), 0)[1].cpu()
c = torch.zeros(9, 50); c[2] = -1; c[5] = -1
amin = torch.min(c.to(dev), dim=0)[1].cpu()
amax = (-c).to(dev).argmax(0).cpu()
summary["T1_ties"] = {"max_dim1_first": bool((mx == 7).all()), "min_dim0_first": bool((amin == 2).all()),
                      "argmax_dim0_first": bool((amax == 2).all()), "max_idx": mx[:3].tolist(), "min_idx": amin[:3].tolist()}

# ----------------------------------------------------------------------------- T5 NMS FMA shape
rng = np.random.default_rng(7)
npair = int(os.environ.get("PROBE_NPAIR", "20000"))
t = 0.65
w = rng.uniform(8, 200, npair).astype(f32)
h = rng.uniform(8, 200, npair).astype(f32)
x0 = rng.uniform(0, 400, npair).astype(f32)
y0 = rng.uniform(0, 400, npair).astype(f32)
dx = (w.astype(np.float64) * (1 - t) / (1 + t) * (1 + rng.uniform(-1e-6, 1e-6, npair))).astype(f32)
A = np.stack([x0, y0, (x0 + w).astype(f32), (y0 + h).astype(f32)], 1).astype(f32)
Bx = np.stack([(x0 + dx).astype(f32), y0, (x0 + dx + w).astype(f32), (y0 + h).astype(f32)], 1).astype(f32)
sup = np.zeros(npair, bool)
for i in range(npair):
    boxes = torch.from_numpy(np.stack([A[i], Bx[i]])).to(dev)
    keep = torchvision.ops.nms(boxes, torch.tensor([1.0, 0.5], device=dev), t)
    sup[i] = keep.numel() == 1


def iou_variants(a, b):
    a64 = a.astype(np.float64); b64 = b.astype(np.float64)
    wi = np.maximum((np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0])).astype(f32), f32(0))
    hi = np.maximum((np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1])).astype(f32), f32(0))
    inter = (wi * hi).astype(f32)
    wa = (a[:, 2] - a[:, 0]).astype(f32); ha = (a[:, 3] - a[:, 1]).astype(f32)
    wb = (b[:, 2] - b[:, 0]).astype(f32); hb = (b[:, 3] - b[:, 1]).astype(f32)
    Sa = (wa * ha).astype(f32); Sb = (wb * hb).astype(f32)
    out = {}
    u0 = ((Sa + Sb).astype(f32) - inter).astype(f32)
    out["nofma"] = (inter / u0).astype(f32)
    tb = (wb.astype(np.float64) * hb.astype(np.float64) + Sa.astype(np.float64)).astype(f32)
    out["fma_b"] = (inter / (tb - inter).astype(f32)).astype(f32)
    ta = (wa.astype(np.float64) * ha.astype(np.float64) + Sb.astype(np.float64)).astype(f32)
    out["fma_a"] = (inter / (ta - inter).astype(f32)).astype(f32)
    return out


res = {}
for name, v in iou_variants(A, Bx).items():
    res[name + "_thr_f32"] = int(((v > f32(t)) != sup).sum())
    res[name + "_thr_f64"] = int(((v.astype(np.float64) > t) != sup).sum())
res["n_suppressed"] = int(sup.sum())
summary["T5_nms_fma_mismatches_of_%d" % npair] = res
dump["nms_A"] = A; dump["nms_B"] = Bx; dump["nms_sup"] = sup

# does CUDA batched_nms take the coordinate-trick branch and agree with manual trick + nms?
nb = 3000
bx = torch.rand(nb, 2, generator=g) * 600
wh = torch.rand(nb, 2, generator=g) * 80 + 4
boxes = torch.cat([bx, bx + wh], 1).float().to(dev)
scores = torch.rand(nb, generator=g).float().to(dev)
cls = torch.randint(0, 80, (nb,), generator=g).float().to(dev)
k1 = torchvision.ops.batched_nms(boxes, scores, cls, 0.65)
mc = boxes.max()
off = cls * (mc + 1)
k2 = torchvision.ops.nms(boxes + off[:, None], scores, 0.65)
summary["batched_nms_is_trick"] = bool(k1.numel() == k2.numel() and (k1 == k2).all())

np.savez_compressed(os.path.join(OUT, "probe_dump.npz"), **dump)
with open(os.path.join(OUT, "probe_summary.json"), "w") as f:
    json.dump(summary, f, indent=1)
print(json.dumps(summary, indent=1))
