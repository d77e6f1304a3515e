#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REAL REFERENCE (imported from /root/reference, CPU).

Run in the build container only (`python oracle/gen_golden.py`); the GPU box has no
/root/reference.  Nothing is restated here: decode/postprocess/SimOTA outputs come from
`YOLOXLoss.__call__`, `YOLOXDecoder.__call__` and `postprocess` themselves; SimOTA's per-image
results are captured by wrapping `dynamic_k_matching` in the reference module's namespace while
`YOLOXLoss.__call__` (training mode) runs its own loop (yolox_loss.py:54-139).

Inputs are regenerated from seeds by pl_yolo_b200.synth (sha256 stored for a guard); only the
reference's OUTPUTS (and the decoded boxes SimOTA consumed) are stored.
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
REF = os.environ.get("PLYOLO_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import torch  # noqa: E402
import torchvision  # noqa: E402

from models.evaluators.postprocess import postprocess as ref_postprocess  # noqa: E402
from models.layers.losses.iou_loss import bboxes_iou as ref_bboxes_iou  # noqa: E402
from models.losses.yolox import yolox_loss as ref_loss_mod  # noqa: E402
from models.losses.yolox.yolox_decoder import YOLOXDecoder  # noqa: E402
from models.losses.yolox.yolox_loss import YOLOXLoss  # noqa: E402
from pl_yolo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
STRIDES = [8, 16, 32]
torch.manual_seed(0)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def save(name, meta, **arrays):
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items() if k != "meta"})


def pad_dets(outs, max_det=300):
    B = len(outs)
    d = np.zeros((B, max_det, 6), np.float32)
    c = np.zeros(B, np.int32)
    for i, o in enumerate(outs):
        if o is not None:
            c[i] = o.shape[0]
            d[i, : o.shape[0]] = o.numpy()
    return d, c


# ------------------------------------------------------------------------------- decode
def gen_decode(name, B, size, seed, mode="clustered", C=80, stride_rows=1):
    heads = synth.make_heads(B, size, C, seed, mode=mode)
    loss = YOLOXLoss(C, STRIDES)  # training mode
    preds_t, ori, xs, ys, es = loss.decode([T(h).clone() for h in heads])
    loss.eval()
    preds_e = loss([T(h).clone() for h in heads], torch.zeros(B, 1, 5))
    dec = YOLOXDecoder(C, STRIDES)([T(h).clone() for h in heads])
    assert torch.equal(dec, preds_e)
    rows = np.arange(0, preds_e.shape[1], stride_rows)
    meta = dict(kind="decode", B=B, size=size, C=C, seed=seed, mode=mode, strides=STRIDES,
                stride_rows=stride_rows, sha_heads=synth.digest(*heads))
    save(name, meta,
         train_boxes=preds_t[..., :4].numpy(), ori_boxes_equal_raw=np.array([1]),
         ori_rows=ori[:, rows].numpy(),
         train_rows=preds_t[:, rows].numpy(), eval_rows=preds_e[:, rows].numpy(),
         eval_boxes=preds_e[..., :4].numpy(), eval_obj=preds_e[..., 4].numpy(),
         x_shifts=xs.numpy(), y_shifts=ys.numpy(), expanded_strides=es.numpy(), rows=rows)


# ------------------------------------------------------------------------------- postprocess
def gen_post(name, preds, meta, conf, nms, agnostic=False):
    p = T(preds)
    outs = ref_postprocess(p.clone(), conf, nms, agnostic)
    d, c = pad_dets(outs)
    # the coordinate-trick branch evaluated with the CPU kernel (pins the oracle's trick path for Nk > 1000)
    dt = np.zeros_like(d)
    ct = np.zeros_like(c)
    ncand = np.zeros_like(c)
    for i in range(p.shape[0]):
        ip = p[i]
        cc, cp = torch.max(ip[:, 5:], 1, keepdim=True)
        cf = ip[:, 4] * cc.squeeze(1)
        det = torch.cat((ip[:, :4], cf.unsqueeze(-1), cp.float()), 1)[cf >= conf][:10000]
        ncand[i] = det.shape[0]
        if det.shape[0] == 0:
            continue
        if agnostic:
            k = torchvision.ops.nms(det[:, :4], det[:, 4], nms)
        else:
            k = torchvision.ops.boxes._batched_nms_coordinate_trick(det[:, :4], det[:, 4], det[:, 5], nms)
        o = det[k][:300]
        ct[i] = o.shape[0]
        dt[i, : o.shape[0]] = o.numpy()
    meta = dict(meta, kind="postprocess", conf=conf, nms=nms, agnostic=agnostic, sha_preds=synth.digest(preds))
    save(name, meta, dets=d, counts=c, dets_trick=dt, counts_trick=ct, n_cand=ncand)


def post_cases():
    for name, B, A, seed, kw in [
        ("post_small", 4, 525, 11, {}),
        ("post_640", 2, 8400, 12, {}),
        ("post_dense", 1, 8400, 13, dict(p_obj=0.9)),
        ("post_1280_trunc", 1, 33600, 14, dict(p_obj=0.6, size=1280.0, n_clusters=40)),
    ]:
        preds = synth.make_eval_preds(B, A, 80, seed, **kw)
        m = dict(gen="make_eval_preds", B=B, A=A, seed=seed, kw=kw)
        gen_post(name + "_c001_n065", preds, m, 0.01, 0.65)
        if name in ("post_small", "post_640"):
            gen_post(name + "_c05_n045", preds, m, 0.5, 0.45)
            gen_post(name + "_agn", preds, m, 0.01, 0.65, agnostic=True)
    # edge cases built by hand (stored as raw inputs: tiny)
    rng = np.random.default_rng(5)
    A = 64
    e = np.zeros((6, A, 85), np.float32)
    # 0: nothing passes
    e[0, :, 4] = 0.001
    # 1: identical boxes, identical scores, same class -> one survivor (lowest index)
    e[1, :, :4] = [10, 10, 50, 50]; e[1, :, 4] = 0.5; e[1, :, 5 + 3] = 0.5
    # 2: identical scores, distinct far-apart boxes, class ties in argmax (all classes equal -> class 0)
    for a in range(A):
        e[2, a, :4] = [a * 9, 0, a * 9 + 8, 8]
    e[2, :, 4] = 0.9; e[2, :, 5:] = 0.25
    # 3: saturated scores (1.0) + negative corners reaching into the neighbouring class offset
    e[3, :, :4] = rng.uniform(-40, 600, (A, 4)); e[3, :, 2:4] = e[3, :, :2] + rng.uniform(1, 300, (A, 2))
    e[3, :, 4] = 1.0; e[3, np.arange(A), 5 + rng.integers(0, 80, A)] = 1.0
    e[3, :8, :4] = [-30, -30, 20, 20]; e[3, 8:16, :4] = [560, 560, 640, 640]
    # 4: single candidate
    e[4, 17, :4] = [1, 2, 3, 4]; e[4, 17, 4] = 0.3; e[4, 17, 5 + 79] = 0.9
    # 5: heavy overlap chains, two classes
    base = rng.uniform(100, 300, (A, 2))
    e[5, :, :2] = base; e[5, :, 2:4] = base + 80
    e[5, :, 4] = rng.uniform(0.2, 1, A); e[5, :, 5 + 1] = 0.9 * (np.arange(A) % 2); e[5, :, 5 + 2] = 0.9 * ((np.arange(A) + 1) % 2)
    gen_post("post_edges", e, dict(gen="raw", B=6, A=A), 0.01, 0.65)
    np.save(os.path.join(OUT, "post_edges_input.npy"), e)


# ------------------------------------------------------------------------------- SimOTA
class Recorder:
    def __init__(self):
        self.calls = []
        self.orig = ref_loss_mod.dynamic_k_matching

    def __call__(self, fg_mask, cost, pair_wise_ious, gt_classes, num_gt):
        n_cand = int(fg_mask.sum())
        k = min(10, pair_wise_ious.size(1))
        srt, _ = pair_wise_ious.sort(descending=True)
        dyn = torch.clamp(srt[:, :k].sum(1).int(), min=1)
        out = self.orig(fg_mask, cost, pair_wise_ious, gt_classes, num_gt)
        fg, num_fg, mgi, gmc, pious = out
        self.calls.append(dict(fg=fg.clone().numpy(), num_fg=int(num_fg), mgi=mgi.numpy().copy(),
                               gmc=gmc.numpy().copy(), pious=pious.numpy().copy(), n_cand=n_cand,
                               dyn=dyn.numpy().copy(), cost_min=float(cost.min()) if cost.numel() else 0.0))
        return out


def run_ref_simota(heads, labels, C=80):
    rec = Recorder()
    ref_loss_mod.dynamic_k_matching = rec
    try:
        loss = YOLOXLoss(C, STRIDES)
        captured = {}
        orig_decode = loss.decode

        def decode_spy(inputs):
            r = orig_decode(inputs)
            captured["preds"] = r[0].clone()
            return r

        loss.decode = decode_spy
        out = loss([T(h).clone() for h in heads], T(labels))
    finally:
        ref_loss_mod.dynamic_k_matching = rec.orig
    preds = captured["preds"]
    B, A = preds.shape[:2]
    nlabel = (T(labels).sum(dim=2) > 0).sum(dim=1).numpy()
    fg = np.zeros((B, A), np.uint8)
    mg = np.full((B, A), -1, np.int32)
    mi = np.zeros((B, A), np.float32)
    nfg = np.zeros(B, np.int32)
    ncand = np.zeros(B, np.int32)
    dyn = np.zeros((B, labels.shape[1]), np.int32)
    it = iter(rec.calls)
    for b in range(B):
        if nlabel[b] == 0:
            continue
        c = next(it)
        fg[b] = c["fg"]
        idx = np.nonzero(c["fg"])[0]
        mg[b, idx] = c["mgi"]
        mi[b, idx] = c["pious"]
        nfg[b] = c["num_fg"]
        ncand[b] = c["n_cand"]
        dyn[b, : nlabel[b]] = c["dyn"]
        assert np.array_equal(c["gmc"], labels[b, c["mgi"], 0])
    losses = {k: float(v) for k, v in out.items()}
    return preds.numpy(), dict(fg_mask=fg, matched_gt=mg, matched_iou=mi, num_fg=nfg, num_gt=nlabel.astype(np.int32),
                               n_cand=ncand, dyn_k=dyn), losses


def gen_simota(name, B, size, seed_h, seed_l, max_labels, min_gt=1, max_gt=None, labels=None, C=80,
               objects=12, note=""):
    heads = synth.make_heads(B, size, C, seed_h, objects_per_image=objects)
    if labels is None:
        labels = synth.make_labels(B, size, max_labels, C, seed_l, min_gt, max_gt)
        lab_meta = dict(gen="make_labels", seed=seed_l, max_labels=max_labels, min_gt=min_gt, max_gt=max_gt)
    else:
        lab_meta = dict(gen="raw")
    preds, res, losses = run_ref_simota(heads, labels, C)
    meta = dict(kind="simota", B=B, size=size, C=C, seed_heads=seed_h, objects=objects, labels=lab_meta, strides=STRIDES,
                sha_heads=synth.digest(*heads), sha_labels=synth.digest(labels), losses=losses, note=note)
    extra = {} if lab_meta["gen"] != "raw" else {"labels": labels}
    save(name, meta, ref_boxes=preds[..., :4], **res, **extra)


def simota_cases():
    gen_simota("simota_640_b4", 4, 640, 21, 22, 120)
    gen_simota("simota_640_dense", 2, 640, 23, 24, 120, min_gt=100, max_gt=120)
    gen_simota("simota_320_b8", 8, 320, 25, 26, 60)
    gen_simota("simota_160_b8", 8, 160, 27, 28, 30)
    # hand-made edge cases on a 160^2 grid (A = 525)
    S = 160
    L = np.zeros((8, 12, 5), np.float32)
    # 0: no GT at all
    # 1: GT centre far outside the image -> no candidate (Q6)
    L[1, 0] = [3, 900, 900, 20, 20]
    # 2: one tiny GT between cell centres (in-box empty, centre prior only)
    L[2, 0] = [5, 80.0, 80.0, 2, 2]
    # 3: duplicate GTs (cost ties across g -> argmin picks the lowest g)
    L[3, 0] = [7, 60, 70, 50, 40]; L[3, 1] = [7, 60, 70, 50, 40]; L[3, 2] = [9, 64, 72, 48, 44]
    # 4: GT at the corner so only 1-3 anchors are candidates (Q3: k >= Nc-1 takes all)
    L[4, 0] = [1, 2.0, 2.0, 3.0, 3.0]
    # 5: many overlapping GTs of different classes
    for g in range(12):
        L[5, g] = [g * 5 % 80, 60 + 3 * g, 80 - 2 * g, 70 + g, 50 + 2 * g]
    # 6: GT covering the whole image + a small one
    L[6, 0] = [0, 80, 80, 160, 160]; L[6, 1] = [79, 40, 40, 10, 10]
    # 7: a valid row after an all-zero row (reference counts rows, then takes the FIRST count rows)
    L[7, 0] = [2, 50, 50, 30, 30]; L[7, 2] = [4, 100, 100, 40, 40]
    gen_simota("simota_edges", 8, S, 29, 0, 12, labels=L)


# ------------------------------------------------------------------------------- loss tail gradients (N2)
def gen_lossgrad(name, B, size, seed_h, seed_l, max_labels, C=80, use_l1=False):
    """The real reference's training loss and its autograd gradients with respect to the head outputs
    (YOLOXLoss.__call__ in training mode, yolox_loss.py:20-173; decode mutates its inputs in place, quirk Q1, so the
    leaves are cloned first)."""
    heads = synth.make_heads(B, size, C, seed_h)
    labels = synth.make_labels(B, size, max_labels, C, seed_l)
    leaves = [T(h).clone().requires_grad_(True) for h in heads]
    out = YOLOXLoss(C, STRIDES, use_l1=use_l1)([x.clone() for x in leaves], T(labels))
    out["loss"].backward()
    losses = {k: float(v) for k, v in out.items()}
    save(name, dict(kind="lossgrad", B=B, size=size, C=C, strides=STRIDES, seed_heads=seed_h, use_l1=bool(use_l1),
                    labels=dict(gen="synth", seed=seed_l, max_labels=max_labels, min_gt=1, max_gt=None),
                    sha_heads=synth.digest(*heads), sha_labels=synth.digest(labels), losses=losses),
         **{"grad%d" % l: x.grad.numpy() for l, x in enumerate(leaves)})


if __name__ == "__main__":
    which = sys.argv[1:] or ["decode", "post", "simota"]
    if "decode" in which:
        gen_decode("decode_160_b2", 2, 160, 1)
        gen_decode("decode_640_b1", 1, 640, 2, stride_rows=16)
        gen_decode("decode_320_sparse", 2, 320, 3, mode="sparse", stride_rows=4)
    if "post" in which:
        post_cases()
    if "simota" in which:
        simota_cases()
    if "lossgrad" in which:  # added with N2; not part of the default set so the older fixtures stay byte-identical
        gen_lossgrad("lossgrad_160_b2", 2, 160, 51, 52, 12)
        gen_lossgrad("lossgrad_320_b2", 2, 320, 53, 54, 40)
    if "lossgrad_l1" in which:  # use_l1=True (the reference's own configs never enable it)
        gen_lossgrad("lossgrad_l1_160_b2", 2, 160, 55, 56, 12, use_l1=True)
        gen_lossgrad("lossgrad_l1_320_b3", 3, 320, 57, 58, 30, use_l1=True)
