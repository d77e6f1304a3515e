#!/usr/bin/env python
"""Goldens for the stand-alone boundary functions, produced by RUNNING THE REAL REFERENCE (CPU) in the build container:
get_in_boxes_info / dynamic_k_matching (models/losses/yolox/yolox_loss.py:231-370) and format_outputs
(models/evaluators/postprocess.py:95-138).  Inputs are stored with the outputs (small).  `python oracle/gen_golden_boundary.py`"""
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

from models.evaluators.postprocess import format_outputs as ref_format_outputs  # noqa: E402
from models.losses.yolox.yolox_loss import dynamic_k_matching as ref_dynk  # noqa: E402
from models.losses.yolox.yolox_loss import get_in_boxes_info as ref_in_boxes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def save(name, meta, **arrays):
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items() if k != "meta"})


def grid(size, strides=(8, 16, 32)):
    xs, ys, es = [], [], []
    for s in strides:
        h = w = size // s
        yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        xs.append(xv.reshape(1, -1).float()); ys.append(yv.reshape(1, -1).float()); es.append(torch.full((1, h * w), float(s)))
    return torch.cat(xs, 1), torch.cat(ys, 1), torch.cat(es, 1)


def gen_in_boxes():
    xs, ys, es = grid(160)
    rng = np.random.default_rng(11)
    gt = np.array([[80, 80, 60, 40], [12, 12, 8, 8], [300, 300, 20, 20], [100.0, 60.0, 3.0, 3.0], [36.0, 36.0, 8.0, 8.0],
                   [0.0, 0.0, 50.0, 50.0], [159.0, 159.0, 100.0, 10.0], [44.0, 76.0, 0.0, 0.0]], np.float32)
    gt = np.concatenate([gt, np.stack([rng.uniform(0, 160, 12), rng.uniform(0, 160, 12), rng.uniform(2, 90, 12), rng.uniform(2, 90, 12)], 1).astype(np.float32)])
    G, A = gt.shape[0], xs.shape[1]
    fg, both = ref_in_boxes(torch.from_numpy(gt), es, xs, ys, A, G)
    save("ref_in_boxes_160", dict(kind="ref_in_boxes", size=160), gt=gt, x_shifts=xs.numpy(), y_shifts=ys.numpy(), expanded_strides=es.numpy(),
         fg_mask=fg.numpy(), both=both.numpy())


def gen_dynk():
    rng = np.random.default_rng(5)
    cases = {}
    for name, (G, Nc, A) in {"a": (5, 40, 90), "b": (3, 3, 10), "c": (4, 2, 6), "d": (12, 300, 525), "e": (1, 1, 4), "f": (6, 33, 64)}.items():
        ious = rng.uniform(0, 1, (G, Nc)).astype(np.float32) ** 2
        if name == "d":
            ious[3] *= 0.02   # dynamic k = 1
            ious[5] = np.clip(ious[5] * 3, 0, 0.99)
        cost = rng.uniform(0.5, 20, (G, Nc)).astype(np.float32)
        if G > 1:  # provoke conflicts: two GT rows that prefer the same anchors
            cost[1] = cost[0] + rng.uniform(-0.01, 0.01, Nc).astype(np.float32)
            ious[1] = ious[0]
        cost[rng.uniform(0, 1, (G, Nc)) < 0.3] += 100000.0
        fg = np.zeros(A, bool)
        fg[rng.choice(A, Nc, replace=False)] = True
        cls = rng.integers(0, 80, G).astype(np.float32)
        fgt = torch.from_numpy(fg.copy())
        out = ref_dynk(fgt, torch.from_numpy(cost), torch.from_numpy(ious), torch.from_numpy(cls), G)
        assert out[0] is fgt
        cases.update({name + "_cost": cost, name + "_ious": ious, name + "_fg_in": fg, name + "_cls": cls, name + "_fg_out": fgt.numpy(),
                      name + "_num_fg": np.array(int(out[1])), name + "_gt": out[2].numpy(), name + "_mcls": out[3].numpy(), name + "_iou": out[4].numpy()})
    save("ref_dynk", dict(kind="ref_dynk", cases=["a", "b", "c", "d", "e", "f"]), **cases)


def gen_format():
    rng = np.random.default_rng(9)
    B, n_cls = 5, 20
    counts = [300, 0, 17, 1, 120]
    outs, raw = [], []
    for n in counts:
        if n == 0:
            outs.append(None); raw.append(np.zeros((0, 6), np.float32)); continue
        x1 = rng.uniform(-20, 600, n); y1 = rng.uniform(-20, 600, n)
        d = np.stack([x1, y1, x1 + rng.uniform(1, 300, n), y1 + rng.uniform(1, 300, n), np.sort(rng.uniform(0.01, 1, n))[::-1],
                      rng.integers(0, n_cls, n)], 1).astype(np.float32)
        raw.append(d.copy()); outs.append(torch.from_numpy(d.copy()))
    ids = [11, 12, 13, 14, 15]
    hws = [[480, 375, 640, 1080, 333], [640, 500, 480, 1920, 500]]   # (heights, widths) as the dataloader batches them
    val_size = (640, 640)
    class_ids = list(range(1, n_cls + 1))
    json_list, det_list = ref_format_outputs(outs, ids, hws, val_size, class_ids, None)
    arrays = {"in_%d" % i: raw[i] for i in range(B)}
    arrays["json_image_id"] = np.array([j["image_id"] for j in json_list], np.int64)
    arrays["json_category_id"] = np.array([j["category_id"] for j in json_list], np.int64)
    arrays["json_bbox"] = np.array([j["bbox"] for j in json_list], np.float64)
    arrays["json_score"] = np.array([j["score"] for j in json_list], np.float64)
    for i in range(B):
        arrays["scaled_%d" % i] = outs[i].numpy() if outs[i] is not None else np.zeros((0, 6), np.float32)  # the in-place `bboxes /= scale`
        for c in range(n_cls):
            arrays["det_%d_%d" % (i, c)] = np.asarray(det_list[i][c], np.float64)
    save("ref_format_outputs", dict(kind="ref_format", ids=ids, hws=hws, val_size=list(val_size), class_ids=class_ids, counts=counts), **arrays)


def gen_voc():
    """tpfp_default / average_precision of models/evaluators/eval_voc.py (terminaltables, absent here, is only used to
    print the summary table: stubbed for the import)."""
    import types
    sys.modules.setdefault("terminaltables", types.SimpleNamespace(AsciiTable=object))
    from models.evaluators.eval_voc import average_precision, tpfp_default
    rng = np.random.default_rng(21)
    B, C, max_det, Gmax = 6, 20, 300, 40
    dets = np.zeros((B, max_det, 6), np.float32); counts = np.zeros(B, np.int32)
    gts = np.zeros((B, Gmax, 5), np.float32); gcnt = np.zeros(B, np.int32)
    for b in range(B):
        k = int(rng.integers(0, Gmax + 1)) if b != 2 else 0
        gx = rng.uniform(0, 500, k); gy = rng.uniform(0, 400, k); gw = rng.uniform(10, 200, k); gh = rng.uniform(10, 200, k)
        gc = rng.integers(0, C, k)
        gts[b, :k] = np.stack([gx, gy, gx + gw, gy + gh, gc], 1); gcnt[b] = k
        n = int(rng.integers(0, max_det + 1)) if b != 4 else 0
        rows = []
        for i in range(n):
            if k and rng.uniform() < 0.6:
                j = int(rng.integers(0, k)); jit = rng.normal(0, 0.12, 4) * np.array([gw[j], gh[j], gw[j], gh[j]])
                box = np.array([gx[j], gy[j], gx[j] + gw[j], gy[j] + gh[j]]) + jit
                c = gc[j] if rng.uniform() < 0.85 else rng.integers(0, C)
            else:
                x = rng.uniform(0, 500); y = rng.uniform(0, 400)
                box = np.array([x, y, x + rng.uniform(5, 150), y + rng.uniform(5, 150)]); c = rng.integers(0, C)
            rows.append(list(box) + [0, c])
        rows = np.array(rows, np.float32).reshape(-1, 6)
        rows[:, 4] = np.sort(rng.uniform(0.01, 1, n).astype(np.float32))[::-1]
        dets[b, :n] = rows; counts[b] = n
    tp = np.zeros((B, max_det), np.uint8)
    aps = np.zeros(C, np.float32); ngts = np.zeros(C, np.int32)
    for c in range(C):
        tps, fps, scs = [], [], []
        for b in range(B):
            n = counts[b]; m = dets[b, :n, 5].astype(int) == c
            d = dets[b, :n][m][:, :5]; g = gts[b, :gcnt[b]][gts[b, :gcnt[b], 4].astype(int) == c][:, :4]
            t, f = tpfp_default(d, g, 0.5)                                   # eval_voc.py:75
            tp[b, np.nonzero(m)[0]] = t.astype(np.uint8)
            assert np.array_equal(f, 1 - t)
            tps.append(t); fps.append(f); scs.append(d[:, 4]); ngts[c] += g.shape[0]
        si = np.argsort(-np.hstack(scs))                                     # :38-50, VOCEvaluator's own arithmetic
        t = np.cumsum(np.hstack(tps)[si], axis=0); f = np.cumsum(np.hstack(fps)[si], axis=0)
        eps = np.finfo(np.float32).eps; num = np.zeros(1, dtype=int); num[0] = ngts[c]
        rec = t / np.maximum(num, eps); pre = t / np.maximum(t + f, eps)
        aps[c] = average_precision(rec, pre, "area") if t.shape[0] else 0    # :108
    save("ref_voc_tpfp", dict(kind="ref_voc", B=B, C=C, max_det=max_det, Gmax=Gmax, iou_thr=0.5), dets=dets, counts=counts, gts=gts,
         gt_counts=gcnt, tp=tp, ap=aps, num_gts=ngts)


if __name__ == "__main__":
    gen_voc()
    gen_in_boxes()
    gen_dynk()
    gen_format()
