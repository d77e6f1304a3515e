#!/usr/bin/env python
"""Goldens for the NMS call sites of the sibling heads, produced by RUNNING THE REAL REFERENCE classes
(CPU) in the build container: YOLOv3Decoder.__call__ (models/losses/yolov3/yolov3_decoder.py) and
YOLOv5Decoder.__call__ (models/losses/yolov5/yolov5_decoder.py).  The decoded `predictions` tensor each class feeds
to its NMS loop is captured (v3: the torch.cat of :61, cloned before the in-place corner conversion; v5: decode() is
replaced by the stored tensor) and stored with the class's own output.  `python oracle/gen_golden_siblings.py`"""
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

from models.losses.yolov3.yolov3_decoder import YOLOv3Decoder  # noqa: E402
from models.losses.yolov5.yolov5_decoder import YOLOv5Decoder  # noqa: E402
from pl_yolo_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
ANCHORS = [[[10, 13], [16, 30], [33, 23]], [[30, 61], [62, 45], [59, 119]], [[116, 90], [156, 198], [373, 326]]]
STRIDES = [8, 16, 32]


def save(name, meta, **arrays):
    arrays["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote", name, {k: v.shape for k, v in arrays.items() if k != "meta"})


def pad(outs, max_det):
    d = np.zeros((len(outs), max_det, 7), np.float32)
    c = np.zeros(len(outs), np.int32)
    for i, o in enumerate(outs):
        if o is not None and o.shape[0]:
            c[i] = o.shape[0]
            d[i, : o.shape[0]] = o.numpy()
    return d, c


def gen_v3(name, B, size, C, conf, nms, seed, max_nms=10000, max_det=300):
    rng = np.random.default_rng(seed)
    heads = []
    for s in STRIDES:
        h = size // s
        m = rng.normal(0, 1.2, (B, 3 * (5 + C), h, h)).astype(np.float32)
        m[:, 4::(5 + C)] -= 0.5
        # a few cells repeat their neighbour's outputs (overlapping boxes of one class: NMS has work to do)
        if h > 2:
            m[:, :, 1:, :] = np.where(rng.uniform(0, 1, (B, 1, h - 1, h)) < 0.35, m[:, :, :-1, :] + rng.normal(0, 0.05, (B, 3 * (5 + C), h - 1, h)), m[:, :, 1:, :]).astype(np.float32)
        heads.append(torch.from_numpy(m))
    dec = YOLOv3Decoder(C, ANCHORS, STRIDES)
    dec.max_nms, dec.max_det, dec.time_limit = max_nms, max_det, 1e9
    captured = []
    real_cat = torch.cat

    def spy(tensors, dim=0, **kw):
        out = real_cat(tensors, dim, **kw)
        if dim == 1 and out.dim() == 3 and out.shape[2] == 5 + C and not captured:
            captured.append(out.clone())
        return out

    torch.cat = spy
    try:
        outs = dec([h.clone() for h in heads], conf, nms)
    finally:
        torch.cat = real_cat
    d, c = pad(outs, max_det)
    save(name, dict(kind="sib_v3", C=C, conf=conf, nms=nms, max_nms=max_nms, max_det=max_det), predictions=captured[0].numpy(), dets=d, counts=c)


def gen_v5(name, B, N, C, conf, nms, seed, agnostic=False, max_nms_note=30000):
    p = synth.make_eval_preds(B, N, C, seed, size=640.0, n_clusters=10, p_obj=0.35)
    x = p.copy()
    x[..., 0] = (p[..., 0] + p[..., 2]) / 2
    x[..., 1] = (p[..., 1] + p[..., 3]) / 2
    x[..., 2] = p[..., 2] - p[..., 0]
    x[..., 3] = p[..., 3] - p[..., 1]
    x = np.ascontiguousarray(x, np.float32)
    dec = YOLOv5Decoder(C, ANCHORS, STRIDES)
    dec.decode = lambda inputs: torch.from_numpy(x.copy())
    outs = dec(None, conf, nms, multi_label=False, agnostic=agnostic)
    d, c = pad(outs, 300)
    save(name, dict(kind="sib_v5", C=C, conf=conf, nms=nms, agnostic=agnostic, max_nms=30000, max_det=300), predictions=x, dets=d, counts=c)


if __name__ == "__main__":
    gen_v3("sib_v3_a", 3, 128, 20, 0.3, 0.45, 1)
    gen_v3("sib_v3_b", 3, 128, 20, 0.62, 0.3, 5)
    gen_v3("sib_v3_trunc", 2, 128, 20, 0.2, 0.5, 2, max_nms=60, max_det=25)   # n > max_nms: the 60 best by conf survive
    gen_v5("sib_v5_a", 3, 1500, 80, 0.25, 0.45, 3)
    gen_v5("sib_v5_agnostic", 2, 900, 80, 0.1, 0.6, 4, agnostic=True)
