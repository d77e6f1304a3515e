"""Seeded synthetic inputs of the shapes BASELINE.json names.

numpy-only (PCG64 streams are reproducible across machines for a given numpy), so the
same seed yields bit-identical head maps / labels in the build container and on the GPU
box; tests/golden/ stores a sha256 of each regenerated input next to the reference's
outputs and the tests check it before trusting a fixture.

Layouts follow the reference:
  head maps  : list of [B, 5+C, H_l, W_l] fp32, channels reg(4) | obj(1) | cls(C)
               (models/heads/decoupled_head.py:93)
  labels     : [B, Lmax, 5] fp32 rows (cls, cx, cy, w, h), valid rows first, zero padded
               (models/data/data_augments.py:41-47)
"""
from __future__ import annotations

import hashlib
from typing import List, Sequence, Tuple

import numpy as np

STRIDES = (8, 16, 32)


def level_shapes(size: int, strides: Sequence[int] = STRIDES) -> List[Tuple[int, int]]:
    return [(size // s, size // s) for s in strides]


def num_anchors(size: int, strides: Sequence[int] = STRIDES) -> int:
    return sum(h * w for h, w in level_shapes(size, strides))


def make_heads(batch: int, size: int = 640, num_classes: int = 80, seed: int = 0,
               objects_per_image: int = 12, mode: str = "clustered",
               strides: Sequence[int] = STRIDES) -> List[np.ndarray]:
    """"clustered": background noise plus `objects_per_image` objects whose cells predict
    consistent boxes (gives ~1900 candidates/img at conf 0.01 and 300 dets/img at 640^2);
    "sparse": background only with obj ~ N(-7,1) (a handful of candidates per image)."""
    rng = np.random.default_rng(seed)
    C = num_classes
    out = []
    if mode == "clustered":
        ocx = rng.uniform(0, size, (batch, objects_per_image))
        ocy = rng.uniform(0, size, (batch, objects_per_image))
        ow = (0.05 + 0.45 * rng.uniform(0, 1, (batch, objects_per_image))) * size
        oh = (0.05 + 0.45 * rng.uniform(0, 1, (batch, objects_per_image))) * size
        ocls = rng.integers(0, C, (batch, objects_per_image))
    for s in strides:
        H = W = size // s
        m = np.empty((batch, 5 + C, H, W), np.float32)
        m[:, 0:2] = rng.uniform(0, 1, (batch, 2, H, W))
        m[:, 2:4] = rng.normal(0.5, 0.5, (batch, 2, H, W))
        if mode == "sparse":
            m[:, 4] = rng.normal(-7.0, 1.0, (batch, H, W))
        else:
            m[:, 4] = rng.normal(-6.0, 1.5, (batch, H, W))
        m[:, 5:] = rng.normal(-3.0, 1.0, (batch, C, H, W))
        if mode == "clustered":
            gx = (np.arange(W, dtype=np.float64))[None, :].repeat(H, 0)
            gy = (np.arange(H, dtype=np.float64))[:, None].repeat(W, 1)
            ccx = (gx + 0.5) * s
            ccy = (gy + 0.5) * s
            for b in range(batch):
                for o in range(objects_per_image):
                    inside = (np.abs(ccx - ocx[b, o]) <= 0.25 * ow[b, o]) & (np.abs(ccy - ocy[b, o]) <= 0.25 * oh[b, o])
                    n = int(inside.sum())
                    if n == 0:
                        continue
                    m[b, 0][inside] = (ocx[b, o] / s - gx[inside]) + rng.normal(0, 0.05, n)
                    m[b, 1][inside] = (ocy[b, o] / s - gy[inside]) + rng.normal(0, 0.05, n)
                    m[b, 2][inside] = np.log(ow[b, o] / s) + rng.normal(0, 0.1, n)
                    m[b, 3][inside] = np.log(oh[b, o] / s) + rng.normal(0, 0.1, n)
                    m[b, 4][inside] = rng.normal(1.0, 1.5, n)
                    m[b, 5 + int(ocls[b, o])][inside] = rng.normal(2.0, 1.0, n)
        out.append(np.ascontiguousarray(m))
    return out


def make_labels(batch: int, size: int = 640, max_labels: int = 120, num_classes: int = 80,
                seed: int = 1, min_gt: int = 1, max_gt: int | None = None) -> np.ndarray:
    """G_b ~ U{min_gt..max_gt}; class ~ U{0..C-1}; wh = (0.02 + 0.4 U^2) S; centre ~ U(wh/2, S - wh/2)."""
    rng = np.random.default_rng(seed)
    max_gt = max_labels if max_gt is None else max_gt
    lab = np.zeros((batch, max_labels, 5), np.float32)
    for b in range(batch):
        g = int(rng.integers(min_gt, max_gt + 1)) if max_gt >= min_gt else 0
        if g == 0:
            continue
        cls = rng.integers(0, num_classes, g)
        w = (0.02 + 0.4 * rng.uniform(0, 1, g) ** 2) * size
        h = (0.02 + 0.4 * rng.uniform(0, 1, g) ** 2) * size
        cx = rng.uniform(w / 2, size - w / 2)
        cy = rng.uniform(h / 2, size - h / 2)
        lab[b, :g] = np.stack([cls, cx, cy, w, h], 1).astype(np.float32)
    return lab


def make_eval_preds(batch: int, anchors: int, num_classes: int = 80, seed: int = 0, size: float = 640.0,
                    n_clusters: int = 12, p_obj: float = 0.25) -> np.ndarray:
    """Already-decoded inference predictions [B, A, 5+C] = (x1,y1,x2,y2,obj,cls..) in (0,1),
    generated WITHOUT transcendentals so they can be regenerated bit-exactly anywhere:
    boxes jittered around `n_clusters` centres so that NMS has real work to do."""
    rng = np.random.default_rng(seed)
    C = num_classes
    p = np.empty((batch, anchors, 5 + C), np.float32)
    ccx = rng.uniform(0.1 * size, 0.9 * size, (batch, n_clusters))
    ccy = rng.uniform(0.1 * size, 0.9 * size, (batch, n_clusters))
    cw = rng.uniform(0.05 * size, 0.4 * size, (batch, n_clusters))
    chh = rng.uniform(0.05 * size, 0.4 * size, (batch, n_clusters))
    ccl = rng.integers(0, C, (batch, n_clusters))
    which = rng.integers(0, n_clusters, (batch, anchors))
    bi = np.arange(batch)[:, None]
    cx = ccx[bi, which] + rng.normal(0, 0.04 * size, (batch, anchors))
    cy = ccy[bi, which] + rng.normal(0, 0.04 * size, (batch, anchors))
    w = cw[bi, which] * rng.uniform(0.7, 1.3, (batch, anchors))
    h = chh[bi, which] * rng.uniform(0.7, 1.3, (batch, anchors))
    p[..., 0] = cx - w / 2
    p[..., 1] = cy - h / 2
    p[..., 2] = cx + w / 2
    p[..., 3] = cy + h / 2
    hot = rng.uniform(0, 1, (batch, anchors)) < p_obj
    p[..., 4] = np.where(hot, rng.uniform(0.05, 1.0, (batch, anchors)), rng.uniform(0, 0.02, (batch, anchors)))
    p[..., 5:] = rng.uniform(0, 0.08, (batch, anchors, C))
    top = np.where(rng.uniform(0, 1, (batch, anchors)) < 0.8, ccl[bi, which], rng.integers(0, C, (batch, anchors)))
    np.put_along_axis(p[..., 5:], top[..., None], rng.uniform(0.2, 1.0, (batch, anchors, 1)).astype(np.float32), axis=2)
    return p


def make_train_preds(heads: Sequence[np.ndarray], boxes: np.ndarray) -> np.ndarray:
    """Training-mode decode output [B,A,5+C]: given decoded (cx,cy,w,h) `boxes` [B,A,4] and the
    head maps (for the raw obj/cls logits, copied through unchanged by yolox_loss.py:210-227)."""
    B = heads[0].shape[0]
    ch = heads[0].shape[1]
    raw = np.concatenate([h.reshape(B, ch, -1).transpose(0, 2, 1) for h in heads], 1)
    out = np.ascontiguousarray(raw, dtype=np.float32).copy()
    out[..., :4] = boxes
    return out


def digest(*arrays: np.ndarray) -> str:
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode())
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()
