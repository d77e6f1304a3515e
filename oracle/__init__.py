"""ctypes front-end of the CPU oracle (oracle/plyolo_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under pl_yolo_b200/ may import this package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libplyolo_oracle.so")

FLAVOR_CUDA = 0  # the reference on CUDA tensors (coordinate trick, fused Sb, ATen CUDA sum tree)
NMS_RULE_CPU, IOU_NOFMA, THR_F64 = 1, 2, 4
FLAVOR_CPU = 7   # the reference on CPU tensors (per-class NMS when Nk > 1000, no FMA, double threshold)

_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "plyolo_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.plyolo_oracle_aten_cuda_sum.restype = ctypes.c_float
        _lib.plyolo_oracle_aten_cuda_sum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_long]
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ivec(v: Sequence[int]):
    return (ctypes.c_int * len(v))(*[int(x) for x in v])


def aten_cuda_sum(e: np.ndarray, n_out: int) -> float:
    e = _f32(e)
    return float(lib().plyolo_oracle_aten_cuda_sum(_p(e), int(e.size), int(n_out)))


def decode(heads: Sequence[np.ndarray], strides: Sequence[int], inference: bool, want_ori: bool = True):
    """-> (preds [B,A,5+C], ori_boxes [B,A,4] | None)   (yolox_loss.py:175-228, :25-36)"""
    heads = [_f32(h) for h in heads]
    B, ch = heads[0].shape[:2]
    hs = [h.shape[2] for h in heads]
    ws = [h.shape[3] for h in heads]
    A = sum(a * b for a, b in zip(hs, ws))
    preds = np.empty((B, A, ch), np.float32)
    ori = np.empty((B, A, 4), np.float32) if want_ori else None
    ptrs = (ctypes.c_void_p * len(heads))(*[h.ctypes.data for h in heads])
    rc = lib().plyolo_oracle_decode(ptrs, _ivec(hs), _ivec(ws), _ivec(strides), len(heads), B, ch - 5,
                                    _p(preds), _p(ori), int(bool(inference)))
    assert rc == 0
    return preds, ori


def postprocess(preds: np.ndarray, conf_thre: float = 0.7, nms_thre: float = 0.45, class_agnostic: bool = False,
                max_nms: int = 10000, max_det: int = 300, flavor: int = FLAVOR_CUDA):
    """-> dict(dets [B,max_det,6], counts [B], keep_idx [B,max_det] anchor ids, n_cand [B])
    (models/evaluators/postprocess.py:7-48)"""
    preds = _f32(preds)
    B, A, ch = preds.shape
    dets = np.zeros((B, max_det, 6), np.float32)
    counts = np.zeros(B, np.int32)
    keep = np.full((B, max_det), -1, np.int32)
    ncand = np.zeros(B, np.int32)
    fn = lib().plyolo_oracle_postprocess
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                   ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                   ctypes.c_void_p, ctypes.c_void_p]
    rc = fn(_p(preds), B, A, ch - 5, float(conf_thre), float(nms_thre), int(class_agnostic), int(max_nms),
            int(max_det), int(flavor), _p(dets), _p(counts), _p(keep), _p(ncand))
    assert rc == 0
    return {"dets": dets, "counts": counts, "keep_idx": keep, "n_cand": ncand}


def simota(preds: np.ndarray, labels: np.ndarray, hw: Sequence[Sequence[int]], strides: Sequence[int],
           sum_order: int = 0):
    """-> dict(fg_mask [B,A] u8, matched_gt [B,A] i32, matched_iou [B,A] f32, num_fg [B], num_gt [B],
               dyn_k [B,Lmax], n_cand [B])     (yolox_loss.py:43-118, :231-370; iou_loss.py:391-414)"""
    preds = _f32(preds)
    labels = _f32(labels)
    B, A, ch = preds.shape
    Lmax = labels.shape[1]
    fg = np.zeros((B, A), np.uint8)
    mg = np.zeros((B, A), np.int32)
    mi = np.zeros((B, A), np.float32)
    nfg = np.zeros(B, np.int32)
    ngt = np.zeros(B, np.int32)
    dk = np.zeros((B, Lmax), np.int32)
    nc = np.zeros(B, np.int32)
    hs = [int(h) for h, _ in hw]
    ws = [int(w) for _, w in hw]
    fn = lib().plyolo_oracle_simota
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 3 + \
                  [ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 7
    rc = fn(_p(preds), _p(labels), B, A, ch - 5, Lmax, _ivec(hs), _ivec(ws), _ivec(strides), len(hs),
            int(sum_order), _p(fg), _p(mg), _p(mi), _p(nfg), _p(ngt), _p(dk), _p(nc))
    assert rc == 0, "oracle simota: anchor count does not match the level shapes"
    return {"fg_mask": fg, "matched_gt": mg, "matched_iou": mi, "num_fg": nfg, "num_gt": ngt,
            "dyn_k": dk, "n_cand": nc}


def bboxes_iou(a: np.ndarray, b: np.ndarray, xyxy: bool = True) -> np.ndarray:
    """iou_loss.py:391-414; raises IndexError on a last dim != 4 like the reference."""
    a = _f32(a)
    b = _f32(b)
    if a.shape[1] != 4 or b.shape[1] != 4:
        raise IndexError
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().plyolo_oracle_bboxes_iou(_p(a), a.shape[0], _p(b), b.shape[0], int(xyxy), _p(out))
    return out
