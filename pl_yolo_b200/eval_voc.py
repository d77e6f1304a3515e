"""VOC evaluator statistics on the device — drop-in pieces of models/evaluators/eval_voc.py.

`tpfp_default` keeps the reference signature (numpy in, numpy out, one (image, class) pair); `voc_tpfp_dense` is the
batch form the validation loop should use: it takes the dense detections that `postprocess` / `format_outputs` already
hold on the device plus padded ground truth, and returns the TP flags of every detection of every image and class in
one launch (the reference: a multiprocessing.Pool(8) over numpy arrays per class, eval_voc.py:18-31).  `voc_ap` then
follows VOCEvaluator's arithmetic (:37-58) and `average_precision` (:108-150, mode 'area') on the gathered arrays.
Across GPUs the TP flags travel with the padded detections (same all-gather) and `num_gts` is all-reduced
(`gather_voc_stats`)."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops


def voc_tpfp_dense(dets: torch.Tensor, counts: torch.Tensor, gts: torch.Tensor, gt_counts: torch.Tensor, iou_thr: float = 0.5,
                   num_classes: int = 20) -> Tuple[torch.Tensor, torch.Tensor]:
    """dets [B,max_det,6] (x1,y1,x2,y2,score,class) score-descending per image, counts [B] i32; gts [B,Gmax,5]
    (x1,y1,x2,y2,class), gt_counts [B] i32 -> (tp [B,max_det] bool, num_gts [C] i32).  fp = valid & ~tp."""
    d = ops._check_cuda_f32(dets, "dets")
    g = ops._check_cuda_f32(gts, "gts")
    if d.dim() != 3 or d.shape[2] != 6 or g.dim() != 3 or g.shape[2] != 5 or g.shape[0] != d.shape[0]:
        raise ValueError("dets must be [B,max_det,6] and gts [B,Gmax,5]")
    if counts.dtype != torch.int32 or gt_counts.dtype != torch.int32:
        raise TypeError("counts / gt_counts must be int32")
    B, max_det, _ = d.shape
    dev = d.device
    tp = torch.empty((B, max_det), dtype=torch.uint8, device=dev)
    num_gts = torch.empty((num_classes,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_voc_tpfp_f32(d.data_ptr(), counts.contiguous().data_ptr(), B, max_det, g.data_ptr(),
                                            gt_counts.contiguous().data_ptr(), g.shape[1], float(iou_thr), num_classes,
                                            tp.data_ptr(), num_gts.data_ptr(), ops._stream_ptr(dev))
    _lib.check(rc, "plyolo_voc_tpfp_f32")
    return tp.view(torch.bool), num_gts


def tpfp_default(det_bboxes: np.ndarray, gt_bboxes: np.ndarray, iou_thr: float = 0.5):
    """eval_voc.py:75 — det_bboxes [n,5] (x1,y1,x2,y2,score), gt_bboxes [k,4] of ONE image and class
    -> (tp [n] float32, fp [n] float32).  Convenience wrapper (one launch per call): use voc_tpfp_dense in the loop."""
    n, k = det_bboxes.shape[0], gt_bboxes.shape[0]
    tp = np.zeros(n, dtype=np.float32)
    fp = np.zeros(n, dtype=np.float32)
    if n == 0:
        return tp, fp
    if k == 0:
        fp[...] = 1
        return tp, fp
    order = np.argsort(-det_bboxes[:, -1], kind="stable")  # the kernel expects score-descending rows
    dev = torch.device("cuda", torch.cuda.current_device())
    d = torch.zeros((1, n, 6), dtype=torch.float32)
    d[0, :, :5] = torch.from_numpy(np.ascontiguousarray(det_bboxes[order, :5], dtype=np.float32))
    g = torch.zeros((1, k, 5), dtype=torch.float32)
    g[0, :, :4] = torch.from_numpy(np.ascontiguousarray(gt_bboxes[:, :4], dtype=np.float32))
    t, _ = voc_tpfp_dense(d.to(dev), torch.tensor([n], dtype=torch.int32, device=dev), g.to(dev),
                          torch.tensor([k], dtype=torch.int32, device=dev), iou_thr, 1)
    t = t[0].cpu().numpy()
    tp[order] = t.astype(np.float32)
    fp[order] = 1.0 - tp[order]
    return tp, fp


def average_precision(recalls: np.ndarray, precisions: np.ndarray) -> np.float32:
    """eval_voc.py:108-150, mode 'area', one scale."""
    mrec = np.hstack((np.zeros(1, recalls.dtype), recalls, np.ones(1, recalls.dtype)))
    mpre = np.hstack((np.zeros(1, precisions.dtype), precisions, np.zeros(1, precisions.dtype)))
    for i in range(mpre.shape[0] - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    ind = np.where(mrec[1:] != mrec[:-1])[0]
    return np.float32(np.sum((mrec[ind + 1] - mrec[ind]) * mpre[ind + 1]))


def voc_ap(dets: torch.Tensor, counts: torch.Tensor, tp: torch.Tensor, num_gts: torch.Tensor) -> Dict[str, object]:
    """VOCEvaluator's arithmetic (eval_voc.py:37-66) on the gathered dense results: per class, detections of all images
    sorted by score, cumulative TP / FP, recall / precision, AP; mean over the classes that have ground truth.
    One device->host copy of four small arrays."""
    B, max_det, _ = dets.shape
    valid = (torch.arange(max_det, device=dets.device)[None, :] < counts[:, None]).cpu().numpy().reshape(-1)
    score = dets[..., 4].cpu().numpy().reshape(-1)[valid]
    cls = dets[..., 5].cpu().numpy().reshape(-1)[valid].astype(np.int64)
    tpf = tp.cpu().numpy().reshape(-1)[valid].astype(np.float32)
    ngt = num_gts.cpu().numpy()
    results = []
    eps = np.finfo(np.float32).eps
    for c in range(ngt.shape[0]):
        m = cls == c
        sort_inds = np.argsort(-score[m])                       # :38
        t = np.cumsum(tpf[m][sort_inds], axis=0)                # :40-44
        f = np.cumsum((1.0 - tpf[m])[sort_inds].astype(np.float32), axis=0)
        num = np.zeros(1, dtype=int)
        num[0] = int(ngt[c])
        recalls = t / np.maximum(num, eps)                      # :46
        precisions = t / np.maximum((t + f), eps)               # :47
        ap = average_precision(recalls, precisions) if t.shape[0] else np.float32(0.0)
        results.append({"num_gts": int(ngt[c]), "num_dets": int(m.sum()), "recall": recalls, "precision": precisions, "ap": ap})
    aps = [r["ap"] for r in results if r["num_gts"] > 0]
    return {"mean_ap": float(np.array(aps).mean()) if aps else 0.0, "results": results}


def gather_voc_stats(tp: torch.Tensor, num_gts: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Across ranks: all-gather of the TP flags (image order == the detection all-gather's) and all-reduce of the
    per-class ground-truth counts.  Equal shards per rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return tp, num_gts
    t8 = tp.view(torch.uint8).contiguous()
    out = t8.new_empty((world * t8.shape[0],) + tuple(t8.shape[1:]))
    dist.all_gather_into_tensor(out, t8, group=group)
    n = num_gts.clone()
    dist.all_reduce(n, group=group)
    return out.view(torch.bool), n
