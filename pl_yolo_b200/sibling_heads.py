"""NMS call sites and the SimOTA matching step of the sibling heads, on the same kernels.

`yolov3_nms` / `yolov5_nms` replace the per-image loops of models/losses/yolov3/yolov3_decoder.py:72-116 and
models/losses/yolov5/yolov5_decoder.py:30-87 (multi_label == False: the decoders' default): they take the decoded
`predictions [B, N, 5+C]` tensor both decoders build and return the same list.  `yolov7_matching` replaces the dynamic-k
/ conflict block of YOLOv7's build_targets (models/losses/yolov7/yolov7_loss.py:236-262) on the cost / IoU matrices it
computed.  INTEGRATION.md shows the edits in the reference files.
"""
from __future__ import annotations

from typing import List, Optional

import torch

from . import _lib, ops


def _check_counts(counts):
    n = counts.tolist()  # the one device->host synchronisation of the batch
    if any(v < 0 for v in n):
        raise _lib.PlyoloError("an image has more than 16384 NMS candidates: beyond what plyolo_postprocess_yolo_f32 sorts")
    return n


def yolov3_nms(predictions: torch.Tensor, conf_thre: float = 0.7, nms_thre: float = 0.45, max_nms: int = 10000,
               max_det: int = 300, classes=None, multi_label: bool = False) -> List[Optional[torch.Tensor]]:
    """yolov3_decoder.py:72-116 -> list of [n_i, 7] rows (x1,y1,x2,y2, obj, conf, class) or None."""
    if multi_label or classes is not None:
        raise NotImplementedError("multi_label / classes filtering are not on the hot path (the reference runs multi_label=False)")
    dets, counts, _ = ops.postprocess_yolo_raw(predictions, conf_thre, nms_thre, _lib.NMS_YOLOV3, True, max_nms, max_det)
    n = _check_counts(counts)
    return [dets[i, : n[i]] if n[i] else None for i in range(len(n))]


def yolov5_nms(predictions: torch.Tensor, conf_thre: float = 0.7, nms_thre: float = 0.45, agnostic: bool = False,
               max_nms: int = 30000, max_det: int = 300, multi_label: bool = False) -> List[torch.Tensor]:
    """yolov5_decoder.py:30-87 -> list of [n_i, 7] rows (x1,y1,x2,y2, obj, best class score, class); empty images give
    the reference's zeros((0, 7))."""
    if multi_label:
        raise NotImplementedError("multi_label is not on the hot path (the reference default is False)")
    dets, counts, _ = ops.postprocess_yolo_raw(predictions, conf_thre, nms_thre, _lib.NMS_YOLOV5, agnostic, max_nms, max_det)
    n = _check_counts(counts)
    return [dets[i, : n[i]] for i in range(len(n))]


def yolov7_matching(cost: torch.Tensor, pair_wise_iou: torch.Tensor):
    """yolov7_loss.py:236-262: dynamic k from the top-10 IoUs, torch.topk(cost, k, largest=False) per GT, conflicts to the
    argmin of the cost column -> (fg_mask_inboxes [N] bool, matched_gt_inds [num_fg] int64)."""
    sel, mg, _, _, _ = ops.dynamic_k_matching_raw(cost, pair_wise_iou, exact_k=True)
    return sel, mg[sel].to(torch.int64)
