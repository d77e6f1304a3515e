"""Drop-in for models/evaluators/postprocess.py:7-48 (`postprocess`) and :51-92 (`demo_postprocess`).

Same signature, same return structure (a Python list with one `[n_i <= 300, 6]` tensor per image,
`None` where nothing survives), but the per-image Python loop, the boolean-mask gathers and
torchvision's batched_nms are replaced by two kernel launches for the whole batch and ONE
device->host copy (the per-image counts).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, ops

MAX_DET = 300    # postprocess.py:8
MAX_NMS = 10000  # postprocess.py:9


class LazyPredictions:
    """What `YOLOXLoss(..., lazy_eval=True)` returns in eval mode: the head maps plus the decode
    recipe.  `postprocess` recognises it and takes the fused route (head maps read once, the
    [B,A,5+C] tensor never written); anything else that touches it (`.tensor`, indexing, `.shape`)
    materialises the reference's tensor with the decode kernel."""

    def __init__(self, inputs: Sequence[torch.Tensor], strides: Sequence[int]):
        self.inputs = list(inputs)
        self.strides = [int(s) for s in strides]
        self._tensor: Optional[torch.Tensor] = None

    @property
    def tensor(self) -> torch.Tensor:
        if self._tensor is None:
            self._tensor, _ = ops.decode_raw(self.inputs, self.strides, True)
        return self._tensor

    @property
    def shape(self):
        x = self.inputs[0]
        return torch.Size((x.shape[0], sum(t.shape[2] * t.shape[3] for t in self.inputs), x.shape[1]))

    def __getitem__(self, idx):
        return self.tensor[idx]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """torch functions applied to the lazy object see the materialised reference tensor."""
        def unwrap(a):
            if isinstance(a, LazyPredictions):
                return a.tensor
            if isinstance(a, (list, tuple)):
                return type(a)(unwrap(x) for x in a)
            return a
        return func(*unwrap(tuple(args)), **{k: unwrap(v) for k, v in (kwargs or {}).items()})


def postprocess_dense(predictions: Union[torch.Tensor, LazyPredictions], conf_thre: float = 0.7, nms_thre: float = 0.45,
                      class_agnostic: bool = False, flavor: int = _lib.FLAVOR_CUDA):
    """Batch-level result without any host synchronisation:
    (dets [B,300,6] zero padded, counts [B] int32, keep_idx [B,300] anchor ids)."""
    if isinstance(predictions, LazyPredictions) and predictions._tensor is None:
        return ops.decode_postprocess_raw(predictions.inputs, predictions.strides, conf_thre, nms_thre, class_agnostic,
                                          MAX_NMS, MAX_DET, flavor)
    p = predictions.tensor if isinstance(predictions, LazyPredictions) else predictions
    return ops.postprocess_raw(p, conf_thre, nms_thre, class_agnostic, MAX_NMS, MAX_DET, flavor)


class DetList(list):
    """What `postprocess` returns: the reference's list (one `[n_i, 6]` tensor or None per image) that also
    remembers the dense batch result, so that `format_outputs` can rescale / convert the whole batch with one
    kernel and one device->host copy instead of one `.cpu()` per detection."""
    dense = None   # (dets [B,max_det,6], counts [B] int32) on the device
    counts_host = None


def postprocess(predictions, conf_thre=0.7, nms_thre=0.45, class_agnostic=False) -> List[Optional[torch.Tensor]]:
    """postprocess.py:7 — rows (x1, y1, x2, y2, confidence, class_pred), score-descending."""
    B = predictions.shape[0]
    output = DetList([None for _ in range(B)])
    if B == 0 or predictions.shape[1] == 0:
        return output
    dets, counts, _ = postprocess_dense(predictions, conf_thre, nms_thre, class_agnostic)
    n = counts.tolist()  # the only device->host synchronisation of the batch
    for i in range(B):
        if n[i]:
            output[i] = dets[i, : n[i]]
    output.dense, output.counts_host = (dets, counts), n
    return output


def format_outputs(outputs, ids, hws, val_size, class_ids, labels=None):
    """Drop-in for models/evaluators/postprocess.py:95-138.  Same arguments, same results — `json_list` (COCO
    dicts, prediction order) and `det_list[image][class]` (VOC arrays of x1,y1,x2,y2,score) — and, like the
    reference, the boxes of `outputs` end up divided by the image's scale in place.  The per-detection Python
    loop with a `.cpu()` per box is replaced by one kernel (`bboxes /= scale`, xyxy2xywh) and one copy."""
    n_img = len(outputs)
    json_list = []
    det_list = [[np.empty(shape=[0, 5]) for _ in range(len(class_ids))] for _ in range(n_img)]
    if n_img == 0:
        return json_list, det_list
    # like the reference's zip(outputs, hws[0], hws[1], ids), the shortest argument decides how many images are visited
    n_img = min(n_img, len(hws[0]), len(hws[1]), len(ids))
    dense = getattr(outputs, "dense", None)
    if dense is not None:
        # the cached batch result is only valid while every entry still is the view `postprocess` returned
        d0, cnt0 = dense[0], outputs.counts_host
        same = len(cnt0) == len(outputs) and all(
            (o is None and cnt0[i] == 0) or (o is not None and cnt0[i] == o.shape[0] and o.data_ptr() == d0[i].data_ptr())
            for i, o in enumerate(outputs))
        if not same:
            dense = None
    outputs = list(outputs)[:n_img] if dense is None else outputs
    if dense is None:  # a plain list (e.g. produced elsewhere, or edited since): pad it to the dense layout first
        first = next((o for o in outputs if o is not None), None)
        if first is None:
            return json_list, det_list
        max_det = max(o.shape[0] for o in outputs if o is not None)
        dets = first.new_zeros((n_img, max_det, 6))
        cnt = [0 if o is None else int(o.shape[0]) for o in outputs]
        for i, o in enumerate(outputs):
            if o is not None:
                dets[i, : cnt[i]] = o
        counts = torch.tensor(cnt, dtype=torch.int32, device=dets.device)
    else:
        (dets, counts), cnt = dense, outputs.counts_host
        if n_img < dets.shape[0]:
            dets, counts, cnt = dets[:n_img], counts[:n_img].contiguous(), cnt[:n_img]
    # scale = min(val_size[0] / float(img_w), val_size[1] / float(img_h)) in Python doubles (postprocess.py:111)
    scales = [min(val_size[0] / float(w), val_size[1] / float(h)) for h, w in zip(hws[0], hws[1])][:n_img]
    # `tensor /= python_float` on CUDA multiplies by the reciprocal taken in double and rounded to fp32
    sc = torch.tensor([1.0 / v for v in scales], dtype=torch.float64).to(torch.float32).to(dets.device)
    rows = ops.format_dets_raw(dets, counts, sc)
    # the reference's in-place `bboxes /= scale` is visible to the caller through `outputs`
    for i, o in zip(range(n_img), outputs):
        if o is not None:
            o[:, 0:4] = rows[i, : cnt[i], 0:4]
    host = rows.cpu().numpy()  # the one device->host copy
    for i, img_id in zip(range(n_img), ids):
        if outputs[i] is None:
            continue
        r = host[i, : cnt[i]]
        cls = r[:, 7].astype(np.int64)
        xywh = r[:, [0, 1, 4, 5]]
        for j in range(cnt[i]):
            json_list.append({"image_id": int(img_id), "category_id": class_ids[int(cls[j])],
                              "bbox": xywh[j].tolist(), "score": r[j, 6].item(), "segmentation": []})
        for c in range(len(class_ids)):
            det_list[i][c] = r[cls == c][:, [0, 1, 2, 3, 6]]
    return json_list, det_list


demo_postprocess = postprocess  # postprocess.py:51-92 is a verbatim copy of :7-48
