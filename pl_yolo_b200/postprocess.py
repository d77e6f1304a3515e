"""Drop-in for models/evaluators/postprocess.py:7-48 (`postprocess`) and :51-92 (`demo_postprocess`).

Same signature, same return structure (a Python list with one `[n_i <= 300, 6]` tensor per image,
`None` where nothing survives), but the per-image Python loop, the boolean-mask gathers and
torchvision's batched_nms are replaced by two kernel launches for the whole batch and ONE
device->host copy (the per-image counts).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

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

    def __torch_function__(self, func, types, args=(), kwargs=None):  # pragma: no cover - convenience only
        kwargs = kwargs or {}
        args = [a.tensor if isinstance(a, LazyPredictions) else a for a in args]
        return func(*args, **kwargs)


def postprocess_dense(predictions: Union[torch.Tensor, LazyPredictions], conf_thre: float = 0.7, nms_thre: float = 0.45,
                      class_agnostic: bool = False, flavor: int = _lib.FLAVOR_CUDA):
    """Batch-level result without any host synchronisation:
    (dets [B,300,6] zero padded, counts [B] int32, keep_idx [B,300] anchor ids)."""
    if isinstance(predictions, LazyPredictions) and predictions._tensor is None:
        return ops.decode_postprocess_raw(predictions.inputs, predictions.strides, conf_thre, nms_thre, class_agnostic,
                                          MAX_NMS, MAX_DET, flavor)
    p = predictions.tensor if isinstance(predictions, LazyPredictions) else predictions
    return ops.postprocess_raw(p, conf_thre, nms_thre, class_agnostic, MAX_NMS, MAX_DET, flavor)


def postprocess(predictions, conf_thre=0.7, nms_thre=0.45, class_agnostic=False) -> List[Optional[torch.Tensor]]:
    """postprocess.py:7 — rows (x1, y1, x2, y2, confidence, class_pred), score-descending."""
    B = predictions.shape[0]
    output: List[Optional[torch.Tensor]] = [None for _ in range(B)]
    if B == 0 or predictions.shape[1] == 0:
        return output
    dets, counts, _ = postprocess_dense(predictions, conf_thre, nms_thre, class_agnostic)
    n = counts.tolist()  # the only device->host synchronisation of the batch
    for i in range(B):
        if n[i]:
            output[i] = dets[i, : n[i]]
    return output


demo_postprocess = postprocess  # postprocess.py:51-92 is a verbatim copy of :7-48
