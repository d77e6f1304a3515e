"""Drop-in for models/losses/yolox/yolox_loss.py (`YOLOXLoss`).

Same constructor, same `__call__(inputs, labels)` / `decode(inputs)` signatures and return values:
  eval  -> Tensor [B, A, 5+C] (x1,y1,x2,y2, sigmoid(obj), sigmoid(cls))            (:25-36)
  train -> dict {loss, loss_iou, loss_obj, loss_cls, loss_l1, proportion}         (:165-173)
What changed underneath: decode is one kernel (with a hand-written backward so the training graph
still reaches the head), and the per-image / per-GT Python loops of the SimOTA assignment
(:54-118, ~14k launches and ~4k host syncs per batch of 32) are three kernel launches for the whole
batch.  The loss tail (:121-173) runs fused as well (N2): one forward kernel produces the three loss sums, one
backward kernel writes d(loss)/d(head maps) directly (decode backward folded in) — the [B,A,5+C] gradient tensor and
the reference's gathered / concatenated target tensors never exist; `use_l1=True` adds one small kernel each way
(:128-133, :158).  `fused_loss=False` keeps the batched torch code for the tail (same values; the torch path is what
the fused kernels are also tested against).
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .iou_loss import giou_loss
from .postprocess import LazyPredictions


class _DecodeTrain(torch.autograd.Function):
    """preds/ori = decode(head maps) with d(xy) = s*g, d(wh) = out_wh*g, d(obj,cls) = g."""

    @staticmethod
    def forward(ctx, strides, *inputs):
        preds, ori = ops.decode_raw(list(inputs), list(strides), False)
        ctx.strides = list(strides)
        ctx.shapes = [tuple(x.shape) for x in inputs]
        ctx.save_for_backward(preds)
        return preds, ori

    @staticmethod
    def backward(ctx, g_preds, g_ori):
        (preds,) = ctx.saved_tensors
        grads = []
        off = 0
        for (B, ch, h, w), s in zip(ctx.shapes, ctx.strides):
            n = h * w
            g = torch.zeros((B, n, ch), dtype=preds.dtype, device=preds.device) if g_preds is None else g_preds[:, off:off + n].clone()
            g[..., 0:2] *= s
            g[..., 2:4] *= preds[:, off:off + n, 2:4]
            if g_ori is not None:
                g[..., 0:4] += g_ori[:, off:off + n]
            grads.append(g.permute(0, 2, 1).reshape(B, ch, h, w))
            off += n
        return (None, *grads)


class _FusedTrainLoss(torch.autograd.Function):
    """(sum GIoU, sum obj BCE, sum cls BCE[, sum L1]) [3 or 4], num_fg [B], num_gt [B] = decode -> SimOTA -> loss tail of the
    head maps; backward: one kernel straight into the head maps (+ the L1 signs) (yolox_loss.py:22, :43-158)."""

    @staticmethod
    def forward(ctx, strides, use_l1, labels, *inputs):
        strides = list(strides)
        preds, ori = ops.decode_raw(list(inputs), strides, False)
        hw: List[int] = []
        for x in inputs:
            hw += [int(x.shape[2]), int(x.shape[3])]
        labels = labels.to(preds.dtype).contiguous()
        if labels.shape[1] == 0:  # no label rows at all: the reference returns the objectness-only loss (:57-62)
            labels = labels.new_zeros((labels.shape[0], 1, 5))
        fg, mg, miou, nfg, ngt = ops.simota_assign_raw(preds, labels, hw, strides)
        sums = ops.yolox_loss_sums_raw(preds, labels, fg, mg, miou)
        if use_l1:                                                 # :128-133, :158
            sums = torch.cat([sums, ops.yolox_l1_sum_raw(ori, labels, fg, mg, hw, strides)])
        ctx.strides, ctx.hw, ctx.use_l1 = strides, hw, bool(use_l1)
        ctx.save_for_backward(preds, labels, fg, mg, miou, ori if use_l1 else preds.new_empty(0))
        ctx.mark_non_differentiable(nfg, ngt)
        return sums, nfg, ngt

    @staticmethod
    def backward(ctx, g_sums, _g_nfg, _g_ngt):
        preds, labels, fg, mg, miou, ori = ctx.saved_tensors
        g_sums = g_sums.contiguous()
        grads = ops.yolox_loss_backward_raw(preds, labels, fg, mg, miou, g_sums[:3].contiguous(), ctx.hw, ctx.strides)
        if ctx.use_l1:
            ops.yolox_l1_backward_raw(ori, labels, fg, mg, g_sums[3:4].contiguous(), grads, ctx.hw, ctx.strides)
        return (None, None, None, *grads)


class YOLOXLoss(nn.Module):
    def __init__(self, num_classes, strides, use_l1=False, lazy_eval=False, fused_loss=True):
        super().__init__()
        self.num_classes = num_classes
        self.strides = strides
        self.n_anchors = 1
        self.use_l1 = use_l1
        self.lazy_eval = lazy_eval  # eval mode returns LazyPredictions (fused decode+NMS route)
        self.fused_loss = fused_loss  # training: loss tail + backward as kernels (N2), use_l1 included
        self.bcewithlog_loss = nn.BCEWithLogitsLoss(reduction="none")
        self.l1_loss = nn.L1Loss(reduction="none")
        self._grid_cache = {}

    # like the reference, __call__ is overridden directly (module hooks never fire there either)
    def __call__(self, inputs, labels):
        if not self.training:
            if self.lazy_eval:
                return LazyPredictions(inputs, self.strides)
            with torch.no_grad():
                preds, _ = ops.decode_raw(list(inputs), list(self.strides), True)
            return preds
        return self._train(inputs, labels)

    def _grids(self, inputs: Sequence[torch.Tensor]):
        key = (tuple((x.shape[2], x.shape[3]) for x in inputs), inputs[0].device, inputs[0].dtype)
        g = self._grid_cache.get(key)
        if g is None:
            xs, ys, es = [], [], []
            for x, s in zip(inputs, self.strides):
                h, w = x.shape[2], x.shape[3]
                yv, xv = torch.meshgrid(torch.arange(h, device=x.device), torch.arange(w, device=x.device), indexing="ij")
                xs.append(xv.reshape(1, -1).to(x.dtype))
                ys.append(yv.reshape(1, -1).to(x.dtype))
                es.append(torch.full((1, h * w), float(s), dtype=x.dtype, device=x.device))
            g = (torch.cat(xs, 1), torch.cat(ys, 1), torch.cat(es, 1))
            self._grid_cache[key] = g
        return g

    def decode(self, inputs):
        """yolox_loss.py:175 -> (preds [B,A,5+C], ori_boxes [B,A,4], x_shifts, y_shifts, expanded_strides [1,A])."""
        inputs = list(inputs)
        if torch.is_grad_enabled() and any(x.requires_grad for x in inputs):
            preds, ori = _DecodeTrain.apply(tuple(self.strides), *inputs)
        else:
            preds, ori = ops.decode_raw(inputs, list(self.strides), False)
        xs, ys, es = self._grids(inputs)
        return preds, ori, xs, ys, es

    def assign(self, preds: torch.Tensor, labels: torch.Tensor, inputs: Sequence[torch.Tensor]):
        """SimOTA for the batch (yolox_loss.py:43-118): (fg_mask [B,A] bool, matched_gt [B,A] i32,
        matched_iou [B,A], num_fg [B], num_gt [B])."""
        hw: List[int] = []
        for x in inputs:
            hw += [int(x.shape[2]), int(x.shape[3])]
        with torch.no_grad():
            return ops.simota_assign_raw(preds.detach(), labels.to(preds.dtype), hw, list(self.strides))

    def _train_fused(self, inputs, labels):
        """yolox_loss.py:20-173: the sums from the kernels, the scalar arithmetic of :148-171 here."""
        inputs = list(inputs)
        sums, nfg, ngt = _FusedTrainLoss.apply(tuple(self.strides), bool(self.use_l1), labels, *inputs)
        counts = torch.stack([nfg.sum(), ngt.sum()]).tolist()     # one host sync (the reference has thousands)
        num_fgs = max(int(counts[0]), 1)                           # :148
        num_gts = int(counts[1])
        loss_iou = sums[0] / num_fgs                               # :150
        loss_obj = sums[1] / num_fgs                               # :152
        loss_cls = sums[2] / num_fgs                               # :154
        loss_l1 = sums[3] / num_fgs if self.use_l1 else 0.0      # :158-160
        loss = 5.0 * loss_iou + loss_obj + loss_cls + loss_l1      # :162-163
        return {"loss": loss, "loss_iou": loss_iou, "loss_obj": loss_obj, "loss_cls": loss_cls, "loss_l1": loss_l1,
                "proportion": num_fgs / max(num_gts, 1)}

    def _train(self, inputs, labels):
        if self.fused_loss:
            return self._train_fused(inputs, labels)
        preds, oriboxes, x_shifts, y_shifts, expanded_strides = self.decode(inputs)
        B, A, _ = preds.shape
        C = self.num_classes
        bbox_preds = preds[:, :, :4]
        obj_preds = preds[:, :, 4].unsqueeze(-1)
        cls_preds = preds[:, :, 5:]
        labels = labels.to(preds.dtype)
        fg, mg, miou, nfg, ngt = self.assign(preds, labels, inputs)

        # ---- loss tail, yolox_loss.py:121-173, batched ----
        fg_flat = fg.view(-1)
        b_idx, a_idx = torch.nonzero(fg, as_tuple=True)           # ascending (image, anchor) == cat order (:142-145)
        g_idx = mg[b_idx, a_idx].long()
        matched = labels[b_idx, g_idx]                             # [num_fg, 5]
        cls_targets = F.one_hot(matched[:, 0].to(torch.int64), C) * miou[b_idx, a_idx].unsqueeze(-1)   # :123-125
        reg_targets = matched[:, 1:5]                              # :127
        obj_targets = fg_flat.unsqueeze(-1).to(preds.dtype)        # :126, :136
        num_fgs = max(int(nfg.sum()), 1)                           # :148
        num_gts = int(ngt.sum())

        loss_iou = giou_loss(bbox_preds.reshape(-1, 4)[fg_flat], reg_targets).sum() / num_fgs               # :150
        loss_obj = self.bcewithlog_loss(obj_preds.reshape(-1, 1), obj_targets).sum() / num_fgs               # :152
        loss_cls = self.bcewithlog_loss(cls_preds.reshape(-1, C)[fg_flat], cls_targets).sum() / num_fgs      # :154
        if self.use_l1:                                                                                      # :128-133, :158
            s = expanded_strides[0][a_idx]
            l1_t = torch.stack([matched[:, 1] / s - x_shifts[0][a_idx], matched[:, 2] / s - y_shifts[0][a_idx],
                                torch.log(matched[:, 3] / s + 1e-8), torch.log(matched[:, 4] / s + 1e-8)], 1)
            loss_l1 = self.l1_loss(oriboxes.reshape(-1, 4)[fg_flat], l1_t).sum() / num_fgs
        else:
            loss_l1 = 0.0
        loss = 5.0 * loss_iou + loss_obj + loss_cls + loss_l1      # :162-163
        return {"loss": loss, "loss_iou": loss_iou, "loss_obj": loss_obj, "loss_cls": loss_cls, "loss_l1": loss_l1,
                "proportion": num_fgs / max(num_gts, 1)}


def get_in_boxes_info(gt_bboxes_per_image, expanded_strides, x_shifts, y_shifts, total_num_anchors, num_gt):
    """Drop-in for the module-level function of models/losses/yolox/yolox_loss.py:231-315 (same arguments):
    -> (is_in_boxes_or_center [A] bool, is_in_boxes_and_center [num_gt, Nc] bool)."""
    fg, in_boxes, in_centers = ops.in_boxes_info_raw(gt_bboxes_per_image[:num_gt].contiguous(), expanded_strides, x_shifts, y_shifts)
    if fg.numel() != total_num_anchors:
        raise ValueError("total_num_anchors does not match the anchor vectors")
    return fg, in_boxes[:, fg] & in_centers[:, fg]                                    # :312-314


def dynamic_k_matching(fg_mask, cost, pair_wise_ious, gt_classes, num_gt):
    """Drop-in for models/losses/yolox/yolox_loss.py:318-370 (same arguments and return values; `fg_mask` is updated
    in place like the reference does, :361): -> (fg_mask, num_fg, matched_gt_inds, gt_matched_classes,
    pred_ious_this_matching)."""
    sel, mg, mi, _, _ = ops.dynamic_k_matching_raw(cost[:num_gt], pair_wise_ious[:num_gt])
    num_fg = sel.sum().detach()                                                       # :358
    fg_mask[fg_mask.clone()] = sel                                                    # :361 (quirk Q5: in place)
    matched_gt_inds = mg[sel].to(torch.int64)                                         # :363
    gt_matched_classes = gt_classes[matched_gt_inds]                                  # :365
    pred_ious_this_matching = mi[sel]                                                 # :367-369
    return fg_mask, num_fg, matched_gt_inds, gt_matched_classes, pred_ious_this_matching
