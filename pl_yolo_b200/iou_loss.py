"""Drop-in for `bboxes_iou` (models/layers/losses/iou_loss.py:391-414) plus the GIoU loss the YOLOX
loss tail needs (IOUloss, :7-50; plain torch ops with autograd — the loss tail is a "next" row)."""
from __future__ import annotations

import torch

from . import ops


def bboxes_iou(bboxes_a, bboxes_b, xyxy=True):
    if bboxes_a.shape[1] != 4 or bboxes_b.shape[1] != 4:
        raise IndexError  # iou_loss.py:392-393
    return ops.bboxes_iou_raw(bboxes_a, bboxes_b, bool(xyxy))


def giou_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """IOUloss(reduction="none", loss_type="giou") on (cx,cy,w,h) rows — iou_loss.py:13-43."""
    pred = pred.view(-1, 4)
    target = target.view(-1, 4)
    p_lo, p_hi = pred[:, :2] - pred[:, 2:] / 2, pred[:, :2] + pred[:, 2:] / 2
    t_lo, t_hi = target[:, :2] - target[:, 2:] / 2, target[:, :2] + target[:, 2:] / 2
    tl, br = torch.max(p_lo, t_lo), torch.min(p_hi, t_hi)
    area_p = torch.prod(pred[:, 2:], 1)
    area_g = torch.prod(target[:, 2:], 1)
    en = (tl < br).to(tl).prod(dim=1)
    area_i = torch.prod(br - tl, 1) * en
    iou = area_i / (area_p + area_g - area_i + 1e-16)
    c_tl, c_br = torch.min(p_lo, t_lo), torch.max(p_hi, t_hi)
    area_c = torch.prod(c_br - c_tl, 1)
    giou = iou - (area_c - area_i) / area_c.clamp(1e-16)
    return 1 - giou.clamp(min=-1.0, max=1.0)
