"""CPU restatement (numpy) of the loss tail of YOLOXLoss and of its gradient with respect to the head outputs —
TEST INFRASTRUCTURE ONLY (imported by tests/, never by pl_yolo_b200/).

Follows /root/reference:
  models/losses/yolox/yolox_loss.py:121-127  targets: one_hot(class) * IoU, fg mask, matched GT box
  models/losses/yolox/yolox_loss.py:148-163  num_fgs = max(sum num_fg, 1); loss_iou / loss_obj / loss_cls = sum / num_fgs;
                                             loss = 5 loss_iou + loss_obj + loss_cls (+ loss_l1 = 0)
  models/layers/losses/iou_loss.py:7-50      IOUloss(loss_type="giou")
  torch.nn.BCEWithLogitsLoss                  (1 - t) x - log_sigmoid(x)
  models/losses/yolox/yolox_loss.py:217-219  decode: cx = (px + grid) s, w = exp(pw) s  (chain rule into the head maps)
Pinned against the real reference's loss dict and autograd gradients by tests/golden/lossgrad_*.npz
(oracle/gen_golden.py lossgrad) in tests/test_oracle_vs_golden.py.  float64 arithmetic: the pin is 1e-5 relative.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np


def _bce_logits(x, t):
    return (1.0 - t) * x - (np.minimum(x, 0.0) - np.log1p(np.exp(-np.abs(x))))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def giou_loss_and_grad(p: np.ndarray, g: np.ndarray):
    """IOUloss giou of (cx,cy,w,h) rows and d loss / d p (autograd conventions: ties split, clamp inclusive)."""
    cx, cy, w, h = p.T
    gx, gy, gw, gh = g.T
    px1, px2, py1, py2 = cx - w / 2, cx + w / 2, cy - h / 2, cy + h / 2
    tx1, tx2, ty1, ty2 = gx - gw / 2, gx + gw / 2, gy - gh / 2, gy + gh / 2
    tlx, tly, brx, bry = np.maximum(px1, tx1), np.maximum(py1, ty1), np.minimum(px2, tx2), np.minimum(py2, ty2)
    area_p, area_g = w * h, gw * gh
    en = (tlx < brx).astype(p.dtype) * (tly < bry).astype(p.dtype)
    wi, hi = brx - tlx, bry - tly
    area_i = wi * hi * en
    U = area_p + area_g - area_i + 1e-16
    iou = area_i / U
    cx1, cy1, cx2, cy2 = np.minimum(px1, tx1), np.minimum(py1, ty1), np.maximum(px2, tx2), np.maximum(py2, ty2)
    wc, hc = cx2 - cx1, cy2 - cy1
    area_c = wc * hc
    Cc = np.maximum(area_c, 1e-16)
    giou = iou - (area_c - area_i) / Cc
    loss = 1.0 - np.clip(giou, -1.0, 1.0)
    g_giou = np.where((giou >= -1.0) & (giou <= 1.0), -1.0, 0.0)
    g_area_i = g_giou / Cc + g_giou * (1.0 / U + area_i / (U * U))
    g_area_c = -g_giou / Cc + np.where(area_c >= 1e-16, g_giou * (area_c - area_i) / (Cc * Cc), 0.0)
    g_area_p = -g_giou * area_i / (U * U)
    g_wi, g_hi = g_area_i * hi * en, g_area_i * wi * en
    g_wc, g_hc = g_area_c * hc, g_area_c * wc
    sgt = lambda a, b: (a > b) + 0.5 * (a == b)
    slt = lambda a, b: (a < b) + 0.5 * (a == b)
    g_px1 = -g_wi * sgt(px1, tx1) - g_wc * slt(px1, tx1)
    g_px2 = g_wi * slt(px2, tx2) + g_wc * sgt(px2, tx2)
    g_py1 = -g_hi * sgt(py1, ty1) - g_hc * slt(py1, ty1)
    g_py2 = g_hi * slt(py2, ty2) + g_hc * sgt(py2, ty2)
    grad = np.stack([g_px1 + g_px2, g_py1 + g_py2, 0.5 * (g_px2 - g_px1) + g_area_p * h, 0.5 * (g_py2 - g_py1) + g_area_p * w], 1)
    return loss, grad


def loss_tail(preds: np.ndarray, labels: np.ndarray, assign: Dict[str, np.ndarray], hw: Sequence[Sequence[int]],
              strides: Sequence[int], ori: np.ndarray = None):
    """preds [B,A,5+C] training-mode decode output, assign = the SimOTA result (fg_mask, matched_gt, matched_iou, num_fg,
    num_gt).  -> (loss dict, [d loss / d head map l as [B,5+C,H,W]]).  ori [B,A,4] = the raw regression outputs: with it the
    use_l1 term (yolox_loss.py:128-133, :158, get_l1_type :373-378) is added."""
    preds = preds.astype(np.float64)
    labels = labels.astype(np.float64)
    B, A, ch = preds.shape
    C = ch - 5
    fg = assign["fg_mask"].astype(bool)
    num_fgs = max(int(assign["num_fg"].sum()), 1)
    num_gts = int(assign["num_gt"].sum())
    b_idx, a_idx = np.nonzero(fg)
    matched = labels[b_idx, assign["matched_gt"][b_idx, a_idx]]
    tcls = np.zeros((len(b_idx), C))
    tcls[np.arange(len(b_idx)), matched[:, 0].astype(np.int64)] = assign["matched_iou"][b_idx, a_idx]
    giou, ggrad = giou_loss_and_grad(preds[b_idx, a_idx, :4], matched[:, 1:5])
    obj_t = fg.astype(np.float64)
    s_iou = giou.sum()
    s_obj = _bce_logits(preds[..., 4], obj_t).sum()
    s_cls = _bce_logits(preds[b_idx, a_idx, 5:], tcls).sum()
    loss_iou, loss_obj, loss_cls = s_iou / num_fgs, s_obj / num_fgs, s_cls / num_fgs
    loss_l1, gl1 = 0.0, None
    if ori is not None:
        ori = ori.astype(np.float64)
        sv = np.concatenate([np.full(H * W, float(st)) for (H, W), st in zip(hw, strides)])            # expanded_strides
        xs = np.concatenate([np.tile(np.arange(W, dtype=np.float64), H) for (H, W) in hw])              # x_shifts
        ys = np.concatenate([np.repeat(np.arange(H, dtype=np.float64), W) for (H, W) in hw])            # y_shifts
        st = sv[a_idx]
        t = np.stack([matched[:, 1] / st - xs[a_idx], matched[:, 2] / st - ys[a_idx],
                      np.log(matched[:, 3] / st + 1e-8), np.log(matched[:, 4] / st + 1e-8)], 1)         # get_l1_type
        d = ori[b_idx, a_idx] - t
        loss_l1 = np.abs(d).sum() / num_fgs                                                               # :158
        gl1 = np.sign(d) / num_fgs
    losses = {"loss": 5.0 * loss_iou + loss_obj + loss_cls + loss_l1, "loss_iou": loss_iou, "loss_obj": loss_obj,
              "loss_cls": loss_cls, "loss_l1": loss_l1, "proportion": num_fgs / max(num_gts, 1)}
    # d loss / d preds
    gp = np.zeros_like(preds)
    gp[..., 4] = (_sigmoid(preds[..., 4]) - obj_t) / num_fgs
    gp[b_idx, a_idx, :4] = 5.0 * ggrad / num_fgs
    gp[b_idx, a_idx, 5:] = (_sigmoid(preds[b_idx, a_idx, 5:]) - tcls) / num_fgs
    # chain through the decode into the channel-planar head maps
    grads: List[np.ndarray] = []
    off = 0
    for (H, W), s in zip(hw, strides):
        n = H * W
        gl = gp[:, off:off + n].copy()
        gl[..., 0:2] *= s
        gl[..., 2:4] *= preds[:, off:off + n, 2:4]
        if gl1 is not None:  # the raw outputs receive the L1 gradient directly
            g4 = np.zeros((B, A, 4))
            g4[b_idx, a_idx] = gl1
            gl[..., 0:4] += g4[:, off:off + n]
        grads.append(gl.transpose(0, 2, 1).reshape(B, ch, H, W))
        off += n
    return losses, grads
