"""Op-for-op replay of the reference path with the torch / torchvision LIBRARY ops it calls.

TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).  /root/reference does not exist on
the GPU box, so this is how "the reference on CUDA tensors on this B200" is obtained there: the
same ATen / torchvision kernels, launched in the same order, on whatever device the inputs live on.
It is validated in the build container against the REAL reference on CPU
(tests/test_replay_vs_golden.py: bit-for-bit, since on one device the ops are the same binaries).
It is also the reference arm of bench.py (`--impl reference`) and its cpu_baseline: the real cost of
the reference's own implementation on the host cores.

Citations: yolox_loss.py = models/losses/yolox/yolox_loss.py, postprocess.py =
models/evaluators/postprocess.py, iou_loss.py = models/layers/losses/iou_loss.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
import torchvision


def anchor_grid(hw: Sequence[Sequence[int]], strides: Sequence[int], like: torch.Tensor):
    """x_shifts, y_shifts, expanded_strides [1,A] (yolox_loss.py:198-208, :225-227)."""
    xs, ys, ss = [], [], []
    for (h, w), s in zip(hw, strides):
        yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        xs.append(xv.reshape(1, -1).type_as(like))
        ys.append(yv.reshape(1, -1).type_as(like))
        ss.append(torch.zeros(1, h * w).fill_(s).type_as(like))
    return torch.cat(xs, 1), torch.cat(ys, 1), torch.cat(ss, 1)


def decode(heads: Sequence[torch.Tensor], strides: Sequence[int], inference: bool):
    """yolox_loss.py:175-228 (+ :25-36 when inference).  -> preds [B,A,5+C], ori [B,A,4]"""
    outs, oris = [], []
    for x, s in zip(heads, strides):
        B, ch, h, w = x.shape
        yv, xv = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        grid = torch.stack((xv, yv), 2).view(1, h * w, 2).type_as(x)
        p = x.flatten(2).permute(0, 2, 1).contiguous()  # [B, hw, ch]  (:210-213)
        oris.append(p[..., :4].clone())                 # :214
        xy = (p[..., :2] + grid) * s                    # :217
        wh = torch.exp(p[..., 2:4]) * s                 # :219
        outs.append(torch.cat([xy, wh, p[..., 4:]], -1))
    preds = torch.cat(outs, 1)
    ori = torch.cat(oris, 1)
    if inference:
        obj = preds[..., 4:5].sigmoid()                 # :26
        cls = preds[..., 5:].sigmoid()                  # :27
        half_w, half_h = preds[..., 2] / 2, preds[..., 3] / 2
        x1 = preds[..., 0] - half_w                     # :31-34
        y1 = preds[..., 1] - half_h
        x2 = preds[..., 0] + half_w
        y2 = preds[..., 1] + half_h
        preds = torch.cat([torch.stack([x1, y1, x2, y2], -1), obj, cls], -1)
    return preds, ori


def postprocess(predictions: torch.Tensor, conf_thre: float = 0.7, nms_thre: float = 0.45,
                class_agnostic: bool = False, max_det: int = 300, max_nms: int = 10000) -> List[Optional[torch.Tensor]]:
    """postprocess.py:7-48."""
    out: List[Optional[torch.Tensor]] = [None] * predictions.shape[0]
    for i in range(predictions.shape[0]):
        ip = predictions[i]
        if not ip.shape[0]:
            continue
        cc, cp = torch.max(ip[:, 5:], 1, keepdim=True)          # :18
        conf = ip[:, 4] * cc.squeeze(1)                         # :19
        det = torch.cat((ip[:, :4], conf.unsqueeze(-1), cp.float()), 1)[conf >= conf_thre]  # :20-23
        det = det[:max_nms]                                     # :24-25
        if not det.size(0):
            continue
        if class_agnostic:
            keep = torchvision.ops.nms(det[:, :4], det[:, 4], nms_thre)                       # :30-34
        else:
            keep = torchvision.ops.batched_nms(det[:, :4], det[:, 4], det[:, 5], nms_thre)    # :36-41
        out[i] = det[keep][:max_det]                            # :43-46
    return out


def pairwise_iou_cxcywh(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """iou_loss.py:400-414 with xyxy=False."""
    tl = torch.max(a[:, None, :2] - a[:, None, 2:] / 2, b[:, :2] - b[:, 2:] / 2)
    br = torch.min(a[:, None, :2] + a[:, None, 2:] / 2, b[:, :2] + b[:, 2:] / 2)
    area_a = torch.prod(a[:, 2:], 1)
    area_b = torch.prod(b[:, 2:], 1)
    en = (tl < br).type(tl.type()).prod(dim=2)
    area_i = torch.prod(br - tl, 2) * en
    return area_i / (area_a[:, None] + area_b - area_i)


def geometry_prior(gt: torch.Tensor, es: torch.Tensor, xs: torch.Tensor, ys: torch.Tensor):
    """get_in_boxes_info (yolox_loss.py:231-315) with broadcasting instead of repeat (same elementwise ops)."""
    s = es[0]
    xc = (xs[0] * s + 0.5 * s)[None, :]
    yc = (ys[0] * s + 0.5 * s)[None, :]
    gl = (gt[:, 0] - 0.5 * gt[:, 2])[:, None]
    gr = (gt[:, 0] + 0.5 * gt[:, 2])[:, None]
    gtop = (gt[:, 1] - 0.5 * gt[:, 3])[:, None]
    gb = (gt[:, 1] + 0.5 * gt[:, 3])[:, None]
    d = torch.stack([xc - gl, yc - gtop, gr - xc, gb - yc], 2)
    in_box = d.min(dim=-1).values > 0.0
    r = 2.5 * s[None, :]
    cl = gt[:, 0:1] - r
    cr = gt[:, 0:1] + r
    ct = gt[:, 1:2] - r
    cb = gt[:, 1:2] + r
    dc = torch.stack([xc - cl, yc - ct, cr - xc, cb - yc], 2)
    in_ctr = dc.min(dim=-1).values > 0.0
    fg = (in_box.sum(0) > 0) | (in_ctr.sum(0) > 0)
    return fg, in_box[:, fg] & in_ctr[:, fg]


def dynamic_k(cost: torch.Tensor, ious: torch.Tensor, stable: bool):
    """dynamic_k_matching (yolox_loss.py:318-370) up to the matching matrix.  `stable=True` fixes the
    tie rule to lowest index (the north-star rule); False is the reference's plain sort()."""
    G, Nc = cost.shape
    M = torch.zeros_like(cost)
    k = min(10, Nc)
    srt, _ = ious.sort(descending=True)
    dks = torch.clamp(srt[:, :k].sum(1).int(), min=1)           # :338-340
    for g in range(G):
        _, pos = cost[g].sort(stable=stable)                    # :342
        kk = int(dks[g].item())
        if kk < pos.numel() - 1:                                # :343
            pos = pos[:kk]
        M[g][pos] = 1.0                                         # :348
    am = M.sum(0)
    if (am > 1).sum() > 0:                                      # :352-356
        _, amin = torch.min(cost[:, am > 1], dim=0)
        M[:, am > 1] *= 0.0
        M[amin, am > 1] = 1.0
    return M, dks


def simota(preds: torch.Tensor, labels: torch.Tensor, hw: Sequence[Sequence[int]], strides: Sequence[int],
           stable: bool = True):
    """The per-image block of YOLOXLoss.__call__ (yolox_loss.py:43-118) for a batch.
    -> dict of dense per-anchor results like oracle.simota."""
    B, A, ch = preds.shape
    C = ch - 5
    xs, ys, es = anchor_grid(hw, strides, preds)
    nlabel = (labels.sum(dim=2) > 0).sum(dim=1)                 # :43
    fg_out = torch.zeros(B, A, dtype=torch.bool, device=preds.device)
    mg_out = torch.full((B, A), -1, dtype=torch.int32, device=preds.device)
    mi_out = torch.zeros(B, A, dtype=torch.float32, device=preds.device)
    nfg = torch.zeros(B, dtype=torch.int32)
    dyn = torch.zeros(B, labels.shape[1], dtype=torch.int32)
    ncand = torch.zeros(B, dtype=torch.int32)
    for b in range(B):
        G = int(nlabel[b])
        if G == 0:
            continue
        gt = labels[b, :G, 1:5]
        gcls = labels[b, :G, 0]
        fg, both = geometry_prior(gt, es, xs, ys)
        box = preds[b, :, :4][fg]
        cls_ = preds[b, :, 5:][fg]
        obj_ = preds[b, :, 4:5][fg]
        Nc = box.shape[0]
        ncand[b] = Nc
        ious = pairwise_iou_cxcywh(gt, box)                     # :84
        iou_loss = -torch.log(ious + 1e-8)                      # :86
        onehot = F.one_hot(gcls.to(torch.int64), C).float().unsqueeze(1).repeat(1, Nc, 1)
        p = cls_.float().unsqueeze(0).repeat(G, 1, 1).sigmoid_() * obj_.unsqueeze(0).repeat(G, 1, 1).sigmoid_()
        cls_loss = F.binary_cross_entropy(p.sqrt_(), onehot, reduction="none").sum(-1)   # :99-101
        cost = cls_loss + 3.0 * iou_loss + 100000.0 * (~both)   # :104-108
        M, dks = dynamic_k(cost, ious, stable)
        sel = M.sum(0) > 0.0
        nfg[b] = int(sel.sum())
        dyn[b, :G] = dks.cpu()
        idx = torch.nonzero(fg)[:, 0][sel]
        fg_out[b, idx] = True
        mg_out[b, idx] = M[:, sel].argmax(0).to(torch.int32)    # :363
        mi_out[b, idx] = (M * ious).sum(0)[sel]                 # :367-369
    return {"fg_mask": fg_out, "matched_gt": mg_out, "matched_iou": mi_out, "num_fg": nfg,
            "num_gt": nlabel.to(torch.int32).cpu(), "dyn_k": dyn, "n_cand": ncand}


# ---- sibling heads: the NMS call sites of the YOLOv3 / YOLOv5 decoders and YOLOv7's matching block,
# replayed op for op on whatever device the inputs live on (multi_label == False, classes == None)
def yolov3_nms(predictions: torch.Tensor, conf_thre=0.7, nms_thre=0.45, max_nms=10000, max_det=300):
    """models/losses/yolov3/yolov3_decoder.py:63-116 on the decoded predictions [B,N,5+C] (cx,cy,w,h,obj,cls..)."""
    predictions = predictions.clone()
    box_corner = predictions.new(predictions.shape)
    box_corner[:, :, 0] = predictions[:, :, 0] - predictions[:, :, 2] / 2      # :64-68
    box_corner[:, :, 1] = predictions[:, :, 1] - predictions[:, :, 3] / 2
    box_corner[:, :, 2] = predictions[:, :, 0] + predictions[:, :, 2] / 2
    box_corner[:, :, 3] = predictions[:, :, 1] + predictions[:, :, 3] / 2
    predictions[:, :, :4] = box_corner[:, :, :4]
    output = [None for _ in range(len(predictions))]
    for b_idx, image_pred in enumerate(predictions):
        image_pred = image_pred[image_pred[..., 4] > conf_thre]                  # :74
        if not image_pred.size(0):
            continue
        image_pred[:, 5:] *= image_pred[:, 4:5]                                  # :79
        conf, j = image_pred[:, 5:].max(1, keepdim=True)                         # :86
        x = torch.cat((image_pred[:, :5], conf, j.float()), 1)[conf.view(-1) > conf_thre]
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 5].argsort(descending=True)[:max_nms]]                    # :98-100
        boxes, scores = x[:, :4], x[:, 5]
        nms = torchvision.ops.nms(boxes, scores, nms_thre)                       # :106 (class offset never applied)
        if nms.shape[0] > max_det:
            nms = nms[:max_det]
        output[b_idx] = x[nms]
    return output


def yolov5_nms(predictions: torch.Tensor, conf_thre=0.7, nms_thre=0.45, agnostic=False, max_nms=30000, max_det=300):
    """models/losses/yolov5/yolov5_decoder.py:23-87 on the decoded predictions [B,N,5+C]."""
    predictions = predictions.clone()
    obj_mask = predictions[..., 4] > conf_thre                                  # :23
    max_wh = 4096
    output = [torch.zeros((0, 7), device=predictions.device)] * predictions.shape[0]
    for img_idx, x in enumerate(predictions):
        x = x[obj_mask[img_idx]]
        if not x.shape[0]:
            continue
        box = x[:, :4].clone()                                                   # xywh2xyxy, models/utils/bbox.py:5-12
        box[:, 0] = x[:, 0] - x[:, 2] / 2
        box[:, 1] = x[:, 1] - x[:, 3] / 2
        box[:, 2] = x[:, 0] + x[:, 2] / 2
        box[:, 3] = x[:, 1] + x[:, 3] / 2
        x[:, :4] = box
        obj = x[:, 4]
        conf, j = x[:, 5:].max(1, keepdim=True)                                  # :57
        conf_mask = ((obj[:, None] * conf) >= conf_thre).squeeze(-1)             # :58
        x = torch.cat((box, obj[:, None], conf, j.float()), 1)[conf_mask]
        n = x.shape[0]
        if not n:
            continue
        elif n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]                    # :66-67
        c = x[:, 6] * (0 if agnostic else max_wh)                                # :70
        boxes, scores = x[:, :4] + c.unsqueeze(-1), x[:, 4]                      # :71
        i = torchvision.ops.nms(boxes, scores, nms_thre)                         # :72
        if i.shape[0] > max_det:
            i = i[:max_det]
        output[img_idx] = x[i]
    return output


def yolov7_matching(cost: torch.Tensor, pair_wise_iou: torch.Tensor):
    """models/losses/yolov7/yolov7_loss.py:233-262 -> (fg_mask_inboxes [N] bool, matched_gt_inds, dynamic_ks)."""
    top_k, _ = torch.topk(pair_wise_iou, min(10, pair_wise_iou.shape[1]), dim=1)   # :233
    dynamic_ks = torch.clamp(top_k.sum(1).int(), min=1)                            # :234
    matching_matrix = torch.zeros_like(cost)
    for gt_idx in range(cost.shape[0]):
        _, pos_idx = torch.topk(cost[gt_idx], k=dynamic_ks[gt_idx].item(), largest=False)   # :243-246
        matching_matrix[gt_idx][pos_idx] = 1.0
    anchor_matching_gt = matching_matrix.sum(0)
    if (anchor_matching_gt > 1).sum() > 0:                                         # :251-254
        _, cost_argmin = torch.min(cost[:, anchor_matching_gt > 1], dim=0)
        matching_matrix[:, anchor_matching_gt > 1] *= 0.0
        matching_matrix[cost_argmin, anchor_matching_gt > 1] = 1.0
    fg_mask_inboxes = matching_matrix.sum(0) > 0.0
    return fg_mask_inboxes, matching_matrix[:, fg_mask_inboxes].argmax(0), dynamic_ks
