"""torch.library ops `plyolo::*` over the C ABI (include/plyolo.h).

PyTorch is plumbing here: device memory, the current stream and shape checks.  Every op runs
hand-written sm_100a kernels from libplyolo.so on the caller's current CUDA stream; there is no
eager / CPU fallback (CPU tensors raise).

Each op exists twice: `<name>_raw` is the plain Python function (lowest host overhead, used by the
shims) and `<name>` is the same function registered as `torch.ops.plyolo.<name>` with a fake
(meta) kernel for shape inference under tracing.
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch

from . import _lib

_WS: Dict[Tuple[int, int, str], torch.Tensor] = {}


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(kind: str, nbytes: int, device: torch.device) -> torch.Tensor:
    """Scratch is cached per (device, stream, kind): launches on one stream are ordered, so reuse is safe."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream_ptr(device), kind)
    ws = _WS.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        assert ws.data_ptr() % 256 == 0
        _WS[key] = ws
    return ws


def _check_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise _lib.PlyoloError("%s is on %s: plyolo ops are CUDA-only (no CPU fallback)" % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError("%s must be float32, got %s" % (name, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def _levels(inputs: Sequence[torch.Tensor], strides: Sequence[int]):
    if len(inputs) != len(strides) or not 1 <= len(inputs) <= _lib.MAX_LEVELS:
        raise ValueError("need one stride per level and 1..%d levels" % _lib.MAX_LEVELS)
    xs = [_check_cuda_f32(x, "inputs[%d]" % i) for i, x in enumerate(inputs)]
    B, ch = xs[0].shape[0], xs[0].shape[1]
    for x in xs:
        if x.dim() != 4 or x.shape[0] != B or x.shape[1] != ch:
            raise ValueError("every level must be [B, 5+C, H, W] with the same B and C")
    hs = [int(x.shape[2]) for x in xs]
    ws = [int(x.shape[3]) for x in xs]
    A = sum(h * w for h, w in zip(hs, ws))
    return xs, B, ch - 5, hs, ws, A


# ------------------------------------------------------------------------------------------ decode
def decode_raw(inputs: List[torch.Tensor], strides: List[int], inference: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    """YOLOXLoss.decode (+ eval branch when `inference`): -> (preds [B,A,5+C], ori_boxes [B,A,4])."""
    xs, B, C, hs, ws, A = _levels(inputs, strides)
    dev = xs[0].device
    preds = torch.empty((B, A, 5 + C), dtype=torch.float32, device=dev)
    ori = torch.empty((B, A, 4), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_decode_f32(_lib.ptr_array([x.data_ptr() for x in xs]), _lib.int_array(hs),
                                          _lib.int_array(ws), _lib.int_array(strides), len(xs), B, C,
                                          preds.data_ptr(), ori.data_ptr(), int(inference), _stream_ptr(dev))
    _lib.check(rc, "plyolo_decode_f32")
    return preds, ori


decode = torch.library.custom_op("plyolo::decode", decode_raw, mutates_args=())


@decode.register_fake
def _(inputs, strides, inference):
    B, ch = inputs[0].shape[0], inputs[0].shape[1]
    A = sum(x.shape[2] * x.shape[3] for x in inputs)
    return inputs[0].new_empty((B, A, ch)), inputs[0].new_empty((B, A, 4))


# ------------------------------------------------------------------------------------- postprocess
def _det_outputs(B: int, max_det: int, dev: torch.device, out):
    if out is not None:
        dets, counts, keep = out
        assert dets.shape == (B, max_det, 6) and dets.dtype == torch.float32 and dets.is_contiguous() and dets.device == dev
        assert counts.shape == (B,) and counts.dtype == torch.int32 and counts.device == dev
        assert keep.shape == (B, max_det) and keep.dtype == torch.int32 and keep.is_contiguous() and keep.device == dev
        return dets, counts, keep
    return (torch.empty((B, max_det, 6), dtype=torch.float32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
            torch.empty((B, max_det), dtype=torch.int32, device=dev))


def postprocess_raw(preds: torch.Tensor, conf_thre: float, nms_thre: float, class_agnostic: bool, max_nms: int,
                    max_det: int, flavor: int, out=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (dets [B,max_det,6] zero padded, counts [B] i32, keep_idx [B,max_det] i32 anchor ids).
    `out=(dets, counts, keep_idx)` writes into caller-owned buffers (CUDA-graph friendly)."""
    p = _check_cuda_f32(preds, "preds")
    if p.dim() != 3 or p.shape[2] < 6:
        raise ValueError("preds must be [B, A, 5+C]")
    B, A, ch = p.shape
    dev = p.device
    dets, counts, keep = _det_outputs(B, max_det, dev, out)
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L.plyolo_postprocess_workspace_bytes(B, A)
        ws = _workspace("post", nbytes, dev)
        rc = L.plyolo_postprocess_f32(p.data_ptr(), B, A, ch - 5, float(conf_thre), float(nms_thre),
                                      int(class_agnostic), int(max_nms), int(max_det), int(flavor), dets.data_ptr(),
                                      counts.data_ptr(), keep.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_postprocess_f32")
    return dets, counts, keep


def _postprocess_op(preds: torch.Tensor, conf_thre: float, nms_thre: float, class_agnostic: bool, max_nms: int,
                    max_det: int, flavor: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return postprocess_raw(preds, conf_thre, nms_thre, class_agnostic, max_nms, max_det, flavor)


postprocess = torch.library.custom_op("plyolo::postprocess", _postprocess_op, mutates_args=())


@postprocess.register_fake
def _(preds, conf_thre, nms_thre, class_agnostic, max_nms, max_det, flavor):
    B = preds.shape[0]
    return (preds.new_empty((B, max_det, 6)), preds.new_empty((B,), dtype=torch.int32),
            preds.new_empty((B, max_det), dtype=torch.int32))


def postprocess_yolo_raw(preds: torch.Tensor, conf_thre: float, nms_thre: float, variant: int, class_agnostic: bool, max_nms: int,
                         max_det: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """NMS call sites of the YOLOv3 / YOLOv5 decoders on decoded predictions [B,N,5+C] (cx,cy,w,h,obj,cls..):
    -> (dets [B,max_det,7] rows (x1,y1,x2,y2,obj,conf,class) zero padded, counts [B] i32 (-1: more candidates than the
    kernel sorts), keep_idx [B,max_det] i32)."""
    p = _check_cuda_f32(preds, "predictions")
    if p.dim() != 3 or p.shape[2] < 6:
        raise ValueError("predictions must be [B, N, 5+C]")
    B, N, ch = p.shape
    dev = p.device
    dets = torch.empty((B, max_det, 7), dtype=torch.float32, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    keep = torch.empty((B, max_det), dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _workspace("post", L.plyolo_postprocess_workspace_bytes(B, N), dev)
        rc = L.plyolo_postprocess_yolo_f32(p.data_ptr(), B, N, ch - 5, float(conf_thre), float(nms_thre), int(variant),
                                           int(class_agnostic), int(max_nms), int(max_det), dets.data_ptr(), counts.data_ptr(),
                                           keep.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_postprocess_yolo_f32")
    return dets, counts, keep


def decode_postprocess_raw(inputs: List[torch.Tensor], strides: List[int], conf_thre: float, nms_thre: float,
                           class_agnostic: bool, max_nms: int, max_det: int, flavor: int,
                           out=None, peers=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Fused decode + postprocess from the head maps (reads them once; preds never materialised).
    peers = (dets_ptrs, counts_ptrs): device addresses of this rank's block inside the OTHER ranks' gathered buffers
    (peer memory): the NMS kernels store every row there too (pl_yolo_b200.distributed.PeerDetections)."""
    xs, B, C, hs, ws_, A = _levels(inputs, strides)
    dev = xs[0].device
    dets, counts, keep = _det_outputs(B, max_det, dev, out)
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L.plyolo_postprocess_workspace_bytes(B, A)
        ws = _workspace("post", nbytes, dev)
        pd, pc = peers if peers is not None else ((), ())
        rc = L.plyolo_decode_postprocess_bcast_f32(_lib.ptr_array([x.data_ptr() for x in xs]), _lib.int_array(hs),
                                                   _lib.int_array(ws_), _lib.int_array(strides), len(xs), B, C,
                                                   float(conf_thre), float(nms_thre), int(class_agnostic), int(max_nms),
                                                   int(max_det), int(flavor), dets.data_ptr(), counts.data_ptr(),
                                                   keep.data_ptr(), len(pd), _lib.ptr_array(list(pd) or [0]),
                                                   _lib.ptr_array(list(pc) or [0]), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_decode_postprocess_bcast_f32")
    return dets, counts, keep


def _decode_postprocess_op(inputs: List[torch.Tensor], strides: List[int], conf_thre: float, nms_thre: float,
                           class_agnostic: bool, max_nms: int, max_det: int,
                           flavor: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    return decode_postprocess_raw(inputs, strides, conf_thre, nms_thre, class_agnostic, max_nms, max_det, flavor)


decode_postprocess = torch.library.custom_op("plyolo::decode_postprocess", _decode_postprocess_op, mutates_args=())


@decode_postprocess.register_fake
def _(inputs, strides, conf_thre, nms_thre, class_agnostic, max_nms, max_det, flavor):
    B = inputs[0].shape[0]
    x = inputs[0]
    return (x.new_empty((B, max_det, 6)), x.new_empty((B,), dtype=torch.int32),
            x.new_empty((B, max_det), dtype=torch.int32))


# ------------------------------------------------------------------------------------------ SimOTA
def simota_assign_raw(preds: torch.Tensor, labels: torch.Tensor, hw: List[int],
                  strides: List[int]) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """SimOTA for the whole batch.  preds [B,A,5+C] training-mode decode output, labels [B,Lmax,5];
    hw = [H0, W0, H1, W1, ...].  -> (fg_mask [B,A] bool, matched_gt [B,A] i32 (-1 bg),
    matched_iou [B,A] f32, num_fg [B] i32, num_gt [B] i32)."""
    p = _check_cuda_f32(preds, "preds")
    lab = _check_cuda_f32(labels, "labels")
    if p.dim() != 3 or lab.dim() != 3 or lab.shape[2] != 5 or lab.shape[0] != p.shape[0]:
        raise ValueError("preds must be [B,A,5+C] and labels [B,Lmax,5]")
    B, A, ch = p.shape
    Lmax = lab.shape[1]
    hs, ws_ = list(hw[0::2]), list(hw[1::2])
    dev = p.device
    fg = torch.empty((B, A), dtype=torch.uint8, device=dev)
    mg = torch.empty((B, A), dtype=torch.int32, device=dev)
    mi = torch.empty((B, A), dtype=torch.float32, device=dev)
    nfg = torch.empty((B,), dtype=torch.int32, device=dev)
    ngt = torch.empty((B,), dtype=torch.int32, device=dev)
    if Lmax == 0:
        fg.zero_(); mg.fill_(-1); mi.zero_(); nfg.zero_(); ngt.zero_()
        return fg.bool(), mg, mi, nfg, ngt
    L = _lib.lib()
    with torch.cuda.device(dev):
        nbytes = L.plyolo_simota_workspace_bytes(B, A, Lmax, len(hs))
        ws = _workspace("simota", nbytes, dev)
        rc = L.plyolo_simota_f32(p.data_ptr(), lab.data_ptr(), B, A, ch - 5, Lmax, _lib.int_array(hs),
                                 _lib.int_array(ws_), _lib.int_array(strides), len(hs), fg.data_ptr(), mg.data_ptr(),
                                 mi.data_ptr(), nfg.data_ptr(), ngt.data_ptr(), ws.data_ptr(), ws.numel(),
                                 _stream_ptr(dev))
    _lib.check(rc, "plyolo_simota_f32")
    return fg.view(torch.bool), mg, mi, nfg, ngt


simota_assign = torch.library.custom_op("plyolo::simota_assign", simota_assign_raw, mutates_args=())


@simota_assign.register_fake
def _(preds, labels, hw, strides):
    B, A = preds.shape[0], preds.shape[1]
    return (preds.new_empty((B, A), dtype=torch.bool), preds.new_empty((B, A), dtype=torch.int32),
            preds.new_empty((B, A)), preds.new_empty((B,), dtype=torch.int32), preds.new_empty((B,), dtype=torch.int32))


def in_boxes_info_raw(gt: torch.Tensor, expanded_strides: torch.Tensor, x_shifts: torch.Tensor,
                      y_shifts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """get_in_boxes_info (yolox_loss.py:231-315): gt [G,4], the three [A] (or [1,A]) anchor vectors
    -> (fg_mask [A] bool, in_boxes [G,A] bool, in_centers [G,A] bool)."""
    g = _check_cuda_f32(gt, "gt_bboxes_per_image")
    es = _check_cuda_f32(expanded_strides, "expanded_strides").reshape(-1)
    xs = _check_cuda_f32(x_shifts, "x_shifts").reshape(-1)
    ys = _check_cuda_f32(y_shifts, "y_shifts").reshape(-1)
    if g.dim() != 2 or g.shape[1] != 4 or not (es.numel() == xs.numel() == ys.numel()):
        raise ValueError("gt must be [G,4] and the anchor vectors must have one length")
    A, G = es.numel(), g.shape[0]
    dev = g.device
    fg = torch.empty((A,), dtype=torch.uint8, device=dev)
    ib = torch.empty((G, A), dtype=torch.uint8, device=dev)
    ic = torch.empty((G, A), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_in_boxes_info_f32(g.data_ptr(), es.data_ptr(), xs.data_ptr(), ys.data_ptr(), A, G, fg.data_ptr(),
                                                 ib.data_ptr(), ic.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_in_boxes_info_f32")
    return fg.view(torch.bool), ib.view(torch.bool), ic.view(torch.bool)


def dynamic_k_matching_raw(cost: torch.Tensor, ious: torch.Tensor, exact_k: bool = False):
    """dynamic_k_matching (yolox_loss.py:318-370) on a [G,Nc] cost / IoU matrix -> (selected [Nc] bool,
    matched_gt [Nc] i32 (-1 = not selected), matched_iou [Nc], dynamic_ks [G] i32, matching [G,Nc] bool).
    exact_k: the YOLOv7 rule (yolov7_loss.py:236-262, torch.topk: exactly k per GT, no `k >= Nc-1` quirk)."""
    c = _check_cuda_f32(cost, "cost")
    i = _check_cuda_f32(ious, "pair_wise_ious")
    if c.dim() != 2 or c.shape != i.shape:
        raise ValueError("cost and pair_wise_ious must be [G,Nc]")
    G, Nc = c.shape
    dev = c.device
    M = torch.empty((G, Nc), dtype=torch.uint8, device=dev)
    dk = torch.empty((G,), dtype=torch.int32, device=dev)
    sel = torch.zeros((Nc,), dtype=torch.uint8, device=dev)
    mg = torch.full((Nc,), -1, dtype=torch.int32, device=dev)
    mi = torch.zeros((Nc,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_dynamic_k_matching_f32(c.data_ptr(), i.data_ptr(), G, Nc, int(exact_k), M.data_ptr(), dk.data_ptr(),
                                                      sel.data_ptr(), mg.data_ptr(), mi.data_ptr(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_dynamic_k_matching_f32")
    return sel.view(torch.bool), mg, mi, dk, M.view(torch.bool)


# ------------------------------------------------------------------------------------------ loss tail (N2)
def _loss_inputs(preds, labels, fg_mask, matched_gt, matched_iou):
    p = _check_cuda_f32(preds, "preds")
    lab = _check_cuda_f32(labels, "labels")
    mi = _check_cuda_f32(matched_iou, "matched_iou")
    if p.dim() != 3 or lab.dim() != 3 or lab.shape[2] != 5 or lab.shape[0] != p.shape[0]:
        raise ValueError("preds must be [B,A,5+C] and labels [B,Lmax,5]")
    B, A, _ = p.shape
    fg = fg_mask.contiguous()
    fg = fg.view(torch.uint8) if fg.dtype == torch.bool else fg
    if fg.dtype != torch.uint8 or matched_gt.dtype != torch.int32:
        raise TypeError("fg_mask must be bool / uint8 and matched_gt int32 (the outputs of simota_assign)")
    mg = matched_gt.contiguous()
    for t, n in ((fg, "fg_mask"), (mg, "matched_gt"), (mi, "matched_iou")):
        if not t.is_cuda or tuple(t.shape) != (B, A):
            raise ValueError("%s must be a CUDA tensor of shape [B, A]" % n)
    return p, lab, fg, mg, mi


def yolox_loss_sums_raw(preds: torch.Tensor, labels: torch.Tensor, fg_mask: torch.Tensor, matched_gt: torch.Tensor,
                        matched_iou: torch.Tensor) -> torch.Tensor:
    """Loss tail of YOLOXLoss (yolox_loss.py:121-154, use_l1=False): -> [3] = (sum GIoU loss over the foreground
    anchors, sum objectness BCE over all anchors, sum class BCE over the foreground anchors)."""
    p, lab, fg, mg, mi = _loss_inputs(preds, labels, fg_mask, matched_gt, matched_iou)
    B, A, ch = p.shape
    dev = p.device
    sums = torch.empty((3,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _workspace("loss", L.plyolo_yolox_loss_workspace_bytes(B, A), dev)
        rc = L.plyolo_yolox_loss_f32(p.data_ptr(), lab.data_ptr(), fg.data_ptr(), mg.data_ptr(), mi.data_ptr(), B, A,
                                     ch - 5, lab.shape[1], sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
    _lib.check(rc, "plyolo_yolox_loss_f32")
    return sums


def yolox_loss_backward_raw(preds: torch.Tensor, labels: torch.Tensor, fg_mask: torch.Tensor, matched_gt: torch.Tensor,
                            matched_iou: torch.Tensor, grad_sums: torch.Tensor, hw: List[int],
                            strides: List[int]) -> List[torch.Tensor]:
    """d(sums . grad_sums) / d(head maps), chained through the training-mode decode: one [B,5+C,H,W] tensor per level."""
    p, lab, fg, mg, mi = _loss_inputs(preds, labels, fg_mask, matched_gt, matched_iou)
    g = _check_cuda_f32(grad_sums, "grad_sums")
    if g.numel() != 3:
        raise ValueError("grad_sums must have 3 elements")
    B, A, ch = p.shape
    hs, ws_ = list(hw[0::2]), list(hw[1::2])
    if sum(h * w for h, w in zip(hs, ws_)) != A:
        raise ValueError("the level shapes do not add up to A")
    dev = p.device
    grads = [torch.empty((B, ch, h, w), dtype=torch.float32, device=dev) for h, w in zip(hs, ws_)]
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_yolox_loss_backward_f32(
            p.data_ptr(), lab.data_ptr(), fg.data_ptr(), mg.data_ptr(), mi.data_ptr(), B, ch - 5, lab.shape[1],
            g.data_ptr(), _lib.ptr_array([t.data_ptr() for t in grads]), _lib.int_array(hs), _lib.int_array(ws_),
            _lib.int_array(strides), len(hs), _stream_ptr(dev))
    _lib.check(rc, "plyolo_yolox_loss_backward_f32")
    return grads


def _l1_inputs(ori, labels, fg_mask, matched_gt, hw):
    o = _check_cuda_f32(ori, "ori")
    lab = _check_cuda_f32(labels, "labels")
    if o.dim() != 3 or o.shape[2] != 4 or lab.dim() != 3 or lab.shape[2] != 5 or lab.shape[0] != o.shape[0]:
        raise ValueError("ori must be [B, A, 4] and labels [B, Lmax, 5]")
    if fg_mask.shape != o.shape[:2] or matched_gt.shape != o.shape[:2]:
        raise ValueError("fg_mask / matched_gt must be [B, A]")
    if fg_mask.dtype not in (torch.bool, torch.uint8) or matched_gt.dtype != torch.int32:
        raise TypeError("fg_mask must be bool / uint8 and matched_gt int32")
    hs, ws_ = list(hw[0::2]), list(hw[1::2])
    if sum(h * w for h, w in zip(hs, ws_)) != o.shape[1]:
        raise ValueError("the level shapes do not add up to A")
    return o, lab, fg_mask.contiguous(), matched_gt.contiguous(), hs, ws_


def yolox_l1_sum_raw(ori: torch.Tensor, labels: torch.Tensor, fg_mask: torch.Tensor, matched_gt: torch.Tensor, hw: List[int],
                     strides: List[int]) -> torch.Tensor:
    """use_l1 term (yolox_loss.py:128-133, :158): -> [1] = sum over the foreground anchors of |ori - get_l1_type(matched GT)|."""
    o, lab, fg, mg, hs, ws_ = _l1_inputs(ori, labels, fg_mask, matched_gt, hw)
    B, A, _ = o.shape
    dev = o.device
    out = torch.empty((1,), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        ws = _workspace("loss", L.plyolo_yolox_loss_workspace_bytes(B, A), dev)
        rc = L.plyolo_yolox_l1_f32(o.data_ptr(), lab.data_ptr(), fg.data_ptr(), mg.data_ptr(), B, lab.shape[1], _lib.int_array(hs),
                                   _lib.int_array(ws_), _lib.int_array(strides), len(hs), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                   _stream_ptr(dev))
    _lib.check(rc, "plyolo_yolox_l1_f32")
    return out


def yolox_l1_backward_raw(ori: torch.Tensor, labels: torch.Tensor, fg_mask: torch.Tensor, matched_gt: torch.Tensor,
                          grad_sum: torch.Tensor, grads: List[torch.Tensor], hw: List[int], strides: List[int]) -> None:
    """Adds grad_sum * sign(ori - target) to the regression planes of the head-map gradients `grads` (in place)."""
    o, lab, fg, mg, hs, ws_ = _l1_inputs(ori, labels, fg_mask, matched_gt, hw)
    g = _check_cuda_f32(grad_sum, "grad_sum")
    if g.numel() != 1:
        raise ValueError("grad_sum must have 1 element")
    B = o.shape[0]
    for t, h, w in zip(grads, hs, ws_):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous() or t.shape[0] != B or t.shape[2:] != (h, w):
            raise ValueError("grads must be contiguous fp32 CUDA tensors [B, 5+C, H, W], one per level")
    dev = o.device
    with torch.cuda.device(dev):
        rc = _lib.lib().plyolo_yolox_l1_backward_f32(
            o.data_ptr(), lab.data_ptr(), fg.data_ptr(), mg.data_ptr(), B, int(grads[0].shape[1]) - 5, lab.shape[1], g.data_ptr(),
            _lib.ptr_array([t.data_ptr() for t in grads]), _lib.int_array(hs), _lib.int_array(ws_), _lib.int_array(strides),
            len(hs), _stream_ptr(dev))
    _lib.check(rc, "plyolo_yolox_l1_backward_f32")


yolox_loss_sums = torch.library.custom_op("plyolo::yolox_loss_sums", yolox_loss_sums_raw, mutates_args=())


@yolox_loss_sums.register_fake
def _(preds, labels, fg_mask, matched_gt, matched_iou):
    return preds.new_empty((3,))


# ------------------------------------------------------------------------------------- format_dets
def format_dets_raw(dets: torch.Tensor, counts: torch.Tensor, inv_scales: torch.Tensor) -> torch.Tensor:
    """Device part of format_outputs: [B,max_det,6] + counts + per-image (float)(1/scale) -> [B,max_det,8] rows
    (x1,y1,x2,y2,w,h,score,class), boxes rescaled exactly as `bboxes /= scale` does on CUDA tensors."""
    d = _check_cuda_f32(dets, "dets")
    s_ = _check_cuda_f32(inv_scales, "inv_scales")
    if d.dim() != 3 or d.shape[2] != 6 or counts.dtype != torch.int32 or not counts.is_cuda:
        raise ValueError("dets must be [B,max_det,6] fp32 and counts [B] int32, both CUDA")
    B, max_det, _ = d.shape
    out = torch.empty((B, max_det, 8), dtype=torch.float32, device=d.device)
    with torch.cuda.device(d.device):
        rc = _lib.lib().plyolo_format_dets_f32(d.data_ptr(), counts.contiguous().data_ptr(), s_.data_ptr(), B, max_det,
                                               out.data_ptr(), _stream_ptr(d.device))
    _lib.check(rc, "plyolo_format_dets_f32")
    return out


format_dets = torch.library.custom_op("plyolo::format_dets", format_dets_raw, mutates_args=())


@format_dets.register_fake
def _(dets, counts, inv_scales):
    return dets.new_empty((dets.shape[0], dets.shape[1], 8))


# --------------------------------------------------------------------------------------- bboxes_iou
def bboxes_iou_raw(bboxes_a: torch.Tensor, bboxes_b: torch.Tensor, xyxy: bool) -> torch.Tensor:
    a = _check_cuda_f32(bboxes_a, "bboxes_a")
    b = _check_cuda_f32(bboxes_b, "bboxes_b")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        rc = _lib.lib().plyolo_bboxes_iou_f32(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], int(xyxy),
                                              out.data_ptr(), _stream_ptr(a.device))
    _lib.check(rc, "plyolo_bboxes_iou_f32")
    return out


bboxes_iou = torch.library.custom_op("plyolo::bboxes_iou", bboxes_iou_raw, mutates_args=())


@bboxes_iou.register_fake
def _(bboxes_a, bboxes_b, xyxy):
    return bboxes_a.new_empty((bboxes_a.shape[0], bboxes_b.shape[0]))
