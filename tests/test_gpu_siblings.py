"""Sibling heads on the GPU: the YOLOv3 / YOLOv5 decoders' NMS call sites and YOLOv7's matching block."""
import numpy as np
import pytest
import torch

from golden_util import load, names
from oracle import torch_ops_replay as R
from pl_yolo_b200 import _lib, ops, synth
from pl_yolo_b200.sibling_heads import yolov3_nms, yolov5_nms, yolov7_matching

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def same_lists(got, want, none_ok=True):
    assert len(got) == len(want)
    for o, r in zip(got, want):
        n_o = 0 if o is None else o.shape[0]
        n_r = 0 if r is None else r.shape[0]
        assert n_o == n_r
        if n_r:
            assert torch.equal(o, r)


@pytest.mark.parametrize("name", names("sib_"))
def test_sibling_nms_vs_real_decoder_goldens_and_cuda_replay(name):
    meta, g = load(name)
    p = cu(g["predictions"])
    if meta["kind"] == "sib_v3":
        outs = yolov3_nms(p, meta["conf"], meta["nms"], meta["max_nms"], meta["max_det"])
        ref = R.yolov3_nms(p, meta["conf"], meta["nms"], meta["max_nms"], meta["max_det"])   # same op chain, CUDA tensors
    else:
        outs = yolov5_nms(p, meta["conf"], meta["nms"], meta["agnostic"])
        ref = R.yolov5_nms(p, meta["conf"], meta["nms"], meta["agnostic"])
    same_lists(outs, ref)                                       # bit-exact vs the reference ops on the same device
    for i, o in enumerate(outs):                                # and vs what the REAL class returned (CPU run)
        c = int(g["counts"][i])
        assert (0 if o is None else o.shape[0]) == c
        if c:
            np.testing.assert_array_equal(o.cpu().numpy(), g["dets"][i, :c])


def test_sibling_nms_seeded_sweep_vs_cuda_replay():
    for seed, (B, N, C) in enumerate([(2, 700, 20), (1, 3000, 80), (4, 1200, 3), (2, 25200, 80)]):
        x = synth.make_eval_preds(B, N, C, 40 + seed, size=640.0, n_clusters=8, p_obj=0.3)
        p = x.copy()
        p[..., 0], p[..., 1] = (x[..., 0] + x[..., 2]) / 2, (x[..., 1] + x[..., 3]) / 2
        p[..., 2], p[..., 3] = x[..., 2] - x[..., 0], x[..., 3] - x[..., 1]
        p = cu(p.astype(np.float32))
        for conf, nms in [(0.05, 0.45), (0.4, 0.6)]:
            same_lists(yolov3_nms(p, conf, nms), R.yolov3_nms(p, conf, nms))
            same_lists(yolov5_nms(p, conf, nms), R.yolov5_nms(p, conf, nms))
            same_lists(yolov5_nms(p, conf, nms, agnostic=True), R.yolov5_nms(p, conf, nms, agnostic=True))
        # score-sorted truncation and a small max_det
        same_lists(yolov3_nms(p, 0.02, 0.5, max_nms=100, max_det=30), R.yolov3_nms(p, 0.02, 0.5, max_nms=100, max_det=30))
        same_lists(yolov5_nms(p, 0.02, 0.5, max_nms=150, max_det=40), R.yolov5_nms(p, 0.02, 0.5, max_nms=150, max_det=40))
    # nothing passes: v3 gives None entries, v5 empty [0,7] tensors
    assert yolov3_nms(p, 2.0, 0.5) == [None] * p.shape[0]
    assert all(o.shape == (0, 7) for o in yolov5_nms(p, 2.0, 0.5))
    with pytest.raises(NotImplementedError):
        yolov5_nms(p, 0.3, 0.5, multi_label=True)


def test_yolov7_matching_vs_reference_ops():
    rng = np.random.default_rng(12)
    for G, N in [(1, 3), (3, 8), (9, 120), (40, 2000)]:
        iou = cu((rng.uniform(0, 1, (G, N)) ** 3).astype(np.float32))
        cost = cu(rng.uniform(1, 30, (G, N)).astype(np.float32))
        if G > 1:
            cost[1] = cost[0] + 1e-3   # two GTs that want the same anchors: conflicts
        sel, gt = yolov7_matching(cost, iou)
        rsel, rgt, rdk = R.yolov7_matching(cost, iou)
        assert torch.equal(sel, rsel) and torch.equal(gt, rgt), (G, N)
        dk = ops.dynamic_k_matching_raw(cost, iou, exact_k=True)[3]
        assert torch.equal(dk, rdk.to(torch.int32))


# ---- N4: VOC evaluator statistics on the device, pinned to the REAL tpfp_default / average_precision
def test_voc_tpfp_vs_real_reference():
    from pl_yolo_b200.eval_voc import tpfp_default, voc_ap, voc_tpfp_dense
    meta, g = load("ref_voc_tpfp")
    dets, counts = cu(g["dets"]), cu(g["counts"])
    gts, gcnt = cu(g["gts"]), cu(g["gt_counts"])
    tp, ngt = voc_tpfp_dense(dets, counts, gts, gcnt, meta["iou_thr"], meta["C"])
    assert np.array_equal(tp.cpu().numpy().astype(np.uint8), g["tp"])            # every TP / FP decision of the real function
    assert np.array_equal(ngt.cpu().numpy(), g["num_gts"])
    res = voc_ap(dets, counts, tp, ngt)
    ap = np.array([r["ap"] for r in res["results"]], np.float32)
    assert np.array_equal(ap, g["ap"])                                           # same float32 AP per class
    assert res["mean_ap"] == pytest.approx(float(g["ap"][g["num_gts"] > 0].mean()), rel=1e-6)
    # the reference-signature wrapper on one (image, class) pair
    b, c = 0, int(g["dets"][0, 0, 5])
    n = int(g["counts"][b])
    m = g["dets"][b, :n, 5].astype(int) == c
    gm = g["gts"][b, : g["gt_counts"][b]]
    t, f = tpfp_default(g["dets"][b, :n][m][:, :5], gm[gm[:, 4].astype(int) == c][:, :4], 0.5)
    assert np.array_equal(t.astype(np.uint8), g["tp"][b, :n][m]) and np.array_equal(f, 1 - t)
    t0, f0 = tpfp_default(g["dets"][b, :5, :5], np.zeros((0, 4), np.float32), 0.5)       # no GT of the class: all FP (:81-83)
    assert not t0.any() and f0.all()
