"""The reference-facing Python interface (same names / signatures as the reference modules), on GPU."""
import numpy as np
import pytest
import torch

from golden_util import eval_preds_of, heads_of, labels_of, load, names
from oracle import torch_ops_replay as R
from pl_yolo_b200 import LazyPredictions, YOLOXDecoder, YOLOXLoss, bboxes_iou, ops, postprocess, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STRIDES = [8, 16, 32]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("name", names("post"))
def test_postprocess_list_structure_matches_reference(name):
    meta, g = load(name)
    p = cu(eval_preds_of(meta))
    outs = postprocess(p, meta["conf"], meta["nms"], meta["agnostic"])
    ref = R.postprocess(p, meta["conf"], meta["nms"], meta["agnostic"])  # reference op chain on CUDA
    assert len(outs) == len(ref)
    for o, r in zip(outs, ref):
        assert (o is None) == (r is None)
        if o is not None:
            assert o.shape == r.shape and torch.equal(o, r)


def test_eval_call_and_decoder_and_lazy_route():
    heads = [cu(h) for h in synth.make_heads(3, 320, 80, 9)]
    loss = YOLOXLoss(80, STRIDES).eval()
    preds = loss([h.clone() for h in heads], torch.zeros(3, 1, 5, device=DEV))
    assert torch.equal(preds, R.decode(heads, STRIDES, True)[0])
    assert torch.equal(YOLOXDecoder(80, STRIDES)(heads), preds)
    lazy = YOLOXLoss(80, STRIDES, lazy_eval=True).eval()(heads, None)
    assert isinstance(lazy, LazyPredictions) and lazy.shape == preds.shape
    a = postprocess(lazy, 0.01, 0.65)
    b = postprocess(preds, 0.01, 0.65)
    for x, y in zip(a, b):
        assert (x is None and y is None) or torch.equal(x, y)
    assert torch.equal(lazy[1], preds[1])


@pytest.mark.parametrize("name", ["simota_640_b4", "simota_320_b8", "simota_edges"])
def test_training_loss_dict_matches_reference(name):
    meta, g = load(name)
    heads = [cu(h) for h in heads_of(meta)]
    labels = cu(labels_of(meta, g))
    out = YOLOXLoss(80, STRIDES)([h.clone() for h in heads], labels)
    ref = meta["losses"]  # the real reference's loss dict for these inputs (CPU)
    assert set(out.keys()) == set(ref.keys())
    for k in ref:
        assert float(out[k]) == pytest.approx(ref[k], rel=2e-5, abs=1e-6), k


def test_decode_backward_matches_autograd_of_op_chain():
    heads = [cu(h) for h in synth.make_heads(2, 160, 80, 10)]
    a = [h.clone().requires_grad_(True) for h in heads]
    b = [h.clone().requires_grad_(True) for h in heads]
    loss = YOLOXLoss(80, STRIDES)
    p, ori, xs, ys, es = loss.decode(a)
    rp, rori = R.decode(b, STRIDES, False)
    w1 = torch.randn_like(p)
    w2 = torch.randn_like(ori)
    ((p * w1).sum() + (ori * w2).sum()).backward()
    ((rp * w1).sum() + (rori * w2).sum()).backward()
    for x, y in zip(a, b):
        assert torch.allclose(x.grad, y.grad, rtol=1e-6, atol=1e-6)
    gx, gy, ge = R.anchor_grid(synth.level_shapes(160), STRIDES, heads[0])
    assert torch.equal(xs, gx) and torch.equal(ys, gy) and torch.equal(es, ge)


def test_training_step_backward_runs():
    heads = [cu(h).requires_grad_(True) for h in synth.make_heads(2, 160, 80, 11)]
    labels = cu(synth.make_labels(2, 160, 10, 80, 12))
    out = YOLOXLoss(80, STRIDES, use_l1=True)(heads, labels)
    out["loss"].backward()
    assert all(h.grad is not None and torch.isfinite(h.grad).all() for h in heads)


def _grad_close(a, b, what):
    scale = float(b.abs().max())
    err = float((a - b).abs().max())
    assert err <= 1e-5 * scale + 1e-12, "%s: max |diff| %.3e vs max |ref| %.3e" % (what, err, scale)


@pytest.mark.parametrize("B,size,lmax,seed", [(2, 160, 10, 30), (4, 320, 40, 31), (8, 640, 120, 32)])
def test_fused_loss_tail_matches_torch_tail(B, size, lmax, seed):
    """N2: the fused loss tail (forward sums + backward straight into the head maps) against the batched torch
    code of the same shim (itself pinned to the real reference's loss dict by the golden fixtures): values within
    1e-5 relative, head-map gradients within 1e-5 of the largest gradient of the level."""
    heads = synth.make_heads(B, size, 80, seed)
    labels = cu(synth.make_labels(B, size, lmax, 80, seed + 100))
    a = [cu(h).requires_grad_(True) for h in heads]
    b = [cu(h).requires_grad_(True) for h in heads]
    fused = YOLOXLoss(80, STRIDES)(a, labels)
    plain = YOLOXLoss(80, STRIDES, fused_loss=False)(b, labels)
    for k in ("loss", "loss_iou", "loss_obj", "loss_cls", "proportion"):
        assert float(torch.as_tensor(fused[k]).detach()) == pytest.approx(float(torch.as_tensor(plain[k]).detach()), rel=1e-5, abs=1e-7), k
    assert fused["loss_l1"] == 0.0
    fused["loss"].backward()
    plain["loss"].backward()
    for l, (x, y) in enumerate(zip(a, b)):
        assert x.grad.shape == y.grad.shape
        _grad_close(x.grad, y.grad, "level %d" % l)
        for lo, hi, nm in ((0, 4, "box"), (4, 5, "obj"), (5, 85, "cls")):
            if float(y.grad[:, lo:hi].abs().max()) > 0:
                _grad_close(x.grad[:, lo:hi], y.grad[:, lo:hi], "level %d %s" % (l, nm))


@pytest.mark.parametrize("name", names("lossgrad"))
def test_fused_loss_vs_real_reference_gradients(name):
    """N2 against the REAL reference: loss dict and autograd gradients of the head outputs recorded from
    /root/reference's YOLOXLoss on CPU (oracle/gen_golden.py lossgrad)."""
    meta, g = load(name)
    heads = [cu(h).requires_grad_(True) for h in heads_of(meta)]
    labels = cu(labels_of(meta, g))
    out = YOLOXLoss(80, STRIDES, use_l1=bool(meta.get("use_l1", False)))(heads, labels)
    for k, v in meta["losses"].items():
        assert float(out[k]) == pytest.approx(v, rel=2e-5, abs=1e-6), k
    out["loss"].backward()
    for l, h in enumerate(heads):
        _grad_close(h.grad, cu(g["grad%d" % l]), "%s level %d" % (name, l))


def test_fused_l1_term_vs_torch_tail():
    """use_l1=True: the fused kernels (sum + sign gradient into the raw regression planes) against the batched torch tail,
    the L1 term alone and inside the total loss; an image without GTs; the stand-alone ops' argument checks."""
    B, size = 3, 320
    heads = synth.make_heads(B, size, 80, 61)
    lab = synth.make_labels(B, size, 25, 80, 62)
    lab[2] = 0
    labels = cu(lab)
    for pick in ("loss_l1", "loss"):
        a = [cu(h).requires_grad_(True) for h in heads]
        b = [cu(h).requires_grad_(True) for h in heads]
        f = YOLOXLoss(80, STRIDES, use_l1=True)(a, labels)
        g = YOLOXLoss(80, STRIDES, use_l1=True, fused_loss=False)(b, labels)
        for k in ("loss", "loss_iou", "loss_obj", "loss_cls", "loss_l1", "proportion"):
            assert float(torch.as_tensor(f[k]).detach()) == pytest.approx(float(torch.as_tensor(g[k]).detach()), rel=1e-5, abs=1e-7), k
        assert float(f["loss_l1"]) > 0
        f[pick].backward()
        g[pick].backward()
        for l, (x, y) in enumerate(zip(a, b)):
            _grad_close(x.grad, y.grad, "%s level %d" % (pick, l))
    from pl_yolo_b200 import ops
    preds, ori = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    hw = [v for s in STRIDES for v in (size // s, size // s)]
    fg, mg, _, _, _ = ops.simota_assign_raw(preds, labels, hw, STRIDES)
    with pytest.raises(ValueError):
        ops.yolox_l1_sum_raw(ori[:, :-1], labels, fg, mg, hw, STRIDES)
    with pytest.raises(TypeError):
        ops.yolox_l1_sum_raw(ori.double(), labels, fg, mg, hw, STRIDES)


@pytest.mark.parametrize("C", [1, 20, 91])
def test_fused_loss_other_class_counts(C):
    """N2 with class counts that are not a multiple of the warp width (VOC has 20)."""
    heads = synth.make_heads(3, 320, C, 90 + C)
    labels = cu(synth.make_labels(3, 320, 20, C, 190 + C))
    a = [cu(h).requires_grad_(True) for h in heads]
    b = [cu(h).requires_grad_(True) for h in heads]
    fused = YOLOXLoss(C, STRIDES)(a, labels)
    plain = YOLOXLoss(C, STRIDES, fused_loss=False)(b, labels)
    for k in ("loss", "loss_iou", "loss_obj", "loss_cls", "proportion"):
        assert float(torch.as_tensor(fused[k]).detach()) == pytest.approx(float(torch.as_tensor(plain[k]).detach()), rel=1e-5, abs=1e-7), k
    fused["loss"].backward()
    plain["loss"].backward()
    for l, (x, y) in enumerate(zip(a, b)):
        _grad_close(x.grad, y.grad, "C=%d level %d" % (C, l))


def test_fused_loss_component_gradients_and_edges():
    """Each of the three sums separately (different upstream gradients), an image without GTs, and the C-ABI errors."""
    B, size = 3, 160
    heads = synth.make_heads(B, size, 80, 40)
    lab = synth.make_labels(B, size, 12, 80, 41)
    lab[1] = 0  # no GT in image 1: every anchor background, only the objectness term
    labels = cu(lab)
    for w in ([1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.3, 2.0, 0.7]):
        a = [cu(h).requires_grad_(True) for h in heads]
        b = [cu(h).requires_grad_(True) for h in heads]
        f = YOLOXLoss(80, STRIDES)(a, labels)
        g = YOLOXLoss(80, STRIDES, fused_loss=False)(b, labels)
        (w[0] * f["loss_iou"] + w[1] * f["loss_obj"] + w[2] * f["loss_cls"]).backward()
        (w[0] * g["loss_iou"] + w[1] * g["loss_obj"] + w[2] * g["loss_cls"]).backward()
        for l, (x, y) in enumerate(zip(a, b)):
            _grad_close(x.grad, y.grad, "weights %s level %d" % (w, l))
    preds, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    fg, mg, mi, nfg, ngt = ops.simota_assign_raw(preds, labels, [20, 20, 10, 10, 5, 5], STRIDES)
    sums = ops.yolox_loss_sums_raw(preds, labels, fg, mg, mi)
    assert sums.shape == (3,) and torch.isfinite(sums).all()
    assert torch.equal(sums, ops.yolox_loss_sums_raw(preds, labels, fg, mg, mi)), "the sums must be deterministic"
    with pytest.raises(Exception):
        ops.yolox_loss_sums_raw(preds.cpu(), labels, fg, mg, mi)
    with pytest.raises(Exception):
        ops.yolox_loss_backward_raw(preds, labels, fg, mg, mi, sums, [20, 20, 10, 10], STRIDES[:2])


def test_bboxes_iou_guard():
    with pytest.raises(IndexError):
        bboxes_iou(torch.zeros(2, 5, device=DEV), torch.zeros(3, 4, device=DEV))


def _ref_format_outputs(outputs, ids, hws, val_size, class_ids):
    """models/evaluators/postprocess.py:95-138 replayed op for op on the tensors' own device."""
    json_list = []
    det_list = [[np.empty(shape=[0, 5]) for _ in range(len(class_ids))] for _ in range(len(outputs))]
    for i, (output, img_h, img_w, img_id) in enumerate(zip(outputs, hws[0], hws[1], ids)):
        if output is None:
            continue
        bboxes = output[:, 0:4]
        scale = min(val_size[0] / float(img_w), val_size[1] / float(img_h))
        bboxes /= scale
        coco = bboxes.clone()
        coco[:, 2] = bboxes[:, 2] - bboxes[:, 0]
        coco[:, 3] = bboxes[:, 3] - bboxes[:, 1]
        scores, clses = output[:, 4], output[:, 5]
        for cocobox, score, cls in zip(coco, scores, clses):
            json_list.append({"image_id": int(img_id), "category_id": class_ids[int(cls)],
                              "bbox": cocobox.cpu().numpy().tolist(), "score": score.cpu().numpy().item(), "segmentation": []})
        for c in range(len(class_ids)):
            det_list[i][c] = output[clses == c, 0:5].cpu().numpy()
    return json_list, det_list


def test_format_outputs_matches_reference_op_chain():
    """N1: format_outputs — one kernel + one D2H instead of a .cpu() per detection; results and the
    in-place rescale of `outputs` are bit-identical to the reference's op chain on CUDA tensors."""
    from pl_yolo_b200 import format_outputs
    B = 5
    heads = [cu(h) for h in synth.make_heads(B, 320, 80, 61)]
    heads[0][3, 4] = -30.0; heads[1][3, 4] = -30.0; heads[2][3, 4] = -30.0  # image 3: nothing passes -> None
    preds, _ = ops.decode_raw(heads, [8, 16, 32], True)
    ids = [11, 12, 13, 14, 15]
    hws = (torch.tensor([480, 333, 1000, 320, 517]), torch.tensor([640, 500, 1777, 320, 389]))
    class_ids = list(range(1, 81))
    ours = postprocess(preds, 0.3, 0.65)
    assert ours[3] is None
    ref_in = [None if o is None else o.clone() for o in ours]
    jl, dl = format_outputs(ours, ids, hws, (320, 320), class_ids, None)
    rj, rd = _ref_format_outputs(ref_in, ids, hws, (320, 320), class_ids)
    assert jl == rj
    for a, b in zip(dl, rd):
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y)
    for o, r in zip(ours, ref_in):  # the in-place side effect on `outputs`
        assert (o is None) == (r is None)
        if o is not None:
            assert torch.equal(o, r)
    # a plain list (not produced by our postprocess) takes the padding route
    plain = [None if o is None else o.clone() for o in ref_in]
    jl2, _ = format_outputs(plain, ids, hws, (1.0, 1.0), class_ids, None)
    rj2, _ = _ref_format_outputs([None if o is None else o.clone() for o in ref_in], ids, hws, (1.0, 1.0), class_ids)
    assert jl2 == rj2


# ---- the stand-alone boundary functions, pinned to goldens produced by the REAL reference (oracle/gen_golden_boundary.py)
def test_get_in_boxes_info_vs_real_reference():
    from pl_yolo_b200 import get_in_boxes_info
    _, g = load("ref_in_boxes_160")
    gt, xs, ys, es = cu(g["gt"]), cu(g["x_shifts"]), cu(g["y_shifts"]), cu(g["expanded_strides"])
    fg, both = get_in_boxes_info(gt, es, xs, ys, xs.shape[1], gt.shape[0])     # yolox_loss.py:231 signature
    assert fg.dtype == torch.bool and both.dtype == torch.bool
    assert np.array_equal(fg.cpu().numpy(), g["fg_mask"])
    assert both.shape == tuple(g["both"].shape) and np.array_equal(both.cpu().numpy(), g["both"])
    # and against the op-for-op replay on CUDA tensors
    rfg, rboth = R.geometry_prior(gt, es, xs, ys)
    assert torch.equal(fg, rfg) and torch.equal(both, rboth)
    # no GT at all: nothing is foreground
    fg0, both0 = get_in_boxes_info(gt[:0], es, xs, ys, xs.shape[1], 0)
    assert not fg0.any() and both0.shape == (0, 0)


def test_dynamic_k_matching_vs_real_reference():
    from pl_yolo_b200 import dynamic_k_matching
    meta, g = load("ref_dynk")
    for c in meta["cases"]:
        cost, ious, cls = cu(g[c + "_cost"]), cu(g[c + "_ious"]), cu(g[c + "_cls"])
        fg = torch.from_numpy(g[c + "_fg_in"].copy()).to(DEV)
        out = dynamic_k_matching(fg, cost, ious, cls, cost.shape[0])           # yolox_loss.py:318 signature
        assert out[0] is fg, "fg_mask must be updated in place (yolox_loss.py:361)"
        assert np.array_equal(fg.cpu().numpy(), g[c + "_fg_out"]), c
        assert int(out[1]) == int(g[c + "_num_fg"]), c
        assert np.array_equal(out[2].cpu().numpy(), g[c + "_gt"]), c
        assert np.array_equal(out[3].cpu().numpy(), g[c + "_mcls"]), c
        assert np.array_equal(out[4].cpu().numpy(), g[c + "_iou"]), c          # bit-exact


def test_dynamic_k_matching_random_vs_replay():
    """Seeded sweep against the reference op chain on CUDA tensors (stable tie rule == lowest index)."""
    rng = np.random.default_rng(3)
    for G, Nc in [(1, 5), (2, 11), (7, 64), (30, 700), (120, 5000)]:
        ious = cu((rng.uniform(0, 1, (G, Nc)) ** 3).astype(np.float32))
        cost = cu(np.round(rng.uniform(1, 30, (G, Nc)), 1).astype(np.float32))      # rounded: plenty of exact ties
        sel, mg, mi, dk, M = ops.dynamic_k_matching_raw(cost, ious)
        rM, rdk = R.dynamic_k(cost, ious, stable=True)
        assert torch.equal(dk, rdk.to(torch.int32)), (G, Nc)
        assert torch.equal(M, rM > 0), (G, Nc)
        rsel = rM.sum(0) > 0
        assert torch.equal(sel, rsel)
        assert torch.equal(mg[sel].long(), rM[:, rsel].argmax(0))
        assert torch.equal(mi[sel], (rM * ious).sum(0)[rsel])


def test_format_outputs_vs_real_reference():
    """N1 pinned to the real format_outputs (models/evaluators/postprocess.py:95-138, run on CPU tensors by
    oracle/gen_golden_boundary.py).  On CPU `bboxes /= scale` divides, on CUDA it multiplies by the fp32 reciprocal
    (what the kernel reproduces): boxes agree to 1 ulp-ish (2e-7 relative), everything else exactly."""
    from pl_yolo_b200 import format_outputs
    meta, g = load("ref_format_outputs")
    B = len(meta["counts"])
    outs = [None if meta["counts"][i] == 0 else cu(g["in_%d" % i]) for i in range(B)]
    json_list, det_list = format_outputs(outs, meta["ids"], meta["hws"], tuple(meta["val_size"]), meta["class_ids"], None)
    assert [j["image_id"] for j in json_list] == g["json_image_id"].tolist()
    assert [j["category_id"] for j in json_list] == g["json_category_id"].tolist()
    assert all(j["segmentation"] == [] for j in json_list)
    # w = x2 - x1 of two coordinates that each differ by at most one ulp of ~1e3 (6e-5): absolute tolerance
    np.testing.assert_allclose(np.array([j["bbox"] for j in json_list]), g["json_bbox"], rtol=3e-7, atol=2.5e-4)
    assert np.array_equal(np.array([j["score"] for j in json_list]), g["json_score"])
    for i in range(B):
        if outs[i] is not None:   # the reference's in-place rescale is visible to the caller
            np.testing.assert_allclose(outs[i].cpu().numpy(), g["scaled_%d" % i], rtol=3e-7, atol=1e-5)
        for c in range(len(meta["class_ids"])):
            want = g["det_%d_%d" % (i, c)]
            got = np.asarray(det_list[i][c])
            assert got.shape == want.shape, (i, c)
            if want.size:
                np.testing.assert_allclose(got, want, rtol=3e-7, atol=1e-5)


def test_lazy_predictions_torch_function_and_edges():
    heads = [cu(h) for h in synth.make_heads(2, 160, 80, 5)]
    ref = R.decode(heads, STRIDES, True)[0]
    lazy = YOLOXLoss(80, STRIDES, lazy_eval=True).eval()(heads, None)
    assert torch.equal(torch.max(lazy, dim=2).values, ref.max(dim=2).values)       # __torch_function__ materialises
    assert torch.equal(torch.cat([lazy, lazy], 0), torch.cat([ref, ref], 0))
    assert torch.equal(lazy[1, :5], ref[1, :5])


def test_fused_training_loss_without_label_rows():
    """labels [B,0,5]: the reference returns the objectness-only loss (yolox_loss.py:57-62); so does the fused path."""
    heads = [cu(h).requires_grad_(True) for h in synth.make_heads(2, 160, 80, 6)]
    empty = torch.zeros((2, 0, 5), device=DEV)
    fused = YOLOXLoss(80, STRIDES).train()(heads, empty)
    plain = YOLOXLoss(80, STRIDES, fused_loss=False).train()([h.detach() for h in heads], torch.zeros((2, 1, 5), device=DEV))
    assert float(fused["loss_iou"]) == 0.0 and float(fused["loss_cls"]) == 0.0
    assert float(fused["loss"]) == pytest.approx(float(plain["loss"]), rel=1e-5)
    fused["loss"].backward()
    assert all(h.grad is not None and torch.isfinite(h.grad).all() for h in heads)


def test_format_outputs_short_arguments_and_edited_list():
    from pl_yolo_b200 import format_outputs
    p = cu(synth.make_eval_preds(3, 525, 80, 2, size=160.0))
    outs = postprocess(p, 0.3, 0.5)
    ref = [None if o is None else o.clone() for o in outs]
    # fewer image sizes than outputs: zip() semantics, only the first two images are visited
    j, d = format_outputs(outs, [1, 2], [[100, 200], [300, 150]], (160, 160), list(range(80)))
    assert {x["image_id"] for x in j} <= {1, 2} and len(d) == 3
    assert outs[2] is None or torch.equal(outs[2], ref[2])                          # untouched
    # a caller that filtered an entry since postprocess(): the cached dense result must not be trusted
    outs2 = postprocess(p, 0.3, 0.5)
    if outs2[0] is not None and outs2[0].shape[0] > 1:
        outs2[0] = outs2[0][:1].clone()
        j2, _ = format_outputs(outs2, [1, 2, 3], [[100, 200, 50], [300, 150, 60]], (160, 160), list(range(80)))
        assert sum(1 for x in j2 if x["image_id"] == 1) == 1
