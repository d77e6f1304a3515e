"""The reference-facing Python interface (same names / signatures as the reference modules), on GPU."""
import numpy as np
import pytest
import torch

from golden_util import eval_preds_of, heads_of, labels_of, load, names
from oracle import torch_ops_replay as R
from pl_yolo_b200 import LazyPredictions, YOLOXDecoder, YOLOXLoss, bboxes_iou, postprocess, synth

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


def test_bboxes_iou_guard():
    with pytest.raises(IndexError):
        bboxes_iou(torch.zeros(2, 5, device=DEV), torch.zeros(3, 4, device=DEV))
