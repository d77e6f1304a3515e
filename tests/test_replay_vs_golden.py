"""The torch-op replay (oracle/torch_ops_replay.py) is the stand-in for "the reference on CUDA" on the
GPU box.  Here, on CPU, it must reproduce the REAL reference's outputs bit-for-bit (same ATen /
torchvision binaries, same op order) — that is what licenses using it on the B200."""
import numpy as np
import pytest
import torch

from golden_util import eval_preds_of, heads_of, labels_of, load, names
from oracle import torch_ops_replay as R
from pl_yolo_b200 import synth


@pytest.mark.parametrize("name", names("decode"))
def test_decode_replay(name):
    meta, g = load(name)
    heads = [torch.from_numpy(h) for h in heads_of(meta)]
    rows = g["rows"]
    tr, ori = R.decode(heads, meta["strides"], inference=False)
    ev, _ = R.decode(heads, meta["strides"], inference=True)
    assert np.array_equal(tr[..., :4].numpy(), g["train_boxes"])
    assert np.array_equal(tr[:, rows].numpy(), g["train_rows"])
    assert np.array_equal(ori[:, rows].numpy(), g["ori_rows"])
    assert np.array_equal(ev[..., :4].numpy(), g["eval_boxes"])
    assert np.array_equal(ev[:, rows].numpy(), g["eval_rows"])
    xs, ys, es = R.anchor_grid(synth.level_shapes(meta["size"]), meta["strides"], heads[0])
    assert np.array_equal(xs.numpy(), g["x_shifts"]) and np.array_equal(ys.numpy(), g["y_shifts"])
    assert np.array_equal(es.numpy(), g["expanded_strides"])


@pytest.mark.parametrize("name", names("post"))
def test_postprocess_replay(name):
    meta, g = load(name)
    p = torch.from_numpy(eval_preds_of(meta))
    outs = R.postprocess(p, meta["conf"], meta["nms"], meta["agnostic"])
    for b, o in enumerate(outs):
        n = 0 if o is None else o.shape[0]
        assert n == g["counts"][b]
        if n:
            assert np.array_equal(o.numpy(), g["dets"][b, :n])


@pytest.mark.parametrize("name", names("simota"))
def test_simota_replay(name):
    meta, g = load(name)
    heads = heads_of(meta)
    labels = labels_of(meta, g)
    preds = torch.from_numpy(synth.make_train_preds(heads, g["ref_boxes"]))
    o = R.simota(preds, torch.from_numpy(labels), synth.level_shapes(meta["size"]), meta["strides"], stable=False)
    assert np.array_equal(o["num_gt"].numpy(), g["num_gt"])
    assert np.array_equal(o["n_cand"].numpy(), g["n_cand"])
    assert np.array_equal(o["dyn_k"].numpy(), g["dyn_k"])
    assert np.array_equal(o["fg_mask"].numpy().astype(np.uint8), g["fg_mask"])
    assert np.array_equal(o["matched_gt"].numpy(), g["matched_gt"])
    assert np.array_equal(o["matched_iou"].numpy(), g["matched_iou"])
    assert np.array_equal(o["num_fg"].numpy(), g["num_fg"])


@pytest.mark.parametrize("name", names("sib_"))
def test_sibling_nms_replay_matches_the_real_decoders(name):
    """oracle/torch_ops_replay.py yolov3_nms / yolov5_nms vs the real YOLOv3Decoder / YOLOv5Decoder outputs (CPU, bit for bit)."""
    meta, g = load(name)
    p = torch.from_numpy(g["predictions"])
    if meta["kind"] == "sib_v3":
        outs = R.yolov3_nms(p, meta["conf"], meta["nms"], meta["max_nms"], meta["max_det"])
    else:
        outs = R.yolov5_nms(p, meta["conf"], meta["nms"], meta["agnostic"])
    for i, o in enumerate(outs):
        c = int(g["counts"][i])
        assert (0 if o is None else o.shape[0]) == c
        if c:
            assert np.array_equal(o.numpy(), g["dets"][i, :c])
