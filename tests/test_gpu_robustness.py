"""SURVEY T12 (NaN / Inf / zero-area inputs: "unspecified, must not hang or write out of bounds"), property tests over
randomly drawn shapes / thresholds (hypothesis), and the paths the class-split NMS kernel hands to the general one."""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle
from oracle import torch_ops_replay as R
from pl_yolo_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
STRIDES = [8, 16, 32]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _poison(arr, rng, frac=0.002):
    a = arr.copy()
    flat = a.reshape(-1)
    idx = rng.choice(flat.size, max(3, int(flat.size * frac)), replace=False)
    flat[idx] = rng.choice([np.nan, np.inf, -np.inf, 1e30, -1e30, 0.0], idx.size).astype(np.float32)
    return a


def test_t12_nan_inf_heads_do_not_hang_or_corrupt():
    rng = np.random.default_rng(0)
    for seed in range(3):
        heads = [_poison(h, rng) for h in synth.make_heads(4, 160, 80, seed)]
        guard = torch.full((4 * 300 * 6 + 64,), 7.0, device=DEV)
        dets = guard[:4 * 300 * 6].view(4, 300, 6)
        counts = torch.full((4,), -5, dtype=torch.int32, device=DEV)
        keep = torch.full((4, 300), -7, dtype=torch.int32, device=DEV)
        ops.decode_postprocess_raw([cu(h) for h in heads], STRIDES, 0.01, 0.65, False, 10000, 300, 0, out=(dets, counts, keep))
        torch.cuda.synchronize()
        c = counts.cpu().numpy()
        assert ((c >= 0) & (c <= 300)).all()
        assert (guard[4 * 300 * 6:] == 7.0).all()                       # nothing written past the output
        k = keep.cpu().numpy()
        for b in range(4):
            assert (k[b, :c[b]] >= 0).all() and (k[b, :c[b]] < 525).all() and (k[b, c[b]:] == -1).all()
        # materialised predictions with NaN / Inf rows, class-aware and class-agnostic
        preds, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, True)
        for agn in (False, True):
            d, c2, _ = ops.postprocess_raw(preds, 0.01, 0.65, agn, 10000, 300, 0)
            torch.cuda.synchronize()
            assert ((c2 >= 0) & (c2 <= 300)).all()


def test_t12_simota_degenerate_inputs_do_not_hang():
    rng = np.random.default_rng(1)
    B, size, lmax = 3, 160, 12
    heads = synth.make_heads(B, size, 80, 3)
    labels = synth.make_labels(B, size, lmax, 80, 4)
    labels[0, 0, 3:5] = 0.0            # zero-area GT
    labels[0, 1, 1:3] = [1e6, -1e6]    # far outside
    labels[1, 0, 1:5] = np.nan         # NaN GT
    labels[2, 0, 3:5] = [np.inf, 5.0]
    preds, _ = ops.decode_raw([cu(_poison(h, rng)) for h in heads], STRIDES, False)
    hw = [v for s in synth.level_shapes(size) for v in s]
    fg, mg, mi, nfg, ngt = ops.simota_assign_raw(preds, cu(labels), hw, STRIDES)
    torch.cuda.synchronize()
    A = preds.shape[1]
    assert fg.shape == (B, A) and int(nfg.min()) >= 0 and int(nfg.max()) <= A
    m = mg.cpu().numpy()
    assert ((m >= -1) & (m < lmax)).all()
    assert np.array_equal(fg.cpu().numpy(), m >= 0)                      # the mask and the matches stay consistent


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(B=st.integers(1, 4), size=st.sampled_from([96, 128, 160, 224, 320]), C=st.sampled_from([1, 2, 7, 20, 80, 91]),
       conf=st.sampled_from([0.001, 0.01, 0.1, 0.3, 0.7]), nms=st.sampled_from([0.0, 0.3, 0.45, 0.65, 0.9, 1.0]),
       seed=st.integers(0, 1000), objs=st.integers(0, 25), agn=st.booleans())
def test_property_fused_postprocess_equals_reference_ops(B, size, C, conf, nms, seed, objs, agn):
    """decode + postprocess from the head maps == the reference's op chain on the same GPU (bit-exact), any shape."""
    heads = [cu(h) for h in synth.make_heads(B, size, C, seed, objects_per_image=objs)]
    d, c, k = ops.decode_postprocess_raw(heads, STRIDES, conf, nms, agn, 10000, 300, 0)
    preds, _ = R.decode(heads, STRIDES, True)
    ref = R.postprocess(preds, conf, nms, agn)
    cn = c.cpu().numpy()
    for b in range(B):
        n = 0 if ref[b] is None else ref[b].shape[0]
        assert cn[b] == n, (b, cn[b], n)
        if n:
            assert torch.equal(d[b, :n], ref[b])
        assert float(d[b, n:].abs().max()) == 0.0 if n < 300 else True     # zero padding
    # the same through the materialised predictions
    d2, c2, _ = ops.postprocess_raw(ops.decode_raw(heads, STRIDES, True)[0], conf, nms, agn, 10000, 300, 0)
    assert torch.equal(c2, c) and torch.equal(d2, d)


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(B=st.integers(1, 3), size=st.sampled_from([96, 160, 320]), lmax=st.sampled_from([1, 5, 40, 120]), seed=st.integers(0, 500),
       tiny=st.floats(0.0, 0.6))
def test_property_simota_equals_oracle(B, size, lmax, seed, tiny):
    heads = synth.make_heads(B, size, 80, seed)
    labels = synth.make_labels(B, size, lmax, 80, seed + 1)
    rng = np.random.default_rng(seed)
    pick = (labels.sum(2) > 0) & (rng.uniform(0, 1, labels.shape[:2]) < tiny)
    labels[..., 3][pick] = rng.uniform(1, 7, pick.sum()).astype(np.float32)    # tiny GTs: fewer in-both anchors than k
    labels[..., 4][pick] = rng.uniform(1, 7, pick.sum()).astype(np.float32)
    preds, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    hw = [v for s in synth.level_shapes(size) for v in s]
    fg, mg, mi, nfg, ngt = ops.simota_assign_raw(preds, cu(labels), hw, STRIDES)
    o = oracle.simota(preds.cpu().numpy(), labels, synth.level_shapes(size), STRIDES)
    assert np.array_equal(fg.cpu().numpy().astype(np.uint8), o["fg_mask"])
    assert np.array_equal(mg.cpu().numpy(), o["matched_gt"])
    assert np.array_equal(mi.cpu().numpy(), o["matched_iou"])
    assert np.array_equal(nfg.cpu().numpy(), o["num_fg"])


def test_dense_scene_takes_general_path_and_stays_exact():
    """More candidates in one class group than the class-split kernel stages (1536), > 896 kept keys in total, and a
    group that keeps more than max_det: every hand-over to the general kernel is exact."""
    rng = np.random.default_rng(5)
    n = 7000
    boxes = np.stack([rng.uniform(0, 600, n), rng.uniform(0, 600, n)], 1)
    boxes = np.concatenate([boxes, boxes + rng.uniform(2, 6, (n, 2))], 1).astype(np.float32)   # tiny boxes: nearly all kept
    p = np.zeros((1, n, 85), np.float32)
    p[0, :, :4] = boxes
    p[0, :, 4] = rng.uniform(0.3, 1.0, n)
    cls = np.where(rng.uniform(0, 1, n) < 0.5, 0, rng.integers(0, 80, n))      # half of everything in class 0
    p[0, np.arange(n), 5 + cls] = 1.0
    for conf, max_det in [(0.01, 300), (0.8, 300), (0.01, 100)]:
        d, c, k = ops.postprocess_raw(cu(p), conf, 0.65, False, 10000, max_det, 0)
        o = oracle.postprocess(p, conf, 0.65, False, max_det=max_det, flavor=oracle.FLAVOR_CUDA)
        assert np.array_equal(c.cpu().numpy(), o["counts"])
        assert np.array_equal(k.cpu().numpy(), o["keep_idx"]) and np.array_equal(d.cpu().numpy(), o["dets"])
