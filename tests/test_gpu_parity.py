"""Parity tests proper: the CUDA path, called through the C ABI (pl_yolo_b200.ops.*_raw are thin
ctypes calls into libplyolo.so), against
  (1) the golden vectors from the real reference (tests/golden),
  (2) the CPU oracle on the same seeded inputs,
  (3) the torch/torchvision op replay running on the SAME GPU (== the reference on CUDA tensors).
Bars: indices, masks, counts bit-exact; floats that never saw a transcendental bit-exact; decoded
boxes/scores within 1e-5 relative against CPU results and bit-exact against the CUDA replay.
"""
import numpy as np
import pytest
import torch

import oracle
from golden_util import close, eval_preds_of, heads_of, labels_of, load, names
from oracle import torch_ops_replay as R
from pl_yolo_b200 import _lib, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5
STRIDES = [8, 16, 32]


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def hw_flat(size):
    return [v for hw in synth.level_shapes(size) for v in hw]


# ------------------------------------------------------------------------------------------ decode
@pytest.mark.parametrize("name", names("decode"))
def test_decode_vs_golden_and_oracle(name):
    meta, g = load(name)
    heads = heads_of(meta)
    rows = g["rows"]
    hd = [cu(h) for h in heads]
    keep = [h.clone() for h in hd]
    tr, ori = ops.decode_raw(hd, meta["strides"], False)
    ev, _ = ops.decode_raw(hd, meta["strides"], True)
    for a, b in zip(hd, keep):
        assert torch.equal(a, b), "decode must not modify the head maps"
    tr, ori, ev = tr.cpu().numpy(), ori.cpu().numpy(), ev.cpu().numpy()
    # golden (reference on CPU)
    assert np.array_equal(ori[:, rows], g["ori_rows"])
    assert np.array_equal(tr[:, rows, 4:], g["train_rows"][..., 4:])
    assert np.array_equal(tr[..., :2], g["train_boxes"][..., :2])
    assert close(tr[..., 2:4], g["train_boxes"][..., 2:4], RTOL).all()
    scale = np.maximum(np.abs(g["train_boxes"][..., 2:4]), np.abs(g["train_boxes"][..., :2]))
    scale = np.concatenate([scale, scale], -1)
    assert (np.abs(ev[..., :4].astype(np.float64) - g["eval_boxes"]) <= RTOL * scale).all()
    assert close(ev[..., 4], g["eval_obj"], RTOL).all()
    assert close(ev[:, rows, 5:], g["eval_rows"][..., 5:], RTOL).all()
    # oracle
    otr, oori = oracle.decode(heads, meta["strides"], inference=False)
    oev, _ = oracle.decode(heads, meta["strides"], inference=True, want_ori=False)
    assert np.array_equal(ori, oori)
    assert np.array_equal(tr[..., :2], otr[..., :2]) and np.array_equal(tr[..., 4:], otr[..., 4:])
    assert close(tr[..., 2:4], otr[..., 2:4], RTOL).all()
    assert close(ev[..., 4:], oev[..., 4:], RTOL).all()


@pytest.mark.parametrize("B,size,seed", [(2, 160, 5), (3, 320, 6), (4, 640, 7), (1, 96, 8)])
def test_decode_bit_exact_vs_cuda_replay(B, size, seed):
    heads = [cu(h) for h in synth.make_heads(B, size, 80, seed)]
    for inference in (False, True):
        p, o = ops.decode_raw(heads, STRIDES, inference)
        rp, ro = R.decode(heads, STRIDES, inference)
        assert torch.equal(o, ro)
        assert torch.equal(p, rp), "decode differs from the ATen CUDA op chain"


def test_decode_other_class_counts_and_levels():
    rng = np.random.default_rng(0)
    for C, shapes, strides in [(1, [(8, 8)], [8]), (20, [(12, 12), (6, 6)], [8, 16]), (91, [(5, 5), (3, 3), (2, 2), (1, 1)], [4, 8, 16, 32])]:
        heads = [rng.normal(0, 2, (2, 5 + C, h, w)).astype(np.float32) for h, w in shapes]
        hd = [cu(h) for h in heads]
        for inference in (False, True):
            p, o = ops.decode_raw(hd, strides, inference)
            rp, ro = R.decode(hd, strides, inference)
            assert torch.equal(p, rp) and torch.equal(o, ro)


def test_decode_rejects_bad_input():
    with pytest.raises(_lib.PlyoloError):
        ops.decode_raw([torch.zeros(1, 85, 4, 4)], [8], True)  # CPU tensor: no fallback
    with pytest.raises(_lib.PlyoloError):
        ops.decode_raw([torch.zeros(1, 85, 4, 6, device=DEV)], [8], True)  # non-square (reference quirk Q2b)


# ------------------------------------------------------------------------------------- postprocess
def dets_list(dets, counts):
    dets, counts = dets.cpu().numpy(), counts.cpu().numpy()
    return dets, counts


@pytest.mark.parametrize("name", names("post"))
def test_postprocess_vs_golden(name):
    meta, g = load(name)
    p = cu(eval_preds_of(meta))
    # reference-on-CPU arithmetic (per-class branch when Nk > 1000, no FMA, double threshold)
    d, c, k = ops.postprocess_raw(p, meta["conf"], meta["nms"], meta["agnostic"], 10000, 300, _lib.FLAVOR_CPU)
    d, c = dets_list(d, c)
    assert np.array_equal(c, g["counts"])
    assert np.array_equal(d, g["dets"])
    # coordinate-trick branch with the CPU kernel's arithmetic
    d2, c2, _ = ops.postprocess_raw(p, meta["conf"], meta["nms"], meta["agnostic"], 10000, 300,
                                    _lib.IOU_NOFMA | _lib.THR_F64)
    d2, c2 = dets_list(d2, c2)
    assert np.array_equal(c2, g["counts_trick"])
    assert np.array_equal(d2, g["dets_trick"])
    # keep_idx addresses the rows it claims
    k = k.cpu().numpy()
    pn = p.cpu().numpy()
    for b in range(pn.shape[0]):
        assert np.array_equal(pn[b, k[b, : c[b]], :4], d[b, : c[b], :4])
        assert (k[b, c[b]:] == -1).all()


def replay_dets(p, conf, nms, agn):
    outs = R.postprocess(p, conf, nms, agn)
    B = p.shape[0]
    d = np.zeros((B, 300, 6), np.float32)
    c = np.zeros(B, np.int32)
    for i, o in enumerate(outs):
        if o is not None:
            c[i] = o.shape[0]
            d[i, : c[i]] = o.cpu().numpy()
    return d, c


@pytest.mark.parametrize("name", names("post"))
def test_postprocess_bit_exact_vs_cuda_torchvision(name):
    """flavor 0 == torchvision's CUDA batched_nms on this very GPU."""
    meta, g = load(name)
    pn = eval_preds_of(meta)
    p = cu(pn)
    d, c, _ = ops.postprocess_raw(p, meta["conf"], meta["nms"], meta["agnostic"], 10000, 300, _lib.FLAVOR_CUDA)
    d, c = dets_list(d, c)
    rd, rc = replay_dets(p, meta["conf"], meta["nms"], meta["agnostic"])
    assert np.array_equal(c, rc)
    assert np.array_equal(d, rd)
    o = oracle.postprocess(pn, meta["conf"], meta["nms"], meta["agnostic"], flavor=oracle.FLAVOR_CUDA)
    assert np.array_equal(c, o["counts"]) and np.array_equal(d, o["dets"])


@pytest.mark.parametrize("B,size,seed,mode", [(4, 640, 31, "clustered"), (2, 320, 32, "clustered"), (3, 640, 33, "sparse"),
                                               (2, 160, 34, "clustered")])
def test_fused_decode_postprocess(B, size, seed, mode):
    heads = [cu(h) for h in synth.make_heads(B, size, 80, seed, mode=mode)]
    preds, _ = ops.decode_raw(heads, STRIDES, True)
    d1, c1, k1 = ops.postprocess_raw(preds, 0.01, 0.65, False, 10000, 300, 0)
    d2, c2, k2 = ops.decode_postprocess_raw(heads, STRIDES, 0.01, 0.65, False, 10000, 300, 0)
    assert torch.equal(c1, c2) and torch.equal(k1, k2) and torch.equal(d1, d2), "fused path differs from decode->postprocess"
    rd, rc = replay_dets(R.decode(heads, STRIDES, True)[0], 0.01, 0.65, False)
    assert np.array_equal(c2.cpu().numpy(), rc)
    assert np.array_equal(d2.cpu().numpy(), rd), "differs from the reference op chain on CUDA"
    # CPU oracle on its own decode: counts/classes/order equal, floats within tolerance
    hn = [h.cpu().numpy() for h in heads]
    op, _ = oracle.decode(hn, STRIDES, inference=True, want_ori=False)
    o = oracle.postprocess(op, 0.01, 0.65, False, flavor=oracle.FLAVOR_CUDA)
    same = np.array_equal(o["counts"], rc) and np.array_equal(o["keep_idx"], k2.cpu().numpy())
    if same:  # ulp-level libm/libdevice differences can flip a borderline candidate; floats must agree when sets do
        assert close(o["dets"][..., :4], rd[..., :4], RTOL, atol=1e-3).all()
        assert close(o["dets"][..., 4], rd[..., 4], RTOL).all()
        assert np.array_equal(o["dets"][..., 5], rd[..., 5])
    assert (np.abs(o["counts"] - rc) <= 2).all()


@pytest.mark.parametrize("C", [1, 3, 20, 33, 91])
def test_fused_other_class_counts(C):
    """Class counts that are not a multiple of the score kernel's 16-class chunks (VOC has 20): fused route ==
    decode -> postprocess == the reference op chain on CUDA."""
    heads = [cu(h) for h in synth.make_heads(3, 320, C, 80 + C)]
    preds, _ = ops.decode_raw(heads, STRIDES, True)
    d1, c1, k1 = ops.postprocess_raw(preds, 0.01, 0.65, False, 10000, 300, 0)
    d2, c2, k2 = ops.decode_postprocess_raw(heads, STRIDES, 0.01, 0.65, False, 10000, 300, 0)
    assert torch.equal(c1, c2) and torch.equal(k1, k2) and torch.equal(d1, d2), "fused path differs from decode->postprocess"
    rd, rc = replay_dets(R.decode(heads, STRIDES, True)[0], 0.01, 0.65, False)
    assert np.array_equal(c2.cpu().numpy(), rc)
    assert np.array_equal(d2.cpu().numpy(), rd), "differs from the reference op chain on CUDA"
    assert int(c2.sum()) > 0


def test_postprocess_large_max_det():
    """max_det above the cluster merge's shared-memory budget (> 800): every image takes the single-CTA path."""
    rng = np.random.default_rng(13)
    n = 3000
    xy = rng.uniform(0, 3000, (n, 2))
    wh = rng.uniform(4, 60, (n, 2))
    boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
    p = _raw_preds(boxes, rng.uniform(0.05, 1, n).astype(np.float32), rng.integers(0, 80, n), A=n)
    _check_post_vs_oracle(p, max_det=1000, flavors=(0,))
    _check_post_vs_oracle(p, max_det=800, flavors=(0,))


def test_fused_argmax_saturation_and_near_ties():
    """Argmax is taken over the sigmoid VALUES (postprocess.py:18): saturated logits tie at 1.0 and the
    first class wins, logits a few ulp apart may round to the same or even an inverted sigmoid.  The fused kernel
    evaluates only a window of classes around the max logit exactly; it must agree with decode -> postprocess and
    with the reference's op chain on CUDA for adversarial class logits."""
    rng = np.random.default_rng(77)
    B, size = 2, 320
    heads = synth.make_heads(B, size, 80, 35, mode="clustered")
    for h in heads:
        n = h[:, 5:].size
        shape = h[:, 5:].shape
        pick = rng.choice([-40.0, -20.0, -3.0, 0.0, 2.0, 9.0, 13.0, 16.0, 16.5, 16.9, 17.0, 17.3, 17.33, 18.0, 20.0, 40.0, 88.0], n).reshape(shape)
        jitter = rng.choice([0.0, 0.0, 1e-7, -1e-7, 1e-6, 3e-6, -2e-6, 1e-5], n).reshape(shape)
        mask = rng.uniform(0, 1, shape) < 0.5
        h[:, 5:] = np.where(mask, (pick * (1.0 + jitter)).astype(np.float32), h[:, 5:])
        h[:, 4] = rng.normal(0.0, 3.0, h[:, 4].shape)  # many anchors pass the objectness gate
    hd = [cu(h) for h in heads]
    for conf in (0.01, 0.5, 0.999):
        preds, _ = ops.decode_raw(hd, STRIDES, True)
        d1, c1, k1 = ops.postprocess_raw(preds, conf, 0.65, False, 10000, 300, 0)
        d2, c2, k2 = ops.decode_postprocess_raw(hd, STRIDES, conf, 0.65, False, 10000, 300, 0)
        assert torch.equal(c1, c2) and torch.equal(k1, k2) and torch.equal(d1, d2), "fused path differs from decode->postprocess"
        rd, rc = replay_dets(R.decode(hd, STRIDES, True)[0], conf, 0.65, False)
        assert np.array_equal(c2.cpu().numpy(), rc) and np.array_equal(d2.cpu().numpy(), rd)


def test_postprocess_cfg2_properties():
    """BASELINE cfg2 at full size (B=32, 640^2): size-independent properties + CUDA-replay equality."""
    B = 32
    heads = [cu(h) for h in synth.make_heads(B, 640, 80, 0)]
    d, c, k = ops.decode_postprocess_raw(heads, STRIDES, 0.01, 0.65, False, 10000, 300, 0)
    dn, cn, kn = d.cpu().numpy(), c.cpu().numpy(), k.cpu().numpy()
    assert (cn <= 300).all() and (cn > 0).all()
    for b in range(B):
        n = cn[b]
        s = dn[b, :n, 4]
        assert (np.diff(s) <= 0).all(), "detections must be score-descending"
        assert (s >= np.float32(0.01)).all()
        assert (dn[b, n:] == 0).all() and (kn[b, n:] == -1).all()
        assert len(set(kn[b, :n].tolist())) == n
    # idempotence: NMS of its own output keeps everything (no kept box suppresses another kept box)
    A = 8400
    p2 = torch.zeros(B, A, 85, device=DEV)
    p2[:, :300, :4] = d[..., :4]
    p2[:, :300, 4] = d[..., 4]
    cls = d[..., 5].long()
    p2[:, :300, 5:].scatter_(2, cls.unsqueeze(-1), 1.0)
    d3, c3, _ = ops.postprocess_raw(p2, 0.01, 0.65, False, 10000, 300, 0)
    assert torch.equal(c3, c)
    assert torch.equal(d3, d)
    rd, rc = replay_dets(R.decode(heads, STRIDES, True)[0], 0.01, 0.65, False)
    assert np.array_equal(cn, rc) and np.array_equal(dn, rd)


def test_postprocess_truncation_1280():
    """Nk > max_nms: the FIRST 10000 candidates in anchor order survive (postprocess.py:24-25)."""
    p = cu(synth.make_eval_preds(2, 33600, 80, 77, p_obj=0.7, size=1280.0, n_clusters=40))
    d, c, k = ops.postprocess_raw(p, 0.01, 0.65, False, 10000, 300, 0)
    rd, rc = replay_dets(p, 0.01, 0.65, False)
    assert np.array_equal(c.cpu().numpy(), rc) and np.array_equal(d.cpu().numpy(), rd)
    # smaller max_nms through the C ABI only (the reference hard-codes 10000): compare with the oracle
    d, c, k = ops.postprocess_raw(p, 0.01, 0.65, False, 500, 100, 0)
    o = oracle.postprocess(p.cpu().numpy(), 0.01, 0.65, False, max_nms=500, max_det=100, flavor=0)
    assert np.array_equal(c.cpu().numpy(), o["counts"]) and np.array_equal(d.cpu().numpy(), o["dets"])
    assert np.array_equal(k.cpu().numpy(), o["keep_idx"])


def test_postprocess_random_sweep_vs_oracle():
    rng = np.random.default_rng(3)
    for it in range(12):
        B = int(rng.integers(1, 5))
        A = int(rng.choice([64, 84, 525, 1000, 2100, 4000]))
        p = synth.make_eval_preds(B, A, 80, 100 + it, p_obj=float(rng.uniform(0.05, 0.9)), n_clusters=int(rng.integers(1, 20)))
        conf = float(rng.choice([0.001, 0.01, 0.25, 0.5]))
        nms = float(rng.choice([0.3, 0.45, 0.65, 0.9]))
        agn = bool(rng.integers(0, 2))
        for flavor in (0, 7):
            d, c, k = ops.postprocess_raw(cu(p), conf, nms, agn, 10000, 300, flavor)
            o = oracle.postprocess(p, conf, nms, agn, flavor=flavor)
            assert np.array_equal(c.cpu().numpy(), o["counts"]), (it, flavor)
            assert np.array_equal(d.cpu().numpy(), o["dets"]), (it, flavor)
            assert np.array_equal(k.cpu().numpy(), o["keep_idx"]), (it, flavor)


def _raw_preds(boxes, scores, classes, A=None, C=80):
    """eval-mode predictions [1, A, 5+C] holding the given boxes (obj = score, one-hot class = 1)."""
    n = len(boxes)
    A = A or n
    p = np.zeros((1, A, 5 + C), np.float32)
    p[0, :n, :4] = boxes
    p[0, :n, 4] = scores
    p[0, np.arange(n), 5 + np.asarray(classes)] = 1.0
    return p


def _check_post_vs_oracle(p, conf=0.01, nms=0.65, agn=False, flavors=(0, 7), **kw):
    for flavor in flavors:
        d, c, k = ops.postprocess_raw(cu(p), conf, nms, agn, kw.get("max_nms", 10000), kw.get("max_det", 300), flavor)
        o = oracle.postprocess(p, conf, nms, agn, max_nms=kw.get("max_nms", 10000), max_det=kw.get("max_det", 300), flavor=flavor)
        assert np.array_equal(c.cpu().numpy(), o["counts"]), flavor
        assert np.array_equal(k.cpu().numpy(), o["keep_idx"]), flavor
        assert np.array_equal(d.cpu().numpy(), o["dets"]), flavor
        if flavor == 0 and not kw:  # and the real thing: torchvision's CUDA kernels behind the reference's op chain
            rd, rc = replay_dets(cu(p), conf, nms, agn)
            assert np.array_equal(c.cpu().numpy(), rc) and np.array_equal(d.cpu().numpy(), rd)
    return o


def test_postprocess_cross_class_suppression():
    """Coordinate trick (tv:ops/boxes.py:99-103): a box with negative corners of class c+1 lands on top of a
    far-corner box of class c and IS suppressed by it on CUDA tensors.  The class-segmented fast path must
    detect this and hand the image to the exact global sweep."""
    # span = 401: X' = X + 401 = [301, 301, 406, 406] vs Y = [300, 300, 400, 400] -> IoU 0.873
    boxes = np.array([[300, 300, 400, 400], [-100, -100, 5, 5], [10, 10, 60, 60]], np.float32)
    p = _raw_preds(boxes, [0.9, 0.8, 0.7], [0, 1, 5], A=64)
    o = _check_post_vs_oracle(p, flavors=(0,))
    assert o["counts"][0] == 2 and 1 not in o["keep_idx"][0, :2].tolist(), "test construction: X must be suppressed across classes"
    _check_post_vs_oracle(p, flavors=(7,))  # CPU flavor at Nk <= 1000 takes the coordinate trick too
    # chains: the suppressed cross box would otherwise have suppressed a same-class neighbour
    boxes = np.array([[300, 300, 400, 400], [-100, -100, 5, 5], [-98, -98, 6, 6], [-60, -60, 30, 30]], np.float32)
    _check_post_vs_oracle(_raw_preds(boxes, [0.9, 0.8, 0.7, 0.6], [0, 1, 1, 1], A=64))
    # random corner populations, many classes
    rng = np.random.default_rng(11)
    for it in range(8):
        n = int(rng.integers(50, 600))
        tl = rng.uniform(-150, -1, (n // 2, 2)); tl = np.concatenate([tl, tl + rng.uniform(2, 160, (n // 2, 2))], 1)
        M = float(rng.choice([320, 400, 640]))
        br = M - rng.uniform(0, 150, (n - n // 2, 2)); br = np.concatenate([br - rng.uniform(2, 160, (n - n // 2, 2)), br], 1)
        boxes = np.concatenate([tl, br]).astype(np.float32)
        boxes[-1, 2:] = M
        perm = rng.permutation(n)
        p = _raw_preds(boxes[perm], rng.uniform(0.05, 1, n).astype(np.float32), rng.integers(0, int(rng.choice([2, 3, 80])), n), A=max(n, 64))
        _check_post_vs_oracle(p)


def test_postprocess_class_segment_shapes():
    """Class segments of every size (empty, 1, 32, 33, hundreds, > max_det keeps per class), score ties,
    and candidate counts around the fast path's shared-memory capacity."""
    rng = np.random.default_rng(12)
    for n, ncls, spread in [(33, 1, 600), (64, 2, 50), (700, 3, 3000), (2000, 80, 600), (4096, 80, 1200), (4097, 7, 1500),
                            (5000, 80, 900), (1500, 1, 5000), (900, 2, 20000)]:
        xy = rng.uniform(0, spread, (n, 2))
        wh = rng.uniform(4, 80, (n, 2))
        boxes = np.concatenate([xy, xy + wh], 1).astype(np.float32)
        scores = rng.choice(np.linspace(0.05, 1, 40), n).astype(np.float32)  # many exact score ties
        p = _raw_preds(boxes, scores, rng.integers(0, ncls, n), A=n)
        _check_post_vs_oracle(p)
        _check_post_vs_oracle(p, agn=True, flavors=(0,))
        _check_post_vs_oracle(p, max_det=50, max_nms=1000, flavors=(0,))


def test_postprocess_errors():
    p = torch.zeros(1, 64, 85, device=DEV)
    with pytest.raises(_lib.PlyoloError):
        ops.postprocess_raw(p, 0.01, 0.65, False, 10000, 5000, 0)  # max_det too large
    with pytest.raises(_lib.PlyoloError):
        ops.postprocess_raw(p.cpu(), 0.01, 0.65, False, 10000, 300, 0)
    L = _lib.lib()
    rc = L.plyolo_postprocess_f32(p.data_ptr(), 1, 64, 80, 0.01, 0.65, 0, 10000, 300, 0, p.data_ptr(), p.data_ptr(), None,
                                  None, 0, None)
    assert rc == _lib.ERR_WORKSPACE and b"workspace" in L.plyolo_last_error()


# ------------------------------------------------------------------------------------------ SimOTA
def run_simota(preds, labels, size, strides=STRIDES):
    fg, mg, mi, nfg, ngt = ops.simota_assign_raw(cu(preds), cu(labels), hw_flat(size), strides)
    return {"fg_mask": fg.cpu().numpy().astype(np.uint8), "matched_gt": mg.cpu().numpy(), "matched_iou": mi.cpu().numpy(),
            "num_fg": nfg.cpu().numpy(), "num_gt": ngt.cpu().numpy()}


def assert_simota_equal(got, want, what, iou_exact=True):
    assert np.array_equal(got["num_gt"], np.asarray(want["num_gt"])), what
    assert np.array_equal(got["fg_mask"], np.asarray(want["fg_mask"]).astype(np.uint8)), what
    assert np.array_equal(got["num_fg"], np.asarray(want["num_fg"])), what
    assert np.array_equal(got["matched_gt"], np.asarray(want["matched_gt"])), what
    if iou_exact:
        assert np.array_equal(got["matched_iou"], np.asarray(want["matched_iou"])), what


@pytest.mark.parametrize("name", names("simota"))
def test_simota_vs_golden_oracle_replay(name):
    meta, g = load(name)
    heads = heads_of(meta)
    labels = labels_of(meta, g)
    preds = synth.make_train_preds(heads, g["ref_boxes"])
    got = run_simota(preds, labels, meta["size"], meta["strides"])
    assert_simota_equal(got, g, "golden (real reference, CPU)")
    o = oracle.simota(preds, labels, synth.level_shapes(meta["size"]), meta["strides"])
    assert_simota_equal(got, o, "oracle")
    r = R.simota(cu(preds), cu(labels), synth.level_shapes(meta["size"]), meta["strides"], stable=True)
    r = {k: v.cpu().numpy() for k, v in r.items()}
    assert_simota_equal(got, r, "torch-op replay on CUDA")


@pytest.mark.parametrize("B,size,lmax,seed", [(8, 640, 120, 41), (4, 320, 40, 42), (6, 160, 20, 43)])
def test_simota_seeded_vs_oracle_and_replay(B, size, lmax, seed):
    heads = synth.make_heads(B, size, 80, seed)
    labels = synth.make_labels(B, size, lmax, 80, seed + 1000)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    got = run_simota(preds, labels, size)
    o = oracle.simota(preds, labels, synth.level_shapes(size), STRIDES)
    assert_simota_equal(got, o, "oracle")
    r = R.simota(dec, cu(labels), synth.level_shapes(size), STRIDES, stable=True)
    assert_simota_equal(got, {k: v.cpu().numpy() for k, v in r.items()}, "torch-op replay on CUDA")


def test_simota_cfg3_full_size_vs_oracle():
    """BASELINE cfg3: B=32, 640^2, G ~ U{1..120}."""
    heads = synth.make_heads(32, 640, 80, 0)
    labels = synth.make_labels(32, 640, 120, 80, 1)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    got = run_simota(preds, labels, 640)
    o = oracle.simota(preds, labels, synth.level_shapes(640), STRIDES)
    assert_simota_equal(got, o, "oracle cfg3")
    # properties: every fg anchor is matched to a valid GT, counts agree, background carries no match
    fg = got["fg_mask"].astype(bool)
    assert (got["num_fg"] == fg.sum(1)).all()
    assert (got["matched_gt"][~fg] == -1).all() and (got["matched_iou"][~fg] == 0).all()
    for b in range(32):
        assert (got["matched_gt"][b][fg[b]] < got["num_gt"][b]).all() and (got["matched_gt"][b][fg[b]] >= 0).all()
    # permuting the valid label rows permutes matched_gt and nothing else (ties aside: seeded, none here)
    perm_labels = labels.copy()
    inv = []
    rng = np.random.default_rng(9)
    for b in range(32):
        G = got["num_gt"][b]
        pm = rng.permutation(G)
        perm_labels[b, :G] = labels[b, pm]
        inv.append(pm)
    got2 = run_simota(preds, perm_labels, 640)
    assert np.array_equal(got2["fg_mask"], got["fg_mask"])
    for b in range(32):
        m2 = got2["matched_gt"][b][fg[b]]
        assert np.array_equal(inv[b][m2], got["matched_gt"][b][fg[b]])


def test_simota_cfg5_dense_gt_stress_vs_oracle():
    """BASELINE cfg5 shapes at a small batch: 1280^2 (33600 anchors), 250..500 GTs per image, 40 objects per image;
    plus the BASELINE cfg4 label layout (Lmax 120) at batch 8.  Bit-exact against the CPU oracle."""
    heads = synth.make_heads(2, 1280, 80, 51, objects_per_image=40)
    labels = synth.make_labels(2, 1280, 500, 80, 52, min_gt=250, max_gt=500)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    got = run_simota(preds, labels, 1280)
    o = oracle.simota(preds, labels, synth.level_shapes(1280), STRIDES)
    assert_simota_equal(got, o, "oracle cfg5")
    assert (got["num_gt"] >= 250).all() and (got["num_fg"] > 0).all()
    heads = synth.make_heads(8, 640, 80, 53)
    labels = synth.make_labels(8, 640, 120, 80, 54)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    assert_simota_equal(run_simota(preds, labels, 640), oracle.simota(preds, labels, synth.level_shapes(640), STRIDES), "oracle cfg4")


def test_decode_postprocess_cfg5_1280():
    """BASELINE cfg5 decode + NMS: 1280^2, ~16k candidates per image -> max_nms truncation by position, general path."""
    heads = [cu(h) for h in synth.make_heads(2, 1280, 80, 55, objects_per_image=40)]
    d2, c2, k2 = ops.decode_postprocess_raw(heads, STRIDES, 0.01, 0.65, False, 10000, 300, 0)
    preds, _ = ops.decode_raw(heads, STRIDES, True)
    d1, c1, k1 = ops.postprocess_raw(preds, 0.01, 0.65, False, 10000, 300, 0)
    assert torch.equal(c1, c2) and torch.equal(k1, k2) and torch.equal(d1, d2)
    rd, rc = replay_dets(R.decode(heads, STRIDES, True)[0], 0.01, 0.65, False)
    assert np.array_equal(c2.cpu().numpy(), rc) and np.array_equal(d2.cpu().numpy(), rd)


def test_simota_adversarial_edges_vs_oracle():
    """Tiny GTs (k > #in-both -> +1e5 quantised ties), duplicates, Q3 (k >= Nc-1), dense overlaps, G=0."""
    size = 160
    rng = np.random.default_rng(12)
    heads = synth.make_heads(16, size, 80, 55, objects_per_image=6)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    L = np.zeros((16, 24, 5), np.float32)
    for b in range(16):
        G = int(rng.integers(0, 25))
        for g in range(G):
            kind = rng.integers(0, 5)
            if kind == 0:    # tiny box between cell centres
                L[b, g] = [rng.integers(0, 80), rng.uniform(4, 156), rng.uniform(4, 156), rng.uniform(0.5, 3), rng.uniform(0.5, 3)]
            elif kind == 1:  # duplicate of the previous row
                L[b, g] = L[b, g - 1] if g else [3, 50, 50, 20, 20]
            elif kind == 2:  # corner box
                L[b, g] = [rng.integers(0, 80), rng.uniform(0, 6), rng.uniform(0, 6), rng.uniform(2, 10), rng.uniform(2, 10)]
            elif kind == 3:  # huge
                L[b, g] = [rng.integers(0, 80), 80, 80, rng.uniform(100, 200), rng.uniform(100, 200)]
            else:
                L[b, g] = [rng.integers(0, 80), rng.uniform(20, 140), rng.uniform(20, 140), rng.uniform(8, 60), rng.uniform(8, 60)]
        keep = L[b, :, 1:].sum(1) + L[b, :, 0] > 0
        L[b] = np.concatenate([L[b][keep], np.zeros((int((~keep).sum()), 5), np.float32)])
    got = run_simota(preds, L, size)
    o = oracle.simota(preds, L, synth.level_shapes(size), STRIDES)
    assert_simota_equal(got, o, "oracle adversarial")
    r = R.simota(dec, cu(L), synth.level_shapes(size), STRIDES, stable=True)
    assert_simota_equal(got, {k: v.cpu().numpy() for k, v in r.items()}, "replay adversarial")


def test_simota_iou_sweep_routes():
    """The IoU sweep's two routes (per-lane top-4 lists vs the exact warp-wide list) must agree with the oracle:
    forced exact route on seeded inputs, and a construction where the 10 best candidates of a GT sit in the same
    lane of consecutive candidate groups (the lane lists cannot hold them: automatic fallback)."""
    import ctypes
    L = _lib.lib()
    L.plyolo_debug_simota_force_exact.argtypes = [ctypes.c_int]
    size = 320
    heads = synth.make_heads(4, size, 80, 21)
    labels = synth.make_labels(4, size, 40, 80, 22)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    o = oracle.simota(preds, labels, synth.level_shapes(size), STRIDES)
    assert_simota_equal(run_simota(preds, labels, size), o, "lane-list route")
    L.plyolo_debug_simota_force_exact(1)
    try:
        assert_simota_equal(run_simota(preds, labels, size), o, "forced exact route")
    finally:
        L.plyolo_debug_simota_force_exact(0)
    # one GT covering the whole stride-8 level: every anchor is a candidate, candidate n == anchor n on level 0, so
    # anchors 0, 32, 64, ... share a lane.  Their predictions are made near-perfect copies of the GT.
    size = 160
    heads = synth.make_heads(2, size, 80, 23, objects_per_image=0)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy().copy()
    lab = np.zeros((2, 4, 5), np.float32)
    lab[:, 0] = [7, 80, 80, 158, 158]
    lab[1, 1] = [9, 40, 40, 30, 30]
    for j in range(12):
        preds[:, 32 * j + 5, :4] = [80 + 0.01 * j, 80, 158 - 0.3 * j, 158]
    o = oracle.simota(preds, lab, synth.level_shapes(size), STRIDES)
    assert_simota_equal(run_simota(preds, lab, size), o, "same-lane top-10 (fallback)")


def test_simota_more_than_one_sweep_chunk():
    """More than 1024 candidate groups (a GT covering a 1312^2 image: every one of the 35301 anchors is a candidate):
    the IoU sweep lists the groups in several passes."""
    size = 1312
    heads = synth.make_heads(1, size, 80, 71, objects_per_image=20)
    dec, _ = ops.decode_raw([cu(h) for h in heads], STRIDES, False)
    preds = dec.cpu().numpy()
    lab = np.zeros((1, 6, 5), np.float32)
    lab[0, 0] = [3, 656, 656, 1308, 1308]
    lab[0, 1] = [7, 300, 400, 200, 150]
    lab[0, 2] = [9, 900, 1000, 50, 60]
    lab[0, 3] = [11, 1200, 200, 600, 300]
    o = oracle.simota(preds, lab, synth.level_shapes(size), STRIDES)
    assert int(o["n_cand"][0]) > 1024 * 32, "test construction: needs more than 1024 candidate groups"
    assert_simota_equal(run_simota(preds, lab, size), o, "oracle, multi-chunk sweep")


def test_simota_small_class_counts():
    rng = np.random.default_rng(4)
    for C, shapes, strides in [(1, [(8, 8), (4, 4)], [8, 16]), (20, [(16, 16), (8, 8), (4, 4)], [8, 16, 32]), (33, [(16, 16)], [8])]:
        heads = [rng.normal(-1, 2, (3, 5 + C, h, w)).astype(np.float32) for h, w in shapes]
        dec, _ = ops.decode_raw([cu(h) for h in heads], strides, False)
        S = shapes[0][0] * strides[0]
        L = np.zeros((3, 6, 5), np.float32)
        for b in range(3):
            for g in range(int(rng.integers(1, 7))):
                L[b, g] = [rng.integers(0, C), rng.uniform(10, S - 10), rng.uniform(10, S - 10), rng.uniform(6, S / 2), rng.uniform(6, S / 2)]
        hw = [v for s_ in shapes for v in s_]
        fg, mg, mi, nfg, ngt = ops.simota_assign_raw(dec, cu(L), hw, strides)
        o = oracle.simota(dec.cpu().numpy(), L, shapes, strides)
        got = {"fg_mask": fg.cpu().numpy().astype(np.uint8), "matched_gt": mg.cpu().numpy(), "matched_iou": mi.cpu().numpy(),
               "num_fg": nfg.cpu().numpy(), "num_gt": ngt.cpu().numpy()}
        assert_simota_equal(got, o, "C=%d" % C)
        r = R.simota(dec, cu(L), shapes, strides, stable=True)
        assert_simota_equal(got, {k: v.cpu().numpy() for k, v in r.items()}, "replay C=%d" % C)


def test_loss_tail_cabi_vs_oracle():
    """N2 through the raw ops (C ABI): the three sums and the head-map gradients against oracle/loss_oracle.py
    (float64 numpy, pinned to the real reference) at 640^2 with up to 120 GTs; fp32 kernel: 1e-5 relative on the sums,
    1e-5 of the level's largest gradient on the gradients."""
    from oracle import loss_oracle
    B, size = 6, 640
    heads = synth.make_heads(B, size, 80, 61)
    labels = synth.make_labels(B, size, 120, 80, 62)
    hd = [cu(h) for h in heads]
    preds, _ = ops.decode_raw(hd, STRIDES, False)
    fg, mg, mi, nfg, ngt = ops.simota_assign_raw(preds, cu(labels), hw_flat(size), STRIDES)
    sums = ops.yolox_loss_sums_raw(preds, cu(labels), fg, mg, mi).cpu().numpy().astype(np.float64)
    a = {"fg_mask": fg.cpu().numpy().astype(np.uint8), "matched_gt": mg.cpu().numpy(), "matched_iou": mi.cpu().numpy(),
         "num_fg": nfg.cpu().numpy(), "num_gt": ngt.cpu().numpy()}
    losses, grads = loss_oracle.loss_tail(preds.cpu().numpy(), labels, a, synth.level_shapes(size), STRIDES)
    N = max(int(a["num_fg"].sum()), 1)
    assert sums[0] / N == pytest.approx(losses["loss_iou"], rel=1e-5)
    assert sums[1] / N == pytest.approx(losses["loss_obj"], rel=1e-5)
    assert sums[2] / N == pytest.approx(losses["loss_cls"], rel=1e-5)
    gs = torch.tensor([5.0 / N, 1.0 / N, 1.0 / N], device=DEV)
    got = ops.yolox_loss_backward_raw(preds, cu(labels), fg, mg, mi, gs, hw_flat(size), STRIDES)
    for l, (x, y) in enumerate(zip(got, grads)):
        x = x.cpu().numpy().astype(np.float64)
        assert x.shape == y.shape
        assert np.abs(x - y).max() <= 1e-5 * np.abs(y).max(), "level %d" % l


def test_bboxes_iou_vs_replay():
    rng = np.random.default_rng(1)
    a = rng.uniform(0, 100, (17, 4)).astype(np.float32)
    b = rng.uniform(0, 100, (33, 4)).astype(np.float32)
    out = ops.bboxes_iou_raw(cu(a), cu(b), False)
    assert torch.equal(out, R.pairwise_iou_cxcywh(cu(a), cu(b)))
    assert np.array_equal(out.cpu().numpy(), oracle.bboxes_iou(a, b, xyxy=False))
    a[:, 2:] += a[:, :2]
    b[:, 2:] += b[:, :2]
    assert np.array_equal(ops.bboxes_iou_raw(cu(a), cu(b), True).cpu().numpy(), oracle.bboxes_iou(a, b, xyxy=True))


def test_libdevice_matches_aten_probe():
    """The transcendental building blocks equal ATen's CUDA results (via decode: exp and sigmoid)."""
    x = torch.randn(1, 85, 64, 64, device=DEV) * 6
    p, _ = ops.decode_raw([x], [8], True)
    flat = x.flatten(2).permute(0, 2, 1)
    assert torch.equal(p[..., 4:], torch.sigmoid(flat[..., 4:]))
    p2, _ = ops.decode_raw([x], [8], False)
    assert torch.equal(p2[..., 2:4], torch.exp(flat[..., 2:4]) * 8)
