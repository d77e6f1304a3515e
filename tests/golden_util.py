"""Loads tests/golden/*.npz (outputs of the real reference, see oracle/gen_golden.py) and
regenerates their seeded inputs, checking the stored sha256 first."""
import json
import os

import numpy as np

from pl_yolo_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def names(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.endswith(".npz") and f.startswith(kind):
            out.append(f[:-4])
    return out


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, {k: z[k] for k in z.files if k != "meta"}


def heads_of(meta):
    kw = {}
    if "objects" in meta:
        kw["objects_per_image"] = meta["objects"]
    seed = meta["seed"] if "seed" in meta else meta["seed_heads"]
    heads = synth.make_heads(meta["B"], meta["size"], meta["C"], seed, mode=meta.get("mode", "clustered"), **kw)
    assert synth.digest(*heads) == meta["sha_heads"], "synthetic head stream differs on this platform"
    return heads


def labels_of(meta, arrays):
    lm = meta["labels"]
    if lm["gen"] == "raw":
        lab = arrays["labels"]
    else:
        lab = synth.make_labels(meta["B"], meta["size"], lm["max_labels"], meta["C"], lm["seed"], lm["min_gt"], lm["max_gt"])
    assert synth.digest(lab) == meta["sha_labels"], "synthetic label stream differs on this platform"
    return lab


def eval_preds_of(meta):
    if meta["gen"] == "raw":
        p = np.load(os.path.join(GOLDEN, "post_edges_input.npy"))
    else:
        p = synth.make_eval_preds(meta["B"], meta["A"], 80, meta["seed"], **meta["kw"])
    assert synth.digest(p) == meta["sha_preds"], "synthetic prediction stream differs on this platform"
    return p


def close(a, b, rtol=1e-5, atol=0.0):
    """|a-b| <= atol + rtol*max(|a|,|b|) elementwise."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) <= atol + rtol * np.maximum(np.abs(a), np.abs(b))
