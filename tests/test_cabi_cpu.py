"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/plyolo.h declares, validates arguments before touching a device, and fails loudly (no CPU
fallback) when there is no GPU.  No compute is attempted here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from pl_yolo_b200 import _lib, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "plyolo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(plyolo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), "libplyolo.so does not export %s" % s
    assert L.plyolo_version() == 100


def test_workspace_queries_are_pure_host():
    L = _lib.lib()
    assert L.plyolo_postprocess_workspace_bytes(32, 8400) > 32 * 8400 * 24
    assert L.plyolo_simota_workspace_bytes(32, 8400, 120, 3) > 32 * 8400 * 28
    assert L.plyolo_postprocess_workspace_bytes(0, 8400) == 0


def test_argument_validation_happens_before_any_launch():
    L = _lib.lib()
    hs = _lib.int_array([4]); ws = _lib.int_array([6]); st = _lib.int_array([8])
    ptrs = _lib.ptr_array([4096])
    rc = L.plyolo_decode_f32(ptrs, hs, ws, st, 1, 1, 80, 4096, None, 1, None)
    assert rc == _lib.ERR_INVALID and b"square" in L.plyolo_last_error()
    rc = L.plyolo_decode_f32(ptrs, hs, hs, st, 1, 1, 500, 4096, None, 1, None)
    assert rc == _lib.ERR_INVALID
    rc = L.plyolo_simota_f32(4096, 4096, 1, 17, 80, 4, hs, hs, st, 1, 4096, 4096, 4096, 4096, 4096, 4096, 1 << 30, None)
    assert rc == _lib.ERR_INVALID and b"level shapes" in L.plyolo_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    L = _lib.lib()
    hs = _lib.int_array([4]); st = _lib.int_array([8])
    buf = np.zeros(1 << 16, np.float32)
    ptrs = _lib.ptr_array([buf.ctypes.data])
    rc = L.plyolo_decode_f32(ptrs, hs, hs, st, 1, 1, 80, buf.ctypes.data, None, 1, None)
    assert rc == _lib.ERR_NO_DEVICE and b"no CPU fallback" in L.plyolo_last_error()
    with pytest.raises(_lib.PlyoloError):
        ops.decode_raw([torch.zeros(1, 85, 4, 4)], [8], True)
    with pytest.raises(_lib.PlyoloError):
        ops.postprocess_raw(torch.zeros(1, 64, 85), 0.01, 0.65, False, 10000, 300, 0)
    with pytest.raises(_lib.PlyoloError):
        ops.simota_assign_raw(torch.zeros(1, 64, 85), torch.zeros(1, 4, 5), [8, 8], [8])


def test_torch_library_ops_registered_with_fake_kernels():
    for name in ("decode", "postprocess", "decode_postprocess", "simota_assign", "bboxes_iou"):
        assert hasattr(torch.ops.plyolo, name)
    with torch.device("meta"):
        x = [torch.empty(2, 85, 8, 8), torch.empty(2, 85, 4, 4)]
        p, o = torch.ops.plyolo.decode(x, [8, 16], True)
        assert p.shape == (2, 80, 85) and o.shape == (2, 80, 4)
        d, c, k = torch.ops.plyolo.decode_postprocess(x, [8, 16], 0.01, 0.65, False, 10000, 300, 0)
        assert d.shape == (2, 300, 6) and c.shape == (2,) and k.shape == (2, 300)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pl_yolo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "plyolo_oracle" not in text, f
