"""ctypes binding of libplyolo.so (include/plyolo.h).  No CPU fallback: if the library is missing
and cannot be built, importing the ops raises."""
from __future__ import annotations

import ctypes
import os
import threading
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_size_t, c_uint8, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libplyolo.so")

OK, ERR_INVALID, ERR_WORKSPACE, ERR_CUDA, ERR_NO_DEVICE = 0, -1, -2, -3, -4
FLAVOR_CUDA, NMS_RULE_CPU, IOU_NOFMA, THR_F64, FLAVOR_CPU = 0, 1, 2, 4, 7
NMS_YOLOX, NMS_YOLOV3, NMS_YOLOV5 = 0, 3, 5
MAX_LEVELS = 8

_lib = None
_lock = threading.Lock()


class PlyoloError(RuntimeError):
    pass


def _declare(lib):
    vp, ip = c_void_p, POINTER(c_int)
    lib.plyolo_version.restype = c_int
    lib.plyolo_last_error.restype = c_char_p
    lib.plyolo_launch_count.restype = c_ulonglong
    lib.plyolo_peer_alloc.restype = c_int
    lib.plyolo_peer_alloc.argtypes = [c_size_t, POINTER(c_void_p), c_char_p]
    lib.plyolo_peer_open.restype = c_int
    lib.plyolo_peer_open.argtypes = [c_char_p, POINTER(c_void_p)]
    lib.plyolo_peer_close.restype = c_int
    lib.plyolo_peer_close.argtypes = [c_void_p, c_int]
    lib.plyolo_enable_peer_access.restype = c_int
    lib.plyolo_enable_peer_access.argtypes = [c_int]
    lib.plyolo_decode_f32.restype = c_int
    lib.plyolo_decode_f32.argtypes = [POINTER(c_void_p), ip, ip, ip, c_int, c_int, c_int, vp, vp, c_int, vp]
    lib.plyolo_postprocess_workspace_bytes.restype = c_size_t
    lib.plyolo_postprocess_workspace_bytes.argtypes = [c_int, c_int]
    lib.plyolo_postprocess_f32.restype = c_int
    lib.plyolo_postprocess_f32.argtypes = [vp, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_int, c_int,
                                           vp, vp, vp, vp, c_size_t, vp]
    lib.plyolo_postprocess_yolo_f32.restype = c_int
    lib.plyolo_postprocess_yolo_f32.argtypes = [vp, c_int, c_int, c_int, c_double, c_double, c_int, c_int, c_int, c_int,
                                                vp, vp, vp, vp, c_size_t, vp]
    lib.plyolo_decode_postprocess_f32.restype = c_int
    lib.plyolo_decode_postprocess_f32.argtypes = [POINTER(c_void_p), ip, ip, ip, c_int, c_int, c_int, c_double,
                                                  c_double, c_int, c_int, c_int, c_int, vp, vp, vp, vp, c_size_t, vp]
    lib.plyolo_decode_postprocess_bcast_f32.restype = c_int
    lib.plyolo_decode_postprocess_bcast_f32.argtypes = [POINTER(c_void_p), ip, ip, ip, c_int, c_int, c_int, c_double,
                                                        c_double, c_int, c_int, c_int, c_int, vp, vp, vp, c_int,
                                                        POINTER(c_void_p), POINTER(c_void_p), vp, c_size_t, vp]
    lib.plyolo_simota_workspace_bytes.restype = c_size_t
    lib.plyolo_simota_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.plyolo_simota_f32.restype = c_int
    lib.plyolo_simota_f32.argtypes = [vp, vp, c_int, c_int, c_int, c_int, ip, ip, ip, c_int, vp, vp, vp, vp, vp, vp,
                                      c_size_t, vp]
    lib.plyolo_in_boxes_info_f32.restype = c_int
    lib.plyolo_in_boxes_info_f32.argtypes = [vp, vp, vp, vp, c_int, c_int, vp, vp, vp, vp]
    lib.plyolo_dynamic_k_matching_f32.restype = c_int
    lib.plyolo_dynamic_k_matching_f32.argtypes = [vp, vp, c_int, c_int, c_int, vp, vp, vp, vp, vp, vp]
    lib.plyolo_voc_tpfp_f32.restype = c_int
    lib.plyolo_voc_tpfp_f32.argtypes = [vp, vp, c_int, c_int, vp, vp, c_int, c_double, c_int, vp, vp, vp]
    lib.plyolo_format_dets_f32.restype = c_int
    lib.plyolo_format_dets_f32.argtypes = [vp, vp, vp, c_int, c_int, vp, vp]
    lib.plyolo_bboxes_iou_f32.restype = c_int
    lib.plyolo_bboxes_iou_f32.argtypes = [vp, c_int, vp, c_int, c_int, vp, vp]
    lib.plyolo_yolox_loss_workspace_bytes.restype = c_size_t
    lib.plyolo_yolox_loss_workspace_bytes.argtypes = [c_int, c_int]
    lib.plyolo_yolox_loss_f32.restype = c_int
    lib.plyolo_yolox_loss_f32.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, c_int, vp, vp, c_size_t, vp]
    lib.plyolo_yolox_loss_backward_f32.restype = c_int
    lib.plyolo_yolox_loss_backward_f32.argtypes = [vp, vp, vp, vp, vp, c_int, c_int, c_int, vp, POINTER(c_void_p), ip, ip,
                                                   ip, c_int, vp]
    lib.plyolo_yolox_l1_f32.restype = c_int
    lib.plyolo_yolox_l1_f32.argtypes = [vp, vp, vp, vp, c_int, c_int, ip, ip, ip, c_int, vp, vp, c_size_t, vp]
    lib.plyolo_yolox_l1_backward_f32.restype = c_int
    lib.plyolo_yolox_l1_backward_f32.argtypes = [vp, vp, vp, vp, c_int, c_int, c_int, vp, POINTER(c_void_p), ip, ip, ip, c_int, vp]


def lib() -> ctypes.CDLL:
    """Loads the in-tree library; builds it first when it is missing or older than its sources and nvcc is present
    (a stale library with no compiler around — the GPU box always has one — is loaded as it is).  PLYOLO_LIB points
    at another build of the same ABI (kernel variants during tuning)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("PLYOLO_LIB") or SO_PATH
        if path == SO_PATH:
            from . import build as _build
            try:
                if _build.needs_build():
                    _build.build()
            except Exception as e:  # noqa: BLE001
                if not os.path.exists(SO_PATH):
                    raise PlyoloError(
                        "libplyolo.so is missing and could not be built (%s). Run `python -m pl_yolo_b200.build` "
                        "(needs nvcc); there is no CPU or PyTorch fallback for these ops." % (e,)) from e
        handle = ctypes.CDLL(path)
        _declare(handle)
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = lib().plyolo_last_error().decode(errors="replace")
        raise PlyoloError("%s failed (%d): %s" % (what, rc, msg))


def launch_count() -> int:
    return int(lib().plyolo_launch_count())


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def ptr_array(values):
    return (c_void_p * len(values))(*[int(v) for v in values])
