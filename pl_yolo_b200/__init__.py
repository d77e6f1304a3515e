"""pl_yolo_b200 — B200-native YOLOX detection hot path for Iywie/pl_YOLO.

Host-side mirror of the reference interface for this path (same names, arguments, errors):
    YOLOXLoss, YOLOXDecoder, postprocess / demo_postprocess, bboxes_iou
over the C ABI of libplyolo.so (include/plyolo.h): hand-written sm_100a kernels, no CPU fallback.
"""
from .iou_loss import bboxes_iou  # noqa: F401
from .pipeline import Lanes, PostprocessPipeline  # noqa: F401
from .postprocess import LazyPredictions, demo_postprocess, format_outputs, postprocess, postprocess_dense  # noqa: F401
from .yolox_decoder import YOLOXDecoder  # noqa: F401
from .yolox_loss import YOLOXLoss, dynamic_k_matching, get_in_boxes_info  # noqa: F401

__all__ = ["YOLOXLoss", "YOLOXDecoder", "postprocess", "demo_postprocess", "format_outputs", "postprocess_dense", "bboxes_iou",
           "LazyPredictions", "get_in_boxes_info", "dynamic_k_matching", "PostprocessPipeline", "Lanes"]
