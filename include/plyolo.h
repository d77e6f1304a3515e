/*
 * plyolo.h — C ABI of the B200-native YOLOX detection hot path (libplyolo.so).
 *
 * The reference (Iywie/pl_YOLO) has no plugin / FFI layer: its boundary for this path is a set
 * of Python call signatures.  Each entry point below replaces the body of one of
 * them; the Python shims in pl_yolo_b200/ keep the reference signatures and call these through
 * ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every pointer is DEVICE memory owned by the caller unless its name starts with `host_`
 *     or the comment says "host array"; nothing is allocated or freed inside;
 *   - every call is asynchronous on `stream` (a cudaStream_t) and never synchronises;
 *   - return value: PLYOLO_OK or a negative PLYOLO_ERR_* code; plyolo_last_error() gives the
 *     text for the calling thread; arguments are validated before anything is launched;
 *   - thread-safe for distinct streams / workspaces; no global mutable state;
 *   - all tensors are contiguous fp32 unless stated; layouts in [] are row-major.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Arithmetic contract: fp32 IEEE, one rounding per reference op (compiled -fmad=false, no
 * fast-math; libdevice expf/logf/log1pf/sqrtf and IEEE division exactly as ATen's CUDA kernels),
 * ties broken by lowest index.  `flavor` selects which build of torchvision's NMS arithmetic is
 * reproduced (PLYOLO_FLAVOR_*).
 */
#ifndef PLYOLO_H_
#define PLYOLO_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLYOLO_VERSION 100 /* 0.1.0 */

#define PLYOLO_OK 0
#define PLYOLO_ERR_INVALID (-1)     /* bad argument (null pointer, bad shape, unsupported size) */
#define PLYOLO_ERR_WORKSPACE (-2)   /* workspace pointer null / too small / misaligned (256 B) */
#define PLYOLO_ERR_CUDA (-3)        /* CUDA runtime error at launch; see plyolo_last_error() */
#define PLYOLO_ERR_NO_DEVICE (-4)   /* no sm_100 device: there is no CPU fallback */

#define PLYOLO_MAX_LEVELS 8
#define PLYOLO_MAX_PEERS 15         /* other ranks a call can store its detections to (16 GPUs per NVLink domain) */
#define PLYOLO_MAX_CLASSES 91       /* ATen's CUDA reduce tree is reproduced bit-exactly up to 91 inputs */

/* NMS arithmetic flavor: OR of independent bits. 0 reproduces the reference on CUDA tensors
 * (torchvision 0.26 nms_kernel.cu behind batched_nms' coordinate trick); 7 reproduces the
 * reference on CPU tensors (nms_kernel.cpp, per-class loop when 4*Nk > 4000). */
#define PLYOLO_FLAVOR_CUDA 0
#define PLYOLO_NMS_RULE_CPU 1 /* tv:ops/boxes.py:80 — per-class NMS on un-offset boxes when 4*Nk > 4000 */
#define PLYOLO_IOU_NOFMA 2    /* union = (Sa + Sb) - inter, Sb rounded (CPU kernel); default fuses Sb: fma(wb,hb,Sa) */
#define PLYOLO_THR_F64 4      /* fp32 IoU compared against the double threshold (CPU kernel) */
#define PLYOLO_FLAVOR_CPU 7

/* which reference call site's candidate filter / score / offset arithmetic plyolo_postprocess_yolo_f32 reproduces */
#define PLYOLO_NMS_YOLOX 0   /* models/evaluators/postprocess.py:7-48 (plyolo_postprocess_f32) */
#define PLYOLO_NMS_YOLOV3 3  /* models/losses/yolov3/yolov3_decoder.py:72-116 */
#define PLYOLO_NMS_YOLOV5 5  /* models/losses/yolov5/yolov5_decoder.py:30-87 */

typedef void *plyolo_stream_t; /* cudaStream_t */

int plyolo_version(void);
const char *plyolo_last_error(void);
/* Peer-visible buffers for plyolo_decode_postprocess_bcast_f32 (CUDA IPC, one process per GPU):
 *   plyolo_peer_alloc  cudaMalloc + zero + export: *ptr on the current device, handle64 = 64 opaque bytes to send to the
 *                      other ranks;  plyolo_peer_open maps another rank's buffer for kernels of the CURRENT device;
 *   plyolo_peer_close  releases a mapping (opened != 0) or frees an allocation (opened == 0). */
int plyolo_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int plyolo_peer_open(const unsigned char *handle64, void **ptr);
int plyolo_peer_close(void *ptr, int opened);
/* lets kernels running on the current device store to memory of `peer_device` (cudaDeviceEnablePeerAccess; needed once
 * per process and peer before plyolo_decode_postprocess_bcast_f32 is given buffers of that device) */
int plyolo_enable_peer_access(int peer_device);
/* number of kernels the calling thread has launched through this library so far (bench accounting) */
unsigned long long plyolo_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * decode — replaces YOLOXLoss.decode (models/losses/yolox/yolox_loss.py:175-228) and, with
 * inference != 0, also the eval branch of YOLOXLoss.__call__ (:25-36), i.e. all of
 * YOLOXDecoder.__call__ (models/losses/yolox/yolox_decoder.py:16-58).
 *   host_lvl   host array of n_levels device pointers, level l = [B, 5+C, hs[l], ws[l]]
 *              (channels reg(4)|obj|cls(C), models/heads/decoupled_head.py:93)
 *   hs, ws, strides   host arrays [n_levels]; hs[l] == ws[l] is required (reference quirk Q2b)
 *   preds      [B, A, 5+C], A = sum hs*ws.  inference == 0: (cx,cy,w,h, raw obj, raw cls);
 *              inference != 0: (x1,y1,x2,y2, sigmoid(obj), sigmoid(cls))
 *   ori_boxes  [B, A, 4] raw regression outputs (yolox_loss.py:214) or NULL
 * The head maps are not modified (the reference's in-place aliasing, quirk Q1, is not reproduced).
 * ------------------------------------------------------------------------------------------- */
int plyolo_decode_f32(const float *const *host_lvl, const int *hs, const int *ws, const int *strides,
                      int n_levels, int B, int C, float *preds, float *ori_boxes, int inference,
                      plyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * postprocess — replaces postprocess / demo_postprocess (models/evaluators/postprocess.py:7-48,
 * :51-92) including torchvision.ops.batched_nms / nms (tv:ops/boxes.py:51-120, torchvision::nms).
 *   preds      [B, A, 5+C] inference-mode predictions (x1,y1,x2,y2,obj,cls..)
 *   conf_thre, nms_thre   the Python floats the reference receives (doubles; compared as the
 *              reference compares them: conf in fp32, IoU per `flavor`)
 *   max_nms    10000 in the reference (:9): the FIRST max_nms candidates in anchor order survive
 *   max_det    300 in the reference (:8)
 *   dets       [B, max_det, 6] rows (x1,y1,x2,y2,conf,class) score-descending, zero padded
 *   counts     [B] int32, rows valid per image (0 <=> the reference's None entry)
 *   keep_idx   [B, max_det] int32 anchor index of every detection (-1 padded) or NULL
 * Workspace: plyolo_postprocess_workspace_bytes(B, A); 256-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t plyolo_postprocess_workspace_bytes(int B, int A);
int plyolo_postprocess_f32(const float *preds, int B, int A, int C, double conf_thre, double nms_thre,
                           int class_agnostic, int max_nms, int max_det, int flavor, float *dets,
                           int32_t *counts, int32_t *keep_idx, void *workspace, size_t workspace_bytes,
                           plyolo_stream_t stream);

/* The NMS call sites of the sibling heads, multi_label == False:
 *   PLYOLO_NMS_YOLOV3 (yolov3_decoder.py:72-116): obj > conf, conf = max_c(cls_c * obj) > conf (strict), the max_nms
 *     best by conf when there are more (:98-100), plain class-agnostic torchvision nms on conf (:103-106: the class
 *     offset is computed but never applied), rows (x1,y1,x2,y2, obj, conf, class);
 *   PLYOLO_NMS_YOLOV5 (yolov5_decoder.py:30-87): obj > conf (strict), obj * max_c cls_c >= conf, the max_nms best by
 *     OBJECTNESS when there are more (:66-67), nms on boxes + class * 4096 (0 if class_agnostic, :70-72) scored by
 *     objectness, rows (x1,y1,x2,y2, obj, best class score, class).
 *   preds  [B, N, 5+C] decoded predictions (cx,cy,w,h, obj, cls..), all squashed, as both decoders build them
 *   dets   [B, max_det, 7] zero padded; counts [B]; keep_idx [B, max_det] row index into preds (-1 padded) or NULL
 * Limits: min(N, 16384) candidates can be sorted per image; an image with more candidates than that reports
 * counts[b] = -1 (nothing else is written for it).  Workspace: plyolo_postprocess_workspace_bytes(B, N). */
int plyolo_postprocess_yolo_f32(const float *preds, int B, int N, int C, double conf_thre, double nms_thre, int variant,
                                int class_agnostic, int max_nms, int max_det, float *dets, int32_t *counts,
                                int32_t *keep_idx, void *workspace, size_t workspace_bytes, plyolo_stream_t stream);

/* Fused decode + postprocess straight from the head maps (reads them once, never materialises
 * preds): what `postprocess(model(imgs, labels), conf, nms)` computes in validation_step
 * (PL_Modules/pl_detection.py:73-76).  Arguments as the two calls above. */
int plyolo_decode_postprocess_f32(const float *const *host_lvl, const int *hs, const int *ws,
                                  const int *strides, int n_levels, int B, int C, double conf_thre,
                                  double nms_thre, int class_agnostic, int max_nms, int max_det,
                                  int flavor, float *dets, int32_t *counts, int32_t *keep_idx,
                                  void *workspace, size_t workspace_bytes, plyolo_stream_t stream);

/* Multi-GPU evaluation: the same call, but the NMS kernels ALSO store every image's rows and count straight into the
 * other ranks' gathered buffers (peer memory mapped over NVLink / NVSwitch, e.g. CUDA IPC) — the detection all-gather of
 * SURVEY 8e happens inside the kernels, no collective is launched.
 *   host_peer_dets / host_peer_counts   host arrays of n_peers device pointers, each pointing at THIS rank's block
 *                                       ([B, max_det, 6] / [B]) inside a peer's gathered buffer; n_peers <= PLYOLO_MAX_PEERS
 * The caller orders consumption across ranks (a barrier after the step's stream work). */
int plyolo_decode_postprocess_bcast_f32(const float *const *host_lvl, const int *hs, const int *ws, const int *strides,
                                        int n_levels, int B, int C, double conf_thre, double nms_thre, int class_agnostic,
                                        int max_nms, int max_det, int flavor, float *dets, int32_t *counts,
                                        int32_t *keep_idx, int n_peers, float *const *host_peer_dets,
                                        int32_t *const *host_peer_counts, void *workspace, size_t workspace_bytes,
                                        plyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * SimOTA assignment — replaces the per-image block of YOLOXLoss.__call__ (yolox_loss.py:43-118):
 * GT prep (:43,:55-65), get_in_boxes_info (:231-315), bboxes_iou(xyxy=False)
 * (models/layers/losses/iou_loss.py:391-414), the BCE/IoU cost (:84-108) and
 * dynamic_k_matching (:318-370), for the whole batch in one call.
 *   preds       [B, A, 5+C] training-mode decode output (cx,cy,w,h, raw obj, raw cls)
 *   labels      [B, Lmax, 5] rows (class,cx,cy,w,h), valid rows first, zero padded
 *   fg_mask     [B, A] uint8   final foreground mask (the reference's mutated fg_mask, :361)
 *   matched_gt  [B, A] int32   matched GT row per anchor, -1 = background
 *               (matched_gt_inds = matched_gt[fg_mask], ascending anchor order, :363)
 *   matched_iou [B, A] fp32    IoU with the matched GT, 0 for background (pred_ious_this_matching, :367)
 *   num_fg      [B] int32      (:358)        num_gt [B] int32 (:43)
 * Workspace: plyolo_simota_workspace_bytes(B, A, Lmax, n_levels); 256-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t plyolo_simota_workspace_bytes(int B, int A, int Lmax, int n_levels);
int plyolo_simota_f32(const float *preds, const float *labels, int B, int A, int C, int Lmax,
                      const int *hs, const int *ws, const int *strides, int n_levels, uint8_t *fg_mask,
                      int32_t *matched_gt, float *matched_iou, int32_t *num_fg, int32_t *num_gt,
                      void *workspace, size_t workspace_bytes, plyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * The two stand-alone pieces of the assignment that the reference exposes as module-level functions.
 * plyolo_simota_f32 fuses both with the cost computation; these serve callers of the functions themselves.
 *
 * get_in_boxes_info (models/losses/yolox/yolox_loss.py:231-315):
 *   gt [G,4] (cx,cy,w,h); expanded_strides, x_shifts, y_shifts [A] (the reference's [1,A] tensors)
 *   fg_mask    [A] uint8     is_in_boxes_or_center (:310)
 *   in_boxes   [G,A] uint8   is_in_boxes (:276)        in_centers [G,A] uint8  is_in_centers (:306)
 *   (the reference's second return value is (in_boxes & in_centers)[:, fg_mask], :312-314)
 *
 * dynamic_k_matching (:318-370) on a caller-provided cost / IoU matrix:
 *   cost, ious [G,Nc];  matching [G,Nc] uint8 scratch (the final matching_matrix);  dynamic_ks [G] int32 (:340)
 *   selected [Nc] uint8 (fg_mask_inboxes, :357), matched_gt [Nc] int32 (-1 if not selected, :363),
 *   matched_iou [Nc] (:367).  Ties in a cost row go to the lowest index (stable-sort rule).
 *   exact_k == 0: YOLOX (:342-345: a GT whose k >= Nc - 1 takes every candidate, quirk Q3);
 *   exact_k != 0: the YOLOv7 matching (models/losses/yolov7/yolov7_loss.py:236-262: torch.topk(cost, k, largest=False),
 *   exactly k candidates per GT; same conflict rule) on the cost / IoU matrix build_targets computed.
 * ------------------------------------------------------------------------------------------- */
int plyolo_in_boxes_info_f32(const float *gt, const float *expanded_strides, const float *x_shifts,
                             const float *y_shifts, int A, int G, uint8_t *fg_mask, uint8_t *in_boxes,
                             uint8_t *in_centers, plyolo_stream_t stream);
int plyolo_dynamic_k_matching_f32(const float *cost, const float *ious, int G, int Nc, int exact_k, uint8_t *matching,
                                  int32_t *dynamic_ks, uint8_t *selected, int32_t *matched_gt, float *matched_iou,
                                  plyolo_stream_t stream);

/* pairwise IoU — replaces bboxes_iou (models/layers/losses/iou_loss.py:391-414).
 *   a [na,4], b [nb,4] -> out [na,nb]; xyxy != 0: corner format, else (cx,cy,w,h). */
int plyolo_bboxes_iou_f32(const float *a, int na, const float *b, int nb, int xyxy, float *out,
                          plyolo_stream_t stream);

/* device part of format_outputs (models/evaluators/postprocess.py:95-138; xyxy2xywh models/utils/bbox.py:58-63):
 * rescales every detection exactly as `bboxes /= scale` does on CUDA tensors and converts to xywh.
 *   dets [B, max_det, 6], counts [B] as returned by plyolo_postprocess_f32; inv_scales [B] fp32 =
 *   (float)(1.0 / scale) with scale = the reference's min(val_w / img_w, val_h / img_h) in double
 *   out  [B, max_det, 8] rows (x1, y1, x2, y2, w, h, score, class), zero padded past counts[b] */
int plyolo_format_dets_f32(const float *dets, const int32_t *counts, const float *inv_scales, int B, int max_det,
                           float *out, plyolo_stream_t stream);

/* VOC evaluator statistics on the device — replaces tpfp_default (models/evaluators/eval_voc.py:75-105,
 * IoU by bbox_overlaps, models/utils/bbox.py:97-139) for every (image, class) of a batch in one launch, instead of the
 * reference's per-class multiprocessing.Pool(8) over numpy arrays (:18-31).
 *   dets [B,max_det,6] rows (x1,y1,x2,y2,score,class) in score-descending order per image (what postprocess /
 *        format_outputs hold), counts [B];  gts [B,Gmax,5] rows (x1,y1,x2,y2,class), gt_counts [B]
 *   tp   [B,max_det] uint8: 1 = true positive; a valid row with 0 is a false positive (fp = 1 - tp, :81-103)
 *   num_gts [C] int32: ground-truth boxes per class over the batch (:33-35) — all-reduce it across ranks; the TP flags
 *        travel with the padded detections in the detection all-gather. */
int plyolo_voc_tpfp_f32(const float *dets, const int32_t *counts, int B, int max_det, const float *gts,
                        const int32_t *gt_counts, int Gmax, double iou_thr, int C, uint8_t *tp, int32_t *num_gts,
                        plyolo_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * loss tail of YOLOXLoss (N2) — replaces models/losses/yolox/yolox_loss.py:121-163 (the use_l1 term: plyolo_yolox_l1_f32 below): the
 * per-image target building (:123-127, :142-147), IOUloss(loss_type="giou")
 * (models/layers/losses/iou_loss.py:7-50) on the foreground anchors, BCEWithLogitsLoss on the objectness of
 * all B*A anchors (:152) and on the classes of the foreground anchors (:154).
 *   preds, labels            as for plyolo_simota_f32
 *   fg_mask, matched_gt, matched_iou   the outputs of plyolo_simota_f32
 *   sums   [3] fp32 device: sum of the GIoU losses, of the objectness BCE terms, of the class BCE terms
 *          (the reference's loss_iou / loss_obj / loss_cls are these divided by max(sum num_fg, 1), :148-154)
 * Floating point: fp32 sums in a fixed (deterministic) order that differs from ATen's — within 1e-5 relative.
 * Workspace: plyolo_yolox_loss_workspace_bytes(B, A); 256-byte aligned.
 * ------------------------------------------------------------------------------------------- */
size_t plyolo_yolox_loss_workspace_bytes(int B, int A);
int plyolo_yolox_loss_f32(const float *preds, const float *labels, const uint8_t *fg_mask,
                          const int32_t *matched_gt, const float *matched_iou, int B, int A, int C, int Lmax,
                          float *sums, void *workspace, size_t workspace_bytes, plyolo_stream_t stream);

/* Backward of the three sums straight into the head maps: d(sums . grad_sums) / d(head map l), chained through
 * the training-mode decode (yolox_loss.py:217-219) — what autograd computes for
 * YOLOXLoss.__call__(inputs, labels)["loss"].backward() between the loss dict and the head outputs.
 *   grad_sums      [3] fp32 device: upstream gradient of the three sums (5/N, 1/N, 1/N for the reference's loss)
 *   host_grad_lvl  host array of n_levels device pointers, level l = [B, 5+C, hs[l], ws[l]], fully overwritten */
int plyolo_yolox_loss_backward_f32(const float *preds, const float *labels, const uint8_t *fg_mask,
                                   const int32_t *matched_gt, const float *matched_iou, int B, int C, int Lmax,
                                   const float *grad_sums, float *const *host_grad_lvl, const int *hs,
                                   const int *ws, const int *strides, int n_levels, plyolo_stream_t stream);

/* use_l1 term of the loss tail (yolox_loss.py:128-133, :158; get_l1_type :373-378): sum over the foreground anchors of
 * |raw regression output - (gx / s - grid_x, gy / s - grid_y, log(gw / s + 1e-8), log(gh / s + 1e-8))|
 * (the reference's loss_l1 is this divided by max(sum num_fg, 1)).
 *   ori    [B, A, 4] fp32 device, 16-byte aligned: the second output of plyolo_decode_f32(inference=0)
 *   sum    [1] fp32 device;  workspace: plyolo_yolox_loss_workspace_bytes(B, A)
 * Backward: grad_sum[0] * sign(ori - target) is ADDED to the four regression planes of the head-map gradients
 * (call it after plyolo_yolox_loss_backward_f32, which overwrites them). */
int plyolo_yolox_l1_f32(const float *ori, const float *labels, const uint8_t *fg_mask, const int32_t *matched_gt, int B,
                        int Lmax, const int *hs, const int *ws, const int *strides, int n_levels, float *sum,
                        void *workspace, size_t workspace_bytes, plyolo_stream_t stream);
int plyolo_yolox_l1_backward_f32(const float *ori, const float *labels, const uint8_t *fg_mask, const int32_t *matched_gt,
                                 int B, int C, int Lmax, const float *grad_sum, float *const *host_grad_lvl, const int *hs,
                                 const int *ws, const int *strides, int n_levels, plyolo_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PLYOLO_H_ */
