/*
 * plyolo_oracle.c — CPU restatement of pl_YOLO's YOLOX detection hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (pl_yolo_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity pin: the reference has no tests or golden vectors for this path, so the
 * oracle is pinned against outputs of the reference itself, generated in the build container by
 * oracle/gen_golden.py (imports /root/reference on CPU) and committed under tests/golden/.
 * tests/test_oracle_vs_golden.py holds the comparison.
 *
 * Each function cites the reference file:line it restates (paths relative to the reference
 * root; "tv:" = torchvision 0.26 python/ops).  All arithmetic is fp32, one rounding per
 * reference op (the reference runs every op as its own ATen kernel, so nothing is ever
 * contracted into an FMA): compile with -ffp-contract=off.
 *
 * "Flavor" selects which build of the third-party arithmetic is restated:
 *   PLYOLO_FLAVOR_CUDA (0)  the reference on CUDA tensors: batched_nms takes the coordinate
 *                           trick, torchvision's CUDA IoU fuses the column box's area into an
 *                           FMA (confirmed on B200: 0/20000 near-threshold pairs differ),
 *                           threshold compared in fp32, sums follow ATen's CUDA reduce tree.
 *   PLYOLO_FLAVOR_CPU  (7)  the reference on CPU tensors: per-class NMS when Nk > 1000
 *                           (tv:ops/boxes.py:80), no FMA, fp32 IoU compared against the double
 *                           threshold.  The three differences are separate bits (see below).
 * Ties are always broken by lowest index (north-star rule; equals a stable sort).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* flavor = OR of three independent bits; 0 = the reference on CUDA, 7 = the reference on CPU */
#define PLYOLO_FLAVOR_CUDA 0
#define PLYOLO_NMS_RULE_CPU 1 /* batched_nms takes the per-class loop when 4*Nk > 4000 (else: > 100000) */
#define PLYOLO_IOU_NOFMA 2    /* union = (Sa + Sb) - inter with Sb rounded (CPU kernel) */
#define PLYOLO_THR_F64 4      /* fp32 IoU compared against the double threshold (CPU kernel) */
#define PLYOLO_FLAVOR_CPU 7

#if defined(__FP_FAST_FMAF) || 1
#define FMAF(a, b, c) fmaf((a), (b), (c))
#endif

/* ------------------------------------------------------------------------------------------
 * ATen CUDA reduce order for sum over a contiguous last dim (aten/src/ATen/native/cuda/Reduce.cuh,
 * sum_functor, vt0 = 4, block.x over the reduced dim).  n inputs per output, n_out outputs.
 * Bit-exact against torch 2.11 on a B200 for every n <= 91 probed (tools/probe_aten_cuda.py;
 * n >= 128 takes ATen's vectorised-input path, which is not restated here).
 * ---------------------------------------------------------------------------------------- */
static int last_pow2(int n) {
    int p = 1;
    while (p * 2 <= n) p *= 2;
    return n < 1 ? 1 : p;
}

static int aten_block_width(int n, long n_out) {
    const int max_threads = 512;
    int dim0_pow2 = n < max_threads ? last_pow2(n) : max_threads;
    long d1 = n_out < max_threads ? (long)last_pow2((int)(n_out < 1 ? 1 : n_out)) : max_threads;
    int bw = dim0_pow2 < 32 ? dim0_pow2 : 32;
    int bh = (int)(d1 < max_threads / bw ? d1 : max_threads / bw);
    int bw2 = dim0_pow2 < max_threads / bh ? dim0_pow2 : max_threads / bh;
    return bw2;
}

float plyolo_oracle_aten_cuda_sum(const float *e, int n, long n_out) {
    float lane[512];
    int bw = aten_block_width(n, n_out);
    for (int t = 0; t < bw; ++t) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        int idx = t, slot = 0;
        while (idx < n) {
            acc[slot] = acc[slot] + e[idx];
            idx += bw;
            slot = (slot + 1) & 3;
        }
        lane[t] = ((acc[0] + acc[1]) + acc[2]) + acc[3];
    }
    /* shared-memory fold (offset = w/2 while > 32 lanes) then the warp shuffle tree; both halve:
     * lane[t] += lane[t + w/2] for w = bw, bw/2, ..., 2 (measured on B200, torch 2.11). */
    for (int w = bw; w > 1; w >>= 1) {
        const int half = w >> 1;
        for (int t = 0; t < half; ++t) lane[t] = lane[t] + lane[t + half];
    }
    return lane[0];
}

static float seq_sum(const float *e, int n) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s = s + e[i];
    return s;
}

/* ------------------------------------------------------------------------------------------
 * decode — YOLOXLoss.decode (models/losses/yolox/yolox_loss.py:175-228) and, with
 * inference != 0, the eval branch (:25-36) == YOLOXDecoder.__call__
 * (models/losses/yolox/yolox_decoder.py:16-58).
 *   lvl[l]  : [B, 5+C, H_l, W_l] contiguous (head output, decoupled_head.py:93)
 *   preds   : [B, A, 5+C], A = sum H_l*W_l, anchor a = off_l + y*W_l + x
 *   ori     : [B, A, 4] raw regression outputs (yolox_loss.py:214), may be NULL
 * ---------------------------------------------------------------------------------------- */
static float sigmoidf_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

int plyolo_oracle_decode(const float *const *lvl, const int *hs, const int *ws, const int *strides,
                         int n_levels, int B, int C, float *preds, float *ori, int inference) {
    const int ch = 5 + C;
    long A = 0;
    for (int l = 0; l < n_levels; ++l) A += (long)hs[l] * ws[l];
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        long off = 0;
        for (int l = 0; l < n_levels; ++l) {
            const int H = hs[l], W = ws[l];
            const long HW = (long)H * W;
            const float s = (float)strides[l];
            const float *src = lvl[l] + (long)b * ch * HW;
            for (long i = 0; i < HW; ++i) {
                float *o = preds + ((long)b * A + off + i) * ch;
                /* grid: yolox_loss.py:198-200 (x = i % W, y = i / W for square maps; Q2b) */
                const float gx = (float)(i % W), gy = (float)(i / W);
                const float px = src[0 * HW + i], py = src[1 * HW + i];
                const float pw = src[2 * HW + i], ph = src[3 * HW + i];
                if (ori) {
                    float *r = ori + ((long)b * A + off + i) * 4;
                    r[0] = px; r[1] = py; r[2] = pw; r[3] = ph;
                }
                float cx = (px + gx) * s; /* :217 */
                float cy = (py + gy) * s;
                float w = expf(pw) * s;   /* :219 */
                float h = expf(ph) * s;
                if (inference) {
                    o[0] = cx - w / 2; /* :31-34 */
                    o[1] = cy - h / 2;
                    o[2] = cx + w / 2;
                    o[3] = cy + h / 2;
                    for (int c = 4; c < ch; ++c) o[c] = sigmoidf_ref(src[c * HW + i]); /* :26-27 */
                } else {
                    o[0] = cx; o[1] = cy; o[2] = w; o[3] = h;
                    for (int c = 4; c < ch; ++c) o[c] = src[c * HW + i];
                }
            }
            off += HW;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * NMS — torchvision::nms (tv csrc/ops/{cpu/nms_kernel.cpp,cuda/nms_kernel.cu}) behind
 * torchvision.ops.batched_nms (tv:ops/boxes.py:51-120), called at
 * models/evaluators/postprocess.py:29-41.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float key; int idx; } key_idx_t;

static int cmp_desc_stable(const void *pa, const void *pb) {
    const key_idx_t *a = (const key_idx_t *)pa, *b = (const key_idx_t *)pb;
    if (a->key > b->key) return -1;
    if (a->key < b->key) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}
static int cmp_asc_stable(const void *pa, const void *pb) {
    const key_idx_t *a = (const key_idx_t *)pa, *b = (const key_idx_t *)pb;
    if (a->key < b->key) return -1;
    if (a->key > b->key) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx);
}

/* a = kept (higher-scored, "row") box, b = later ("column") box. */
static int iou_exceeds(const float *a, const float *b, double thr, int flavor) {
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h;
    float Sa = (a[2] - a[0]) * (a[3] - a[1]);
    float iou;
    if (!(flavor & PLYOLO_IOU_NOFMA)) {
        /* sm_100 SASS of torchvision 0.26 devIoU: Sb is fused into the sum */
        float t = FMAF(b[2] - b[0], b[3] - b[1], Sa);
        iou = inter / (t - inter);
    } else {
        float Sb = (b[2] - b[0]) * (b[3] - b[1]);
        iou = inter / (Sa + Sb - inter);
    }
    return (flavor & PLYOLO_THR_F64) ? ((double)iou > thr) : (iou > (float)thr);
}

/* greedy NMS over boxes [n,4] in the given (score-descending) order; keep_out gets ORIGINAL
 * indices in score order; returns #kept (all of them; caller truncates). */
static int nms_greedy(const float *boxes, const key_idx_t *order, int n, double thr, int flavor,
                      const float *cls_or_null, int *keep_out) {
    uint8_t *sup = (uint8_t *)calloc((size_t)n + 1, 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (sup[i]) continue;
        const int oi = order[i].idx;
        keep_out[nk++] = oi;
        for (int j = i + 1; j < n; ++j) {
            if (sup[j]) continue;
            const int oj = order[j].idx;
            if (cls_or_null && cls_or_null[oi] != cls_or_null[oj]) continue;
            if (iou_exceeds(boxes + 4 * (long)oi, boxes + 4 * (long)oj, thr, flavor)) sup[j] = 1;
        }
    }
    free(sup);
    return nk;
}

/* ------------------------------------------------------------------------------------------
 * postprocess — models/evaluators/postprocess.py:7-48 (demo_postprocess :51-92 is a verbatim
 * twin).  preds [B,A,5+C] = (x1,y1,x2,y2,sig(obj),sig(cls)...).
 *   dets     [B,max_det,6]  (x1,y1,x2,y2,conf,cls) score-descending, zero padded
 *   counts   [B]            number of valid rows (0 <=> the reference's None)
 *   keep_idx [B,max_det]    anchor index of each detection (debug / parity), may be NULL
 *   n_cand   [B]            Nk after the max_nms truncation (debug), may be NULL
 * ---------------------------------------------------------------------------------------- */
int plyolo_oracle_postprocess(const float *preds, int B, int A, int C, double conf_thre,
                              double nms_thre, int class_agnostic, int max_nms, int max_det,
                              int flavor, float *dets, int *counts, int *keep_idx, int *n_cand) {
    const int ch = 5 + C;
    const float thr_f = (float)conf_thre; /* tensor >= python-scalar compares in fp32 (measured) */
    int rc = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *P = preds + (long)b * A * ch;
        float *D = dets + (long)b * max_det * 6;
        memset(D, 0, sizeof(float) * 6 * max_det);
        if (keep_idx) for (int i = 0; i < max_det; ++i) keep_idx[(long)b * max_det + i] = -1;
        counts[b] = 0;
        if (n_cand) n_cand[b] = 0;
        float *box = (float *)malloc(sizeof(float) * 4 * (size_t)(A + 1));
        float *boxn = (float *)malloc(sizeof(float) * 4 * (size_t)(A + 1));
        float *conf = (float *)malloc(sizeof(float) * (size_t)(A + 1));
        float *cls = (float *)malloc(sizeof(float) * (size_t)(A + 1));
        int *anchor = (int *)malloc(sizeof(int) * (size_t)(A + 1));
        key_idx_t *order = (key_idx_t *)malloc(sizeof(key_idx_t) * (size_t)(A + 1));
        int *keep = (int *)malloc(sizeof(int) * (size_t)(A + 1));
        int n = 0;
        for (int a = 0; a < A && n < max_nms; ++a) { /* :24-25 keeps the FIRST max_nms in anchor order */
            const float *r = P + (long)a * ch;
            float best = r[5]; int bi = 0;          /* :18 torch.max -> first max index */
            for (int c = 1; c < C; ++c) if (r[5 + c] > best) { best = r[5 + c]; bi = c; }
            float cf = r[4] * best;                  /* :19 */
            if (cf >= thr_f) {                       /* :20 */
                box[4 * n + 0] = r[0]; box[4 * n + 1] = r[1]; box[4 * n + 2] = r[2]; box[4 * n + 3] = r[3];
                conf[n] = cf; cls[n] = (float)bi; anchor[n] = a; ++n;
            }
        }
        if (n_cand) n_cand[b] = n;
        int nk = 0;
        if (n > 0) {
            for (int i = 0; i < n; ++i) { order[i].key = conf[i]; order[i].idx = i; }
            qsort(order, (size_t)n, sizeof(key_idx_t), cmp_desc_stable); /* nms: sort_stable desc */
            if (class_agnostic) {
                nk = nms_greedy(box, order, n, nms_thre, flavor, NULL, keep); /* :30-34 */
            } else {
                /* tv:ops/boxes.py:80 branch: numel > 4000 (cpu) / 100000 (cuda) -> per-class loop */
                const long numel = 4L * n;
                const int vanilla = numel > ((flavor & PLYOLO_NMS_RULE_CPU) ? 4000 : 100000);
                if (vanilla) {
                    /* tv:ops/boxes.py:107-120: per class nms on the un-offset boxes, then the
                     * kept set re-sorted by score (sort(descending=True), unstable in torch; lowest
                     * index first here). */
                    nk = nms_greedy(box, order, n, nms_thre, flavor, cls, keep);
                } else {
                    /* tv:ops/boxes.py:99-103 coordinate trick: three separate fp32 roundings */
                    float mx = box[0];
                    for (long i = 1; i < 4L * n; ++i) if (box[i] > mx) mx = box[i];
                    const float span = mx + 1.0f;
                    for (int i = 0; i < n; ++i) {
                        const float off = cls[i] * span;
                        for (int k = 0; k < 4; ++k) boxn[4 * i + k] = box[4 * i + k] + off;
                    }
                    nk = nms_greedy(boxn, order, n, nms_thre, flavor, NULL, keep);
                }
            }
        }
        if (nk > max_det) nk = max_det; /* :44-45 */
        for (int i = 0; i < nk; ++i) {
            const int k = keep[i];
            float *d = D + 6 * i;
            d[0] = box[4 * k]; d[1] = box[4 * k + 1]; d[2] = box[4 * k + 2]; d[3] = box[4 * k + 3];
            d[4] = conf[k]; d[5] = cls[k];
            if (keep_idx) keep_idx[(long)b * max_det + i] = anchor[k];
        }
        counts[b] = nk;
        free(box); free(boxn); free(conf); free(cls); free(anchor); free(order); free(keep);
    }
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * SimOTA — the per-image block of YOLOXLoss.__call__ (yolox_loss.py:43-118):
 *   GT prep (:43, :55-65), get_in_boxes_info (:231-315), bboxes_iou(xyxy=False)
 *   (models/layers/losses/iou_loss.py:391-414), cost (:84-108), dynamic_k_matching (:318-370).
 * preds  [B,A,5+C]  decoded cx,cy,w,h + RAW obj/cls logits (decode output, training mode)
 * labels [B,Lmax,5] (cls,cx,cy,w,h), zero padded
 * Dense per-anchor outputs (the reference's compact vectors are their ascending-anchor compaction):
 *   fg_mask [B,A] u8, matched_gt [B,A] i32 (-1 = background), matched_iou [B,A] f32,
 *   num_fg [B], num_gt [B].
 * Optional debug outputs (may be NULL): dyn_k [B,Lmax], n_cand [B].
 * sum_order: 0 = ATen CUDA tree (default), 1 = sequential.
 * ---------------------------------------------------------------------------------------- */
int plyolo_oracle_simota(const float *preds, const float *labels, int B, int A, int C, int Lmax,
                         const int *hs, const int *ws, const int *strides, int n_levels,
                         int sum_order, uint8_t *fg_mask, int32_t *matched_gt, float *matched_iou,
                         int32_t *num_fg, int32_t *num_gt, int32_t *dyn_k, int32_t *n_cand) {
    const int ch = 5 + C;
    /* anchor geometry: x_shifts / y_shifts / expanded_strides (yolox_loss.py:204-208) */
    float *ax = (float *)malloc(sizeof(float) * (size_t)A);
    float *ay = (float *)malloc(sizeof(float) * (size_t)A);
    float *as = (float *)malloc(sizeof(float) * (size_t)A);
    {
        long off = 0;
        for (int l = 0; l < n_levels; ++l) {
            for (long i = 0; i < (long)hs[l] * ws[l]; ++i) {
                ax[off + i] = (float)(i % ws[l]);
                ay[off + i] = (float)(i / ws[l]);
                as[off + i] = (float)strides[l];
            }
            off += (long)hs[l] * ws[l];
        }
        if (off != A) { free(ax); free(ay); free(as); return -1; }
    }
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        const float *P = preds + (long)b * A * ch;
        const float *L = labels + (long)b * Lmax * 5;
        uint8_t *FG = fg_mask + (long)b * A;
        int32_t *MG = matched_gt + (long)b * A;
        float *MI = matched_iou + (long)b * A;
        memset(FG, 0, (size_t)A);
        for (int a = 0; a < A; ++a) { MG[a] = -1; MI[a] = 0.f; }
        if (dyn_k) for (int g = 0; g < Lmax; ++g) dyn_k[(long)b * Lmax + g] = 0;
        if (n_cand) n_cand[b] = 0;
        /* :43 nlabel = (labels.sum(2) > 0).sum(1); GTs are rows [0, G) (:64-65) */
        int G = 0;
        for (int g = 0; g < Lmax; ++g) {
            float s5 = (sum_order == 0) ? plyolo_oracle_aten_cuda_sum(L + 5 * g, 5, (long)B * Lmax)
                                        : seq_sum(L + 5 * g, 5);
            if (s5 > 0.f) ++G;
        }
        num_gt[b] = G;
        num_fg[b] = 0;
        if (G == 0) continue; /* :57-62 */

        /* ---- get_in_boxes_info (:231-315) ---- */
        uint8_t *inbox = (uint8_t *)malloc((size_t)G * A);
        uint8_t *inctr = (uint8_t *)malloc((size_t)G * A);
        int *cand = (int *)malloc(sizeof(int) * (size_t)A);
        int Nc = 0;
        for (int a = 0; a < A; ++a) {
            const float s = as[a];
            const float xc = ax[a] * s + 0.5f * s; /* :240-247 */
            const float yc = ay[a] * s + 0.5f * s;
            int any = 0;
            for (int g = 0; g < G; ++g) {
                const float gx = L[5 * g + 1], gy = L[5 * g + 2], gw = L[5 * g + 3], gh = L[5 * g + 4];
                /* :249-281 edges rounded first, then deltas */
                float bl = xc - (gx - 0.5f * gw), br = (gx + 0.5f * gw) - xc;
                float bt = yc - (gy - 0.5f * gh), bb = (gy + 0.5f * gh) - yc;
                float m = fminf(fminf(bl, bt), fminf(br, bb));
                int ib = m > 0.0f;
                /* :284-307 centre box of radius 2.5*stride */
                const float rad = 2.5f * s;
                float cl = xc - (gx - rad), cr = (gx + rad) - xc;
                float ct = yc - (gy - rad), cb = (gy + rad) - yc;
                float mc = fminf(fminf(cl, ct), fminf(cr, cb));
                int ic = mc > 0.0f;
                inbox[(long)g * A + a] = (uint8_t)ib;
                inctr[(long)g * A + a] = (uint8_t)ic;
                any |= ib | ic;
            }
            if (any) cand[Nc++] = a; /* :310 */
        }
        if (n_cand) n_cand[b] = Nc;

        /* ---- per-candidate class probabilities (:94-98): sqrt(sig(cls)*sig(obj)) ---- */
        float *prob = (float *)malloc(sizeof(float) * (size_t)(Nc + 1) * C);
        float *l0 = (float *)malloc(sizeof(float) * (size_t)(Nc + 1) * C); /* t=1 term: -max(log p,-100) */
        float *l1 = (float *)malloc(sizeof(float) * (size_t)(Nc + 1) * C); /* t=0 term: -max(log1p(-p),-100) */
        for (int n = 0; n < Nc; ++n) {
            const float *r = P + (long)cand[n] * ch;
            const float so = sigmoidf_ref(r[4]);
            for (int c = 0; c < C; ++c) {
                float p = sqrtf(sigmoidf_ref(r[5 + c]) * so);
                prob[(long)n * C + c] = p;
                /* ATen binary_cross_entropy: (t-1)*max(log1p(-p),-100) - t*max(log(p),-100) */
                float lg = fmaxf(logf(p), -100.f), lm = fmaxf(log1pf(-p), -100.f);
                l0[(long)n * C + c] = (1.f - 1.f) * lm - 1.f * lg;
                l1[(long)n * C + c] = (0.f - 1.f) * lm - 0.f * lg;
            }
        }

        /* ---- pairwise IoU (iou_loss.py:400-414, xyxy=False) and cost (:84-108) ---- */
        float *iou = (float *)malloc(sizeof(float) * (size_t)G * (Nc + 1));
        float *cost = (float *)malloc(sizeof(float) * (size_t)G * (Nc + 1));
        float *terms = (float *)malloc(sizeof(float) * (size_t)C);
        for (int g = 0; g < G; ++g) {
            const float gx = L[5 * g + 1], gy = L[5 * g + 2], gw = L[5 * g + 3], gh = L[5 * g + 4];
            const int gc = (int)L[5 * g + 0]; /* .to(int64) truncates */
            const float area_a = gw * gh;
            for (int n = 0; n < Nc; ++n) {
                const int a = cand[n];
                const float *r = P + (long)a * ch;
                const float px = r[0], py = r[1], pw = r[2], ph = r[3];
                float tlx = fmaxf(gx - gw / 2, px - pw / 2), tly = fmaxf(gy - gh / 2, py - ph / 2);
                float brx = fminf(gx + gw / 2, px + pw / 2), bry = fminf(gy + gh / 2, py + ph / 2);
                float area_b = pw * ph;
                float en = (float)(tlx < brx) * (float)(tly < bry);
                float area_i = ((brx - tlx) * (bry - tly)) * en;
                float v = area_i / (area_a + area_b - area_i);
                iou[(long)g * Nc + n] = v;
                float liou = -logf(v + 1e-8f); /* :86 */
                for (int c = 0; c < C; ++c)
                    terms[c] = (c == gc) ? l0[(long)n * C + c] : l1[(long)n * C + c];
                float lcls = (sum_order == 0) ? plyolo_oracle_aten_cuda_sum(terms, C, (long)G * Nc)
                                              : seq_sum(terms, C); /* :99-101 .sum(-1) */
                const int both = inbox[(long)g * A + a] & inctr[(long)g * A + a]; /* :312-314 */
                cost[(long)g * Nc + n] = (lcls + 3.0f * liou) + 100000.0f * (float)(!both); /* :104-108 */
            }
        }

        /* ---- dynamic_k_matching (:318-370) ---- */
        uint8_t *M = (uint8_t *)calloc((size_t)G * (Nc + 1), 1);
        key_idx_t *row = (key_idx_t *)malloc(sizeof(key_idx_t) * (size_t)(Nc + 1));
        const int kc = Nc < 10 ? Nc : 10; /* :336 */
        for (int g = 0; g < G; ++g) {
            for (int n = 0; n < Nc; ++n) { row[n].key = iou[(long)g * Nc + n]; row[n].idx = n; }
            qsort(row, (size_t)Nc, sizeof(key_idx_t), cmp_desc_stable); /* :338 */
            float top[10];
            for (int i = 0; i < kc; ++i) top[i] = row[i].key;
            float ssum = (sum_order == 0) ? plyolo_oracle_aten_cuda_sum(top, kc, G) : seq_sum(top, kc);
            int k = (int)ssum; /* :340 .int() truncates toward zero */
            if (k < 1) k = 1;
            if (dyn_k) dyn_k[(long)b * Lmax + g] = k;
            for (int n = 0; n < Nc; ++n) { row[n].key = cost[(long)g * Nc + n]; row[n].idx = n; }
            qsort(row, (size_t)Nc, sizeof(key_idx_t), cmp_asc_stable); /* :342, ties -> lowest index */
            int take = (k < Nc - 1) ? k : Nc; /* :343-344 quirk Q3 */
            for (int i = 0; i < take; ++i) M[(long)g * Nc + row[i].idx] = 1; /* :348 */
        }
        int nfg = 0;
        for (int n = 0; n < Nc; ++n) {
            int cnt = 0, gsel = -1;
            for (int g = 0; g < G; ++g) if (M[(long)g * Nc + n]) { ++cnt; if (gsel < 0) gsel = g; }
            if (cnt > 1) { /* :352-356 argmin over ALL GT rows, first minimum */
                float best = cost[n]; gsel = 0;
                for (int g = 1; g < G; ++g) if (cost[(long)g * Nc + n] < best) { best = cost[(long)g * Nc + n]; gsel = g; }
            }
            if (cnt > 0) { /* :357-369 */
                const int a = cand[n];
                FG[a] = 1; MG[a] = gsel; MI[a] = iou[(long)gsel * Nc + n];
                ++nfg;
            }
        }
        num_fg[b] = nfg;
        free(inbox); free(inctr); free(cand); free(prob); free(l0); free(l1);
        free(iou); free(cost); free(terms); free(M); free(row);
    }
    free(ax); free(ay); free(as);
    return 0;
}

/* pairwise IoU helper exposed for unit tests — iou_loss.py:391-414 (both layouts). */
int plyolo_oracle_bboxes_iou(const float *a, int na, const float *b, int nb, int xyxy, float *out) {
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            const float *p = a + 4 * i, *q = b + 4 * j;
            float tlx, tly, brx, bry, aa, ab;
            if (xyxy) {
                tlx = fmaxf(p[0], q[0]); tly = fmaxf(p[1], q[1]);
                brx = fminf(p[2], q[2]); bry = fminf(p[3], q[3]);
                aa = (p[2] - p[0]) * (p[3] - p[1]); ab = (q[2] - q[0]) * (q[3] - q[1]);
            } else {
                tlx = fmaxf(p[0] - p[2] / 2, q[0] - q[2] / 2); tly = fmaxf(p[1] - p[3] / 2, q[1] - q[3] / 2);
                brx = fminf(p[0] + p[2] / 2, q[0] + q[2] / 2); bry = fminf(p[1] + p[3] / 2, q[1] + q[3] / 2);
                aa = p[2] * p[3]; ab = q[2] * q[3];
            }
            float en = (float)(tlx < brx) * (float)(tly < bry);
            float ai = ((brx - tlx) * (bry - tly)) * en;
            out[(long)i * nb + j] = ai / (aa + ab - ai);
        }
    return 0;
}
