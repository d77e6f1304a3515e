// simota.cu — SimOTA label assignment for a whole batch (yolox_loss.py:43-118, :231-370;
// iou_loss.py:391-414).  Three launches, no host synchronisation, no [G,Nc,80] temporaries.
//
//  K1 simota_prep_kernel      one CTA per image
//     GT count (:43), closed-form geometry prior: for every (GT, level) the in-box and in-centre
//     anchors are axis-aligned cell rectangles, found with the reference's own fp32 comparisons
//     (edge rounded first, then the delta, :249-307) and rasterised into shared-memory bitmaps:
//     fg = union of all rectangles (:310), U = union of the (in-box AND in-centre) rectangles.
//     Both bitmaps are compacted in anchor order; the candidates' corner boxes and areas are gathered
//     (16 of the 340 bytes of each prediction row) for the IoU sweep.
//  K2 simota_cost_table_kernel  one warp per anchor of U
//     the 80-class BCE cost of the anchor for EVERY possible GT class from one reduction in ATen's
//     CUDA order (measured on B200: lane t adds classes t, t+32, t+64, then a halving tree): a
//     butterfly over the all-negative leaves leaves each lane the sibling sums of its path, and
//     swapping in the positive leaf re-adds them.  O(|U|*C) transcendentals instead of O(G*Nc*C).
//  K3 simota_match_kernel     one CTA (4 warps) per GT
//     IoU against every candidate with warp-resident top-10 lists (values only, :336-340; the
//     division runs only when a pair can enter the list) -> dynamic k with ATen's reduce tree ->
//     cost only for the <= 25*levels anchors that are both in-box and in-centre (every other cost
//     carries +1e5, :104-108, so the k smallest live there unless the GT is tiny) -> k smallest
//     (cost, anchor) -> per-anchor match count / lowest-GT atomics.
//     The image's last GT CTA to finish (atomic counter) then finalises the image:
//     anchors matched once take that GT; anchors matched more than once take the argmin of the cost
//     over ALL GTs (:352-356, quirk Q4); fg_mask / matched GT / matched IoU are written densely per
//     anchor and num_fg is counted (:357-369).
#include <cfloat>

#include "common.cuh"

namespace plyolo {

constexpr int kPrepThreads = 512;
constexpr int kMatchWarps = 4;   // warps per GT
constexpr int kTableCtas = 48;   // per image
constexpr int kTableWarps = 8;
constexpr int kMaxBoth = 36 * PLYOLO_MAX_LEVELS;  // 5x5 centre cells per level (6x6 if an edge rounds outward)

struct SimParams {
    const float *preds;
    const float *labels;
    int B, A, C, ch, Lmax, n_levels;
    int hw[PLYOLO_MAX_LEVELS], w[PLYOLO_MAX_LEVELS], off[PLYOLO_MAX_LEVELS];
    float stride[PLYOLO_MAX_LEVELS];
    uint8_t *fg_mask;
    int32_t *matched_gt;
    float *matched_iou;
    int32_t *num_fg;
    int32_t *num_gt;
    // workspace
    int *meta;            // [B,8]  G, Nc, |U|, #matched anchors, #finished GT CTAs, #conflict anchors
    int *conf_list;       // [B,A]  anchors claimed by more than one GT (unordered)
    int *cand_anchor;     // [B,A]  (reused as the conflict list by the finalize kernel)
    float4 *cand_box;     // [B,A]  corners (cx-w/2, cy-h/2, cx+w/2, cy+h/2) of candidate n
    float *cand_area;     // [B,A]  w*h of candidate n
    int *u_anchor;        // [B,A]  anchors of U = union_g (in-box AND in-centre), ascending
    int *u_index;         // [B,A]  position in u_anchor or -1
    float *table;         // [B,A,C] class cost per (anchor of U, GT class)
    unsigned *sel_count;  // [B,A]  number of GTs that claimed the anchor
    int *res_g;           // [B,A]  conflict resolution: argmin GT ...
    float *res_iou;       // [B,A]  ... and its IoU
    short *rect;          // [B,Lmax,n_levels,8] in-box x0,x1,y0,y1 | in-centre x0,x1,y0,y1 (inclusive)
};

// ---- arithmetic shared by the three kernels ------------------------------------------------

// labels.sum(2) over the 5 columns in ATen's CUDA order (4 lanes: (e0+e4), e1, e2, e3; halving tree)
__device__ __forceinline__ float row_sum5(const float *r) { return ((r[0] + r[4]) + r[2]) + (r[1] + r[3]); }

// bboxes_iou(gt, pred, xyxy=False): iou_loss.py:400-414
__device__ __forceinline__ float pair_iou(const float gx, const float gy, const float gw, const float gh,
                                          const float4 pb) {
    const float tlx = fmaxf(gx - gw / 2, pb.x - pb.z / 2), tly = fmaxf(gy - gh / 2, pb.y - pb.w / 2);
    const float brx = fminf(gx + gw / 2, pb.x + pb.z / 2), bry = fminf(gy + gh / 2, pb.y + pb.w / 2);
    const float area_a = gw * gh, area_b = pb.z * pb.w;
    const float en = (tlx < brx ? 1.f : 0.f) * (tly < bry ? 1.f : 0.f);
    const float area_i = ((brx - tlx) * (bry - tly)) * en;
    return area_i / ((area_a + area_b) - area_i);
}

// cell-centre coordinate exactly as get_in_boxes_info forms it (:240-247)
__device__ __forceinline__ float cell_center(const int i, const float s) { return (float)i * s + 0.5f * s; }

// First / last cell index whose centre c satisfies (c - lo) > 0 and (hi - c) > 0 (reference compares, :272-281).
__device__ __forceinline__ void cell_range(const float lo, const float hi, const float s, const int W, int &first,
                                           int &last) {
    int e = (int)fminf(fmaxf(floorf(lo / s - 0.5f) + 1.f, 0.f), (float)W);
    if (!(e >= 0 && e <= W)) e = 0;  // NaN edges
    while (e > 0 && (cell_center(e - 1, s) - lo) > 0.0f) --e;
    while (e < W && !((cell_center(e, s) - lo) > 0.0f)) ++e;
    first = e;
    int f = (int)fminf(fmaxf(ceilf(hi / s - 0.5f) - 1.f, -1.f), (float)(W - 1));
    if (!(f >= -1 && f <= W - 1)) f = W - 1;
    while (f < W - 1 && (hi - cell_center(f + 1, s)) > 0.0f) ++f;
    while (f >= 0 && !((hi - cell_center(f, s)) > 0.0f)) --f;
    last = f;
}

struct LaneTerms {
    float neg[3];  // -max(log1p(-p), -100) for class lane + 32 j   (target 0)
    float p[3];    // p = sqrt(sigmoid(cls) * sigmoid(obj))
};

// BCE leaves of one prediction row, lanes over classes (yolox_loss.py:94-101).
__device__ __forceinline__ void lane_terms(const float *row, const int C, const int lane, LaneTerms &t) {
    const float so = sigmoid_ref(__ldg(row + 4));
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
            const float p = sqrtf(sigmoid_ref(__ldg(row + 5 + c)) * so);
            t.p[j] = p;
            t.neg[j] = -fmaxf(log1pf(-p), -100.f);  // ATen BCE, target 0: (0-1)*max(log1p(-p),-100)
        } else {
            t.p[j] = 0.f;
            t.neg[j] = 0.f;
        }
    }
}
__device__ __forceinline__ float pos_term(const float p) { return -fmaxf(logf(p), -100.f); }  // target 1

// ATen's strided accumulators for one lane: wide == false: 32 lanes ((e0+e1)+e2); wide == true: the
// 64-lane block used when there are fewer than 16 outputs, folded to 32 lanes ((e0+e2)+(e1+0)).
__device__ __forceinline__ float lane_combine(const float e0, const float e1, const float e2, const bool wide) {
    return wide ? ((e0 + e2) + e1) : ((e0 + e1) + e2);
}

// Sum of the C BCE leaves for GT class `gc` in ATen's CUDA reduce order; result valid in lane 0.
// C >= 32 path (registers only); smaller C goes through `terms` in shared memory.
__device__ __forceinline__ float cls_cost(const LaneTerms &t, const int C, const int gc, const int lane,
                                          const bool wide, float *terms /* per-warp [96] */) {
    float e[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = (lane + 32 * j == gc) ? pos_term(t.p[j]) : t.neg[j];
    float v;
    int lanes;
    if (C >= 32) {
        v = lane_combine(e[0], e[1], e[2], wide);
        lanes = 32;
    } else {
        // fewer than 32 classes: bw = last_pow2(C) lanes, lane t sums classes t, t+bw (at most two chunks)
        __syncwarp();
        if (lane < C) terms[lane] = e[0];
        __syncwarp();
        int bw = 1;
        while (bw * 2 <= C) bw <<= 1;
        v = 0.f;
        if (lane < bw) v = (terms[lane] + ((lane + bw < C) ? terms[lane + bw] : 0.f));
        lanes = bw;
    }
    for (int h = lanes >> 1; h >= 1; h >>= 1) {
        const float o = __shfl_down_sync(0xffffffffu, v, h);
        if (lane < h) v = v + o;
    }
    return v;
}

// in_boxes_and_center for (GT rects of the anchor's level, anchor cell)
__device__ __forceinline__ bool in_both(const short *r8, const int x, const int y) {
    return x >= r8[0] && x <= r8[1] && y >= r8[2] && y <= r8[3] && x >= r8[4] && x <= r8[5] && y >= r8[6] && y <= r8[7];
}

__device__ __forceinline__ void anchor_cell(const SimParams &p, const int a, int &l, int &x, int &y) {
    l = 0;
    for (int i = 1; i < p.n_levels; ++i)
        if (a >= p.off[i]) l = i;
    const int r = a - p.off[l];
    y = r / p.w[l];
    x = r - y * p.w[l];
}

__device__ __forceinline__ float4 load_box(const float *row) {
    return make_float4(__ldg(row), __ldg(row + 1), __ldg(row + 2), __ldg(row + 3));
}

// ---- K1 ------------------------------------------------------------------------------------
// sets every bit of [p0, p1] in a shared-memory bitmap
__device__ __forceinline__ void set_bits(unsigned *bitmap, const int p0, const int p1) {
    for (int wd = p0 >> 5; wd <= (p1 >> 5); ++wd) {
        const int lo = max(p0, wd << 5) & 31, hi = min(p1, (wd << 5) + 31) & 31;
        atomicOr(&bitmap[wd], (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
    }
}

// Block-wide, order-preserving compaction of the set bits of `bitmap` into `list` (+ optional inverse
// map).  Returns the number of set bits.  All threads of the CTA must call it.
__device__ __forceinline__ int compact_bitmap(const unsigned *bitmap, const int nwords, const int A, int *list,
                                              int *inverse, int *s_warp) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wpt = (nwords + kPrepThreads - 1) / kPrepThreads;
    const int w0 = min(tid * wpt, nwords), w1 = min(w0 + wpt, nwords);
    int local = 0;
    for (int i = w0; i < w1; ++i) {
        unsigned m = bitmap[i];
        if (i == nwords - 1 && (A & 31)) m &= (1u << (A & 31)) - 1u;
        local += __popc(m);
    }
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    __syncthreads();  // s_warp may still be read by a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = inc - local, total = 0;
    for (int w = 0; w < kPrepThreads / 32; ++w) {
        if (w < warp) base += s_warp[w];
        total += s_warp[w];
    }
    for (int i = w0; i < w1; ++i) {
        unsigned m = bitmap[i];
        if (i == nwords - 1 && (A & 31)) m &= (1u << (A & 31)) - 1u;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int a = (i << 5) + bit;
            list[base] = a;
            if (inverse) inverse[a] = base;
            ++base;
        }
    }
    return total;
}

__global__ void __launch_bounds__(kPrepThreads) simota_prep_kernel(const SimParams p) {
    extern __shared__ unsigned bitmap[];  // [2][ceil(A/32)]: fg, U
    __shared__ int s_G, s_warp[kPrepThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int nwords = (p.A + 31) >> 5;
    unsigned *bm_fg = bitmap, *bm_u = bitmap + nwords;
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    if (tid == 0) s_G = 0;
    for (int i = tid; i < 2 * nwords; i += kPrepThreads) bitmap[i] = 0u;
    for (int a = tid; a < p.A; a += kPrepThreads) {
        p.sel_count[(size_t)b * p.A + a] = 0u;
        p.u_index[(size_t)b * p.A + a] = -1;
        // background defaults (:57-62); the last GT CTA of the image overwrites the matched anchors
        p.fg_mask[(size_t)b * p.A + a] = 0;
        p.matched_gt[(size_t)b * p.A + a] = -1;
        p.matched_iou[(size_t)b * p.A + a] = 0.f;
    }
    __syncthreads();
    // :43 nlabel = (labels.sum(2) > 0).sum(1)
    int cnt = 0;
    for (int g = tid; g < p.Lmax; g += kPrepThreads) cnt += row_sum5(L + 5 * g) > 0.f ? 1 : 0;
    if (cnt) atomicAdd(&s_G, cnt);
    __syncthreads();
    const int G = s_G;  // the GTs are rows [0, G) (:64-65)

    // geometry prior: one work item per (GT, level)
    const int items = G * p.n_levels;
    for (int it = tid; it < items; it += kPrepThreads) {
        const int l = it % p.n_levels, g = it / p.n_levels;
        const float gx = L[5 * g + 1], gy = L[5 * g + 2], gw = L[5 * g + 3], gh = L[5 * g + 4];
        const float s = p.stride[l];
        const int W = p.w[l], H = p.hw[l] / p.w[l];
        int bx0, bx1, by0, by1, cx0, cx1, cy0, cy1;
        cell_range(gx - 0.5f * gw, gx + 0.5f * gw, s, W, bx0, bx1);  // :249-268
        cell_range(gy - 0.5f * gh, gy + 0.5f * gh, s, H, by0, by1);
        const float r = 2.5f * s;                                     // :284-298, center_radius = 2.5
        cell_range(gx - r, gx + r, s, W, cx0, cx1);
        cell_range(gy - r, gy + r, s, H, cy0, cy1);
        short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
        r8[0] = (short)bx0; r8[1] = (short)bx1; r8[2] = (short)by0; r8[3] = (short)by1;
        r8[4] = (short)cx0; r8[5] = (short)cx1; r8[6] = (short)cy0; r8[7] = (short)cy1;
        const int o = p.off[l];
        if (bx0 <= bx1)
            for (int y = by0; y <= by1; ++y) set_bits(bm_fg, o + y * W + bx0, o + y * W + bx1);
        if (cx0 <= cx1)
            for (int y = cy0; y <= cy1; ++y) set_bits(bm_fg, o + y * W + cx0, o + y * W + cx1);
        const int ux0 = max(bx0, cx0), ux1 = min(bx1, cx1), uy0 = max(by0, cy0), uy1 = min(by1, cy1);
        if (ux0 <= ux1)
            for (int y = uy0; y <= uy1; ++y) set_bits(bm_u, o + y * W + ux0, o + y * W + ux1);
    }
    __syncthreads();

    // candidates in anchor order: candidate n <-> n-th set bit of fg (:79-82); same for U
    int *ca = p.cand_anchor + (size_t)b * p.A;
    const int Nc = compact_bitmap(bm_fg, nwords, p.A, ca, nullptr, s_warp);
    const int U = compact_bitmap(bm_u, nwords, p.A, p.u_anchor + (size_t)b * p.A, p.u_index + (size_t)b * p.A, s_warp);
    __syncthreads();  // the candidate list (global) is complete for this CTA
    float4 *cb = p.cand_box + (size_t)b * p.A;
    float *car = p.cand_area + (size_t)b * p.A;
    for (int n0 = tid; n0 < Nc; n0 += 4 * kPrepThreads) {
        float4 pb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // 4 independent row gathers in flight per thread
            const int n = n0 + u * kPrepThreads;
            if (n < Nc) pb[u] = load_box(p.preds + ((size_t)b * p.A + ca[n]) * p.ch);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + u * kPrepThreads;
            if (n < Nc) {
                // the candidate-side operands of bboxes_iou(xyxy=False) (iou_loss.py:400-410)
                cb[n] = make_float4(pb[u].x - pb[u].z / 2, pb[u].y - pb[u].w / 2, pb[u].x + pb[u].z / 2, pb[u].y + pb[u].w / 2);
                car[n] = pb[u].z * pb[u].w;
            }
        }
    }
    if (tid == 0) {
        p.meta[b * 8 + 0] = G;
        p.meta[b * 8 + 1] = Nc;
        p.meta[b * 8 + 2] = U;
        p.meta[b * 8 + 3] = 0;
        p.meta[b * 8 + 4] = 0;
        p.meta[b * 8 + 5] = 0;
        p.num_gt[b] = G;
        p.num_fg[b] = 0;
    }
}

// ---- K2 ------------------------------------------------------------------------------------
// Class-cost table: T[u][c] = sum over classes of BCE(p, onehot(c)) for anchor u_anchor[u], in ATen's
// CUDA reduce order (only used when C >= 32; smaller class counts take the pair_cost path).
__global__ void __launch_bounds__(kTableWarps * 32) simota_cost_table_kernel(const SimParams p) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int G = p.meta[b * 8 + 0], Nc = p.meta[b * 8 + 1], U = p.meta[b * 8 + 2];
    const bool wide = (long long)G * Nc < 16 && p.C >= 64;
    const int *ua = p.u_anchor + (size_t)b * p.A;
    for (int u = blockIdx.x * kTableWarps + warp; u < U; u += kTableCtas * kTableWarps) {
        const float *row = p.preds + ((size_t)b * p.A + ua[u]) * p.ch;
        LaneTerms t;
        lane_terms(row, p.C, lane, t);
        float s = lane_combine(t.neg[0], t.neg[1], t.neg[2], wide);
        float sib[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            sib[i] = __shfl_xor_sync(0xffffffffu, s, 16 >> i);
            s = s + sib[i];
        }
        float *T = p.table + ((size_t)b * p.A + u) * p.C;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = lane + 32 * j;
            if (c < p.C) {
                const float e0 = j == 0 ? pos_term(t.p[0]) : t.neg[0];
                const float e1 = j == 1 ? pos_term(t.p[1]) : t.neg[1];
                const float e2 = j == 2 ? pos_term(t.p[2]) : t.neg[2];
                float r = lane_combine(e0, e1, e2, wide);
#pragma unroll
                for (int i = 0; i < 5; ++i) r = r + sib[i];
                T[c] = r;
            }
        }
    }
}

// cost of (GT g, anchor a) exactly as yolox_loss.py:84-108; warp-cooperative, valid in lane 0
__device__ __forceinline__ float pair_cost(const SimParams &p, const int b, const int a, const float gx,
                                           const float gy, const float gw, const float gh, const int gc,
                                           const bool both, const bool wide, const int lane, float *terms) {
    const float *row = p.preds + ((size_t)b * p.A + a) * p.ch;
    LaneTerms t;
    lane_terms(row, p.C, lane, t);
    const float lcls = cls_cost(t, p.C, gc, lane, wide, terms);
    const float iou = pair_iou(gx, gy, gw, gh, load_box(row));
    const float liou = -logf(iou + 1e-8f);                          // :86
    return (lcls + 3.0f * liou) + (both ? 0.0f : 100000.0f);        // :104-108
}

// inserts x into a descending list held one value per lane (lane i = i-th largest)
__device__ __forceinline__ void top_insert(float &top, const float x, const int lane) {
    const float up = __shfl_up_sync(0xffffffffu, top, 1);
    if (top < x) top = (lane == 0 || up >= x) ? x : up;
}

// GT g claims anchor a (matching_matrix[g][a] = 1, :348).  The first claim writes the match
// tentatively; the claim that makes the count 2 reports a conflict, which the claiming CTA resolves
// right away (the argmin over ALL GT rows does not depend on who claimed, :352-356); the image's last
// CTA finally patches the resolved values over the tentative ones.
__device__ __forceinline__ bool claim(const SimParams &p, const int b, const int a, const int g, const float iou) {
    const unsigned old = atomicAdd(&p.sel_count[(size_t)b * p.A + a], 1u);
    if (old == 0u) {
        p.fg_mask[(size_t)b * p.A + a] = 1;
        p.matched_gt[(size_t)b * p.A + a] = g;
        p.matched_iou[(size_t)b * p.A + a] = iou;  // :367
        atomicAdd(&p.meta[b * 8 + 3], 1);
    }
    return old == 1u;
}

__device__ void resolve_conflict(const SimParams &p, const int b, const int a, const int G, const bool wide, float *T);

// ---- K3 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMatchWarps * 32) simota_match_kernel(const SimParams p) {
    __shared__ int s_anchor[kMaxBoth];
    __shared__ float s_cost[kMaxBoth];
    __shared__ float s_terms[96];
    __shared__ float s_iou[kMaxBoth];
    __shared__ float s_top[kMatchWarps][10];
    __shared__ float s_T[kMatchWarps][96];
    __shared__ int s_conf[64];
    __shared__ int s_k, s_nb, s_last, s_nc;
    const int b = blockIdx.y, g = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.meta[b * 8 + 0], Nc = p.meta[b * 8 + 1];
    if (g >= G || Nc == 0) return;
    const float *L = p.labels + ((size_t)b * p.Lmax + g) * 5;
    const int gc = (int)L[0];  // .to(int64) truncates (:89)
    const float gx = L[1], gy = L[2], gw = L[3], gh = L[4];
    const int *ca = p.cand_anchor + (size_t)b * p.A;
    const float4 *cb = p.cand_box + (size_t)b * p.A;
    const float *car = p.cand_area + (size_t)b * p.A;
    // ATen picks a 64-wide block for sum(-1) when the [G,Nc] output has fewer than 16 elements (and C >= 64)
    const bool wide = (long long)G * Nc < 16 && p.C >= 64;

    // ---- top-10 IoU values over all candidates (iou_loss.py:400-414 with xyxy=False; GT-side operands
    // hoisted).  Lists start at +0: a pair that does not overlap has IoU (+/-)0 and can never displace
    // anything; the division runs only when the quotient could exceed the current 10th value.
    const float g_x1 = gx - gw / 2, g_y1 = gy - gh / 2, g_x2 = gx + gw / 2, g_y2 = gy + gh / 2;
    const float area_a = gw * gh;
    float top = 0.f, thresh = 0.f;
    const int per = (((Nc + kMatchWarps - 1) / kMatchWarps) + 31) & ~31;
    const int n_lo = warp * per, n_hi = min(Nc, n_lo + per);
    for (int n0 = n_lo; n0 < n_hi; n0 += 32) {
        const int n = n0 + lane;
        float v = 0.f;
        if (n < n_hi) {
            const float4 c = cb[n];
            const float tlx = fmaxf(g_x1, c.x), tly = fmaxf(g_y1, c.y);
            const float brx = fminf(g_x2, c.z), bry = fminf(g_y2, c.w);
            if (tlx < brx && tly < bry) {  // en == 1
                const float area_i = (brx - tlx) * (bry - tly);
                const float den = (area_a + car[n]) - area_i;
                if (area_i >= (thresh * den) * 0.999f) v = area_i / den;
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, v > thresh);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            top_insert(top, __shfl_sync(0xffffffffu, v, j), lane);
        }
        thresh = __shfl_sync(0xffffffffu, top, 9);
    }
    if (lane < 10) s_top[warp][lane] = top;
    __syncthreads();
    if (warp == 0) {
        const int src = lane + 10;  // the other warps' 30 values
        const float mine = (src < 10 * kMatchWarps) ? s_top[src / 10][src % 10] : 0.f;
        for (int j = 0; j < 10 * (kMatchWarps - 1); ++j) {
            const float x = __shfl_sync(0xffffffffu, mine, j);
            if (x > thresh) {
                top_insert(top, x, lane);
                thresh = __shfl_sync(0xffffffffu, top, 9);
            }
        }
        // dynamic k = clamp(int(sum of the top min(10, Nc)), 1) with ATen's reduce tree (:336-340)
        const int kc = min(10, Nc);
        int bw = 1;
        while (bw * 2 <= kc) bw <<= 1;
        const float hi = __shfl_down_sync(0xffffffffu, top, bw);
        float v = 0.f;
        if (lane < bw) v = top + ((lane + bw < kc) ? hi : 0.f);
        for (int h = bw >> 1; h >= 1; h >>= 1) {
            const float o = __shfl_down_sync(0xffffffffu, v, h);
            if (lane < h) v = v + o;
        }
        const int k = max((int)__shfl_sync(0xffffffffu, v, 0), 1);
        // ---- anchors both in-box and in-centre, in ascending anchor order
        const short *rect = p.rect + ((size_t)b * p.Lmax + g) * p.n_levels * 8;
        int nb = 0;
        if (k < Nc - 1) {
            for (int l = 0; l < p.n_levels; ++l) {
                const short *r8 = rect + l * 8;
                const int x0 = max(r8[0], r8[4]), x1 = min(r8[1], r8[5]);
                const int y0 = max(r8[2], r8[6]), y1 = min(r8[3], r8[7]);
                if (x0 > x1 || y0 > y1) continue;
                const int wx = x1 - x0 + 1, cells = wx * (y1 - y0 + 1);
                for (int i = lane; i < cells; i += 32)
                    if (nb + i < kMaxBoth) s_anchor[nb + i] = p.off[l] + (y0 + i / wx) * p.w[l] + x0 + i % wx;
                nb = min(nb + cells, kMaxBoth);
            }
        }
        if (lane == 0) { s_k = k; s_nb = nb; s_nc = 0; }
    }
    __syncthreads();
    const int k = s_k, nb = s_nb;
    auto push_conflict = [&](const int a) {
        const int i = atomicAdd(&s_nc, 1);
        if (i < 64) s_conf[i] = a;
    };

    if (!(k < Nc - 1)) {  // quirk Q3 (:343): the GT takes EVERY candidate
        for (int n = tid; n < Nc; n += kMatchWarps * 32) {  // Nc <= 11 here (k <= 10)
            const float iou = pair_iou(gx, gy, gw, gh, load_box(p.preds + ((size_t)b * p.A + ca[n]) * p.ch));
            if (claim(p, b, ca[n], g, iou)) push_conflict(ca[n]);
        }
    } else {
        // ---- cost of the in-both anchors (:84-108): class cost from the table, IoU term on the fly
        if (p.C >= 32) {
            const int *uidx = p.u_index + (size_t)b * p.A;
            const int gcc = min(max(gc, 0), p.C - 1);
            for (int i = tid; i < nb; i += kMatchWarps * 32) {
                const int a = s_anchor[i];
                const float lcls = p.table[((size_t)b * p.A + uidx[a]) * p.C + gcc];
                const float iou = pair_iou(gx, gy, gw, gh, load_box(p.preds + ((size_t)b * p.A + a) * p.ch));
                const float liou = -logf(iou + 1e-8f);     // :86
                s_cost[i] = (lcls + 3.0f * liou) + 0.0f;   // :104-108 (in_boxes_and_center -> + 1e5 * 0)
                s_iou[i] = iou;
            }
        } else if (warp == 0) {
            for (int i = 0; i < nb; ++i) {
                const float c = pair_cost(p, b, s_anchor[i], gx, gy, gw, gh, gc, true, wide, lane, s_terms);
                if (lane == 0) {
                    s_cost[i] = c;
                    s_iou[i] = pair_iou(gx, gy, gw, gh, load_box(p.preds + ((size_t)b * p.A + s_anchor[i]) * p.ch));
                }
            }
        }
        __syncthreads();
        if (warp == 0) {
            // ---- k smallest (cost, anchor); ties -> lowest anchor index (stable sort, :342)
            const int take = min(k, nb);
            for (int r = 0; r < take; ++r) {
                unsigned long long best = ~0ull;
                for (int i = lane; i < nb; i += 32) {
                    const float c = s_cost[i];
                    if (c >= 0.f || c < 0.f) {  // not yet taken (taken entries are NaN)
                        const unsigned long long key = ((unsigned long long)float_ordered(c) << 32) | (unsigned)i;
                        best = key < best ? key : best;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                    best = other < best ? other : best;
                }
                const int i = (int)(best & 0xffffffffu);
                if (lane == 0) {
                    if (claim(p, b, s_anchor[i], g, s_iou[i])) push_conflict(s_anchor[i]);
                    s_cost[i] = __int_as_float(0x7fc00000);
                }
                __syncwarp();
            }
            if (k > nb) {
                // ---- tiny GT: fewer in-both anchors than k.  The remaining picks come from the candidates
                // whose cost carries +1e5 (quantised to 1/128, T8): smallest (cost, anchor) over all others.
                const short *rect = p.rect + ((size_t)b * p.Lmax + g) * p.n_levels * 8;
                const int need = k - nb;  // <= 10
                unsigned long long mine = ~0ull;  // lanes 0..need-1 hold the `need` smallest keys, ascending
                for (int n = 0; n < Nc; ++n) {
                    const int a = ca[n];
                    int l, x, y;
                    anchor_cell(p, a, l, x, y);
                    if (in_both(rect + l * 8, x, y)) continue;
                    float c = pair_cost(p, b, a, gx, gy, gw, gh, gc, false, wide, lane, s_terms);
                    c = __shfl_sync(0xffffffffu, c, 0);
                    const unsigned long long key = ((unsigned long long)float_ordered(c) << 32) | (unsigned)a;
                    const unsigned long long upk = __shfl_up_sync(0xffffffffu, mine, 1);
                    if (key < mine) mine = (lane == 0 || upk <= key) ? key : upk;
                }
                if (lane < need && mine != ~0ull) {
                    const int a = (int)(mine & 0xffffffffu);
                    const float iou = pair_iou(gx, gy, gw, gh, load_box(p.preds + ((size_t)b * p.A + a) * p.ch));
                    if (claim(p, b, a, g, iou)) push_conflict(a);
                }
            }
        }
    }

    // ---- resolve the conflicts this CTA created (one warp each)
    __syncthreads();
    const int nc = min(s_nc, 64);
    for (int i = warp; i < nc; i += kMatchWarps) resolve_conflict(p, b, s_conf[i], G, wide, s_T[warp]);

    // ---- the last GT CTA of the image to finish patches the resolved matches over the tentative ones
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&p.meta[b * 8 + 4], 1) == G - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence();
        const int nconf = __ldcg(p.meta + b * 8 + 5);
        for (int i = tid; i < nconf; i += kMatchWarps * 32) {
            const int a = __ldcg(p.conf_list + (size_t)b * p.A + i);
            p.matched_gt[(size_t)b * p.A + a] = __ldcg(p.res_g + (size_t)b * p.A + a);
            p.matched_iou[(size_t)b * p.A + a] = __ldcg(p.res_iou + (size_t)b * p.A + a);
        }
        if (tid == 0) p.num_fg[b] = __ldcg(p.meta + b * 8 + 3);  // :358
    }
}

// ---- conflict resolution (warp-cooperative) -----------------------------------------------------
// anchor a was claimed by several GTs: argmin of the cost column over ALL GT rows, first minimum (:352-356)
__device__ void resolve_conflict(const SimParams &p, const int b, const int a, const int G, const bool wide, float *T) {
    const int lane = threadIdx.x & 31;
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    const float *row = p.preds + ((size_t)b * p.A + a) * p.ch;
    const float4 pb = load_box(row);
    int l, x, y;
    anchor_cell(p, a, l, x, y);
    const bool fast = p.C >= 32;
    if (fast) {
        const int u = p.u_index[(size_t)b * p.A + a];
        if (u >= 0) {
            const float *src = p.table + ((size_t)b * p.A + u) * p.C;
            for (int c = lane; c < 96; c += 32) T[c] = c < p.C ? src[c] : 0.f;
        } else {
            // not in any in-both set (claimed through the +1e5 region): class costs from one reduction,
            // butterfly over the all-negative leaves + positive-leaf swap (same tree as the table kernel)
            LaneTerms t;
            lane_terms(row, p.C, lane, t);
            float s = lane_combine(t.neg[0], t.neg[1], t.neg[2], wide);
            float sib[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                sib[i] = __shfl_xor_sync(0xffffffffu, s, 16 >> i);
                s = s + sib[i];
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = lane + 32 * j;
                if (c < p.C) {
                    const float e0 = j == 0 ? pos_term(t.p[0]) : t.neg[0];
                    const float e1 = j == 1 ? pos_term(t.p[1]) : t.neg[1];
                    const float e2 = j == 2 ? pos_term(t.p[2]) : t.neg[2];
                    float r = lane_combine(e0, e1, e2, wide);
#pragma unroll
                    for (int i = 0; i < 5; ++i) r = r + sib[i];
                    T[c] = r;
                }
            }
        }
        __syncwarp();
    }
    unsigned long long best = ~0ull;
    if (fast) {
        for (int g = lane; g < G; g += 32) {
            const float *gr = L + 5 * g;
            const short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
            const float iou = pair_iou(gr[1], gr[2], gr[3], gr[4], pb);
            const float liou = -logf(iou + 1e-8f);
            const float cost = (T[min(max((int)gr[0], 0), p.C - 1)] + 3.0f * liou) + (in_both(r8, x, y) ? 0.0f : 100000.0f);
            const unsigned long long key = ((unsigned long long)float_ordered(cost) << 32) | (unsigned)g;
            best = key < best ? key : best;
        }
    } else {
        for (int g = 0; g < G; ++g) {
            const float *gr = L + 5 * g;
            const short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
            const float cost = pair_cost(p, b, a, gr[1], gr[2], gr[3], gr[4], (int)gr[0], in_both(r8, x, y), wide, lane, T);
            if (lane == 0) {
                const unsigned long long key = ((unsigned long long)float_ordered(cost) << 32) | (unsigned)g;
                best = key < best ? key : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
    }
    if (lane == 0) {
        const int g = (int)(best & 0xffffffffu);
        const float *gr = L + 5 * g;
        p.res_g[(size_t)b * p.A + a] = g;
        p.res_iou[(size_t)b * p.A + a] = pair_iou(gr[1], gr[2], gr[3], gr[4], pb);
        p.conf_list[(size_t)b * p.A + atomicAdd(&p.meta[b * 8 + 5], 1)] = a;
    }
    __syncwarp();
}

static size_t sim_ws_layout(int B, int A, int Lmax, int n_levels, SimParams *p, unsigned char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_meta = take((size_t)B * 8 * sizeof(int));
    const size_t o_sl = take((size_t)B * A * sizeof(int));
    const size_t o_ri = take((size_t)B * A * sizeof(float));
    const size_t o_ca = take((size_t)B * A * sizeof(int));
    const size_t o_cb = take((size_t)B * A * sizeof(float4));
    const size_t o_car = take((size_t)B * A * sizeof(float));
    const size_t o_ua = take((size_t)B * A * sizeof(int));
    const size_t o_ui = take((size_t)B * A * sizeof(int));
    const size_t o_tab = take((size_t)B * A * PLYOLO_MAX_CLASSES * sizeof(float));
    const size_t o_sc = take((size_t)B * A * sizeof(unsigned));
    const size_t o_sm = take((size_t)B * A * sizeof(unsigned));
    const size_t o_rect = take((size_t)B * Lmax * n_levels * 8 * sizeof(short));
    if (p) {
        p->meta = reinterpret_cast<int *>(base + o_meta);
        p->conf_list = reinterpret_cast<int *>(base + o_sl);
        p->res_iou = reinterpret_cast<float *>(base + o_ri);
        p->cand_anchor = reinterpret_cast<int *>(base + o_ca);
        p->cand_box = reinterpret_cast<float4 *>(base + o_cb);
        p->cand_area = reinterpret_cast<float *>(base + o_car);
        p->u_anchor = reinterpret_cast<int *>(base + o_ua);
        p->u_index = reinterpret_cast<int *>(base + o_ui);
        p->table = reinterpret_cast<float *>(base + o_tab);
        p->sel_count = reinterpret_cast<unsigned *>(base + o_sc);
        p->res_g = reinterpret_cast<int *>(base + o_sm);
        p->rect = reinterpret_cast<short *>(base + o_rect);
    }
    return off;
}

}  // namespace plyolo

extern "C" size_t plyolo_simota_workspace_bytes(int B, int A, int Lmax, int n_levels) {
    if (B < 1 || A < 1 || Lmax < 1 || n_levels < 1) return 0;
    return plyolo::sim_ws_layout(B, A, Lmax, n_levels, nullptr, nullptr);
}

extern "C" int plyolo_simota_f32(const float *preds, const float *labels, int B, int A, int C, int Lmax,
                                 const int *hs, const int *ws, const int *strides, int n_levels, uint8_t *fg_mask,
                                 int32_t *matched_gt, float *matched_iou, int32_t *num_fg, int32_t *num_gt,
                                 void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(preds && labels, "preds / labels is null");
    PLYOLO_REQUIRE(fg_mask && matched_gt && matched_iou && num_fg && num_gt, "an output pointer is null");
    PLYOLO_REQUIRE(B >= 1 && B <= 65535, "B=%d not in [1,65535]", B);
    PLYOLO_REQUIRE(C >= 1 && C <= PLYOLO_MAX_CLASSES, "C=%d not in [1,%d]", C, PLYOLO_MAX_CLASSES);
    PLYOLO_REQUIRE(Lmax >= 1 && Lmax <= 32767, "Lmax=%d not in [1,32767]", Lmax);
    PLYOLO_REQUIRE(hs && ws && strides, "null level description");
    PLYOLO_REQUIRE(n_levels >= 1 && n_levels <= PLYOLO_MAX_LEVELS, "n_levels=%d not in [1,%d]", n_levels,
                   PLYOLO_MAX_LEVELS);
    SimParams p;
    int off = 0;
    for (int l = 0; l < PLYOLO_MAX_LEVELS; ++l) {
        if (l < n_levels) {
            PLYOLO_REQUIRE(hs[l] > 0 && ws[l] > 0 && strides[l] > 0, "level %d has a non-positive dimension", l);
            PLYOLO_REQUIRE(hs[l] == ws[l], "level %d is %dx%d: the reference grid is only defined for square maps", l,
                           hs[l], ws[l]);
            PLYOLO_REQUIRE(hs[l] <= 32767, "level %d too large", l);
            p.hw[l] = hs[l] * ws[l]; p.w[l] = ws[l]; p.off[l] = off; p.stride[l] = (float)strides[l];
            off += p.hw[l];
        } else {
            p.hw[l] = 0; p.w[l] = 1; p.off[l] = off; p.stride[l] = 1.f;
        }
    }
    PLYOLO_REQUIRE(off == A, "A=%d does not match the level shapes (sum H*W = %d)", A, off);
    if (!workspace || ((uintptr_t)workspace & 255) ||
        workspace_bytes < plyolo_simota_workspace_bytes(B, A, Lmax, n_levels)) {
        set_error("workspace null, not 256-byte aligned, or smaller than plyolo_simota_workspace_bytes()");
        return PLYOLO_ERR_WORKSPACE;
    }
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    p.preds = preds; p.labels = labels; p.B = B; p.A = A; p.C = C; p.ch = 5 + C; p.Lmax = Lmax; p.n_levels = n_levels;
    p.fg_mask = fg_mask; p.matched_gt = matched_gt; p.matched_iou = matched_iou; p.num_fg = num_fg; p.num_gt = num_gt;
    sim_ws_layout(B, A, Lmax, n_levels, &p, static_cast<unsigned char *>(workspace));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bm = 2 * (size_t)((A + 31) / 32) * sizeof(unsigned);
    PLYOLO_REQUIRE(bm <= 160 * 1024, "A=%d too large for the candidate bitmaps", A);
    if (bm > 40 * 1024) cudaFuncSetAttribute(simota_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm);
    simota_prep_kernel<<<B, kPrepThreads, bm, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_prep_kernel");
    if (C >= 32) {
        simota_cost_table_kernel<<<dim3(kTableCtas, B), kTableWarps * 32, 0, st>>>(p);
        PLYOLO_CHECK_LAUNCH("simota_cost_table_kernel");
    }
    simota_match_kernel<<<dim3(Lmax, B), kMatchWarps * 32, 0, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_match_kernel");
    return PLYOLO_OK;
}

// ---- bboxes_iou (iou_loss.py:391-414) -------------------------------------------------------
namespace plyolo {
__global__ void bboxes_iou_kernel(const float *a, int na, const float *b, int nb, int xyxy, float *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)na * nb) return;
    const float *p = a + 4 * (i / nb), *q = b + 4 * (i % nb);
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
    const float en = (tlx < brx ? 1.f : 0.f) * (tly < bry ? 1.f : 0.f);
    const float ai = ((brx - tlx) * (bry - tly)) * en;
    out[i] = ai / ((aa + ab) - ai);
}
}  // namespace plyolo

extern "C" int plyolo_bboxes_iou_f32(const float *a, int na, const float *b, int nb, int xyxy, float *out,
                                     plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(na >= 0 && nb >= 0, "negative box count");
    if ((long long)na * nb == 0) return PLYOLO_OK;
    PLYOLO_REQUIRE(a && b && out, "null pointer");
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    const long long n = (long long)na * nb;
    bboxes_iou_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a, na, b, nb, xyxy, out);
    PLYOLO_CHECK_LAUNCH("bboxes_iou_kernel");
    return PLYOLO_OK;
}
