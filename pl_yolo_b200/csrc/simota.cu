// simota.cu — SimOTA label assignment for a whole batch (yolox_loss.py:43-118, :231-370;
// iou_loss.py:391-414).  Three launches, no host synchronisation, no [G,Nc,80] temporaries.
//
//  K1 simota_prep_kernel      one CTA per image
//     GT count (:43), closed-form geometry prior: for every (GT, level) the in-box and in-centre
//     anchors are axis-aligned cell rectangles, found with the reference's own fp32 comparisons
//     (edge rounded first, then the delta, :249-307) and rasterised into a shared-memory bitmap
//     (fg_mask = union, :310); the bitmap is compacted in anchor order into the candidate list and
//     the candidates' decoded boxes are gathered (16 of the 340 bytes of each prediction row).
//  K2 simota_match_kernel     one warp per GT
//     IoU against every candidate with a warp-resident top-10 (values only, :336-340) -> dynamic k
//     with ATen's reduce tree -> cost only for the <= 25*levels anchors that are both in-box and
//     in-centre (every other cost carries +1e5, :104-108, so the k smallest live there unless the GT
//     is tiny) -> k smallest (cost, anchor) -> per-anchor match count / lowest-GT atomics.
//     The 80-class BCE cost is evaluated by the warp with lanes over classes and summed in ATen's
//     CUDA reduce order (measured on B200: lane t adds classes t, t+32, t+64, then a halving tree).
//  K3 simota_finalize_kernel  one CTA per image
//     anchors matched once take that GT; anchors matched more than once take the argmin of the cost
//     over ALL GTs (:352-356, quirk Q4) — the per-class costs of such an anchor come from one
//     butterfly reduction by swapping the positive-class leaf — then fg_mask / matched GT /
//     matched IoU are written densely per anchor and num_fg is counted (:357-369).
#include <cfloat>

#include "common.cuh"

namespace plyolo {

constexpr int kPrepThreads = 512;
constexpr int kMatchWarps = 8;
constexpr int kFinThreads = 512;
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
    int *meta;            // [B,4]  G, Nc, n_conflict
    int *cand_anchor;     // [B,A]  (reused as the conflict list by K3)
    float4 *cand_box;     // [B,A]  decoded (cx,cy,w,h) of candidate n
    unsigned *sel_count;  // [B,A]
    unsigned *sel_ming;   // [B,A]
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
__global__ void __launch_bounds__(kPrepThreads) simota_prep_kernel(const SimParams p) {
    extern __shared__ unsigned bitmap[];  // [ceil(A/32)]
    __shared__ int s_G, s_warp[kPrepThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwords = (p.A + 31) >> 5;
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    if (tid == 0) s_G = 0;
    for (int i = tid; i < nwords; i += kPrepThreads) bitmap[i] = 0u;
    for (int a = tid; a < p.A; a += kPrepThreads) {
        p.sel_count[(size_t)b * p.A + a] = 0u;
        p.sel_ming[(size_t)b * p.A + a] = 0xffffffffu;
    }
    __syncthreads();
    // :43 nlabel = (labels.sum(2) > 0).sum(1)
    int cnt = 0;
    for (int g = tid; g < p.Lmax; g += kPrepThreads) cnt += row_sum5(L + 5 * g) > 0.f ? 1 : 0;
    if (cnt) atomicAdd(&s_G, cnt);
    __syncthreads();
    const int G = s_G;  // the GTs are rows [0, G) (:64-65)

    // geometry prior: (GT, level, {in-box, in-centre}) work items
    const int items = G * p.n_levels * 2;
    for (int it = tid; it < items; it += kPrepThreads) {
        const int which = it & 1;
        const int l = (it >> 1) % p.n_levels;
        const int g = (it >> 1) / p.n_levels;
        const float gx = L[5 * g + 1], gy = L[5 * g + 2], gw = L[5 * g + 3], gh = L[5 * g + 4];
        const float s = p.stride[l];
        const int W = p.w[l], H = p.hw[l] / p.w[l];
        float lo_x, hi_x, lo_y, hi_y;
        if (which == 0) {  // :249-268
            lo_x = gx - 0.5f * gw; hi_x = gx + 0.5f * gw; lo_y = gy - 0.5f * gh; hi_y = gy + 0.5f * gh;
        } else {           // :284-298, center_radius = 2.5
            const float r = 2.5f * s;
            lo_x = gx - r; hi_x = gx + r; lo_y = gy - r; hi_y = gy + r;
        }
        int x0, x1, y0, y1;
        cell_range(lo_x, hi_x, s, W, x0, x1);
        cell_range(lo_y, hi_y, s, H, y0, y1);
        short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8 + which * 4;
        r8[0] = (short)x0; r8[1] = (short)x1; r8[2] = (short)y0; r8[3] = (short)y1;
        if (x0 <= x1)
            for (int y = y0; y <= y1; ++y) {
                const int p0 = p.off[l] + y * W + x0, p1 = p.off[l] + y * W + x1;
                for (int wd = p0 >> 5; wd <= (p1 >> 5); ++wd) {
                    const int lo = max(p0, wd << 5) & 31, hi = min(p1, (wd << 5) + 31) & 31;
                    const unsigned m = (0xffffffffu >> (31 - hi)) & (0xffffffffu << lo);
                    atomicOr(&bitmap[wd], m);
                }
            }
    }
    __syncthreads();

    // compaction in anchor order: candidate n <-> n-th set bit (:79-82)
    const int wpt = (nwords + kPrepThreads - 1) / kPrepThreads;
    const int w0 = min(tid * wpt, nwords), w1 = min(w0 + wpt, nwords);
    int local = 0;
    for (int i = w0; i < w1; ++i) {
        unsigned m = bitmap[i];
        if (i == nwords - 1 && (p.A & 31)) m &= (1u << (p.A & 31)) - 1u;
        local += __popc(m);
    }
    int inc = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = inc - local, total = 0;
    for (int w = 0; w < kPrepThreads / 32; ++w) {
        if (w < warp) base += s_warp[w];
        total += s_warp[w];
    }
    int *ca = p.cand_anchor + (size_t)b * p.A;
    float4 *cb = p.cand_box + (size_t)b * p.A;
    for (int i = w0; i < w1; ++i) {
        unsigned m = bitmap[i];
        if (i == nwords - 1 && (p.A & 31)) m &= (1u << (p.A & 31)) - 1u;
        while (m) {
            const int bit = __ffs(m) - 1;
            m &= m - 1;
            const int a = (i << 5) + bit;
            ca[base] = a;
            cb[base] = load_box(p.preds + ((size_t)b * p.A + a) * p.ch);
            ++base;
        }
    }
    if (tid == 0) {
        p.meta[b * 4 + 0] = G;
        p.meta[b * 4 + 1] = total;
        p.meta[b * 4 + 2] = 0;
        p.num_gt[b] = G;
    }
}

// ---- K2 ------------------------------------------------------------------------------------
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

__global__ void __launch_bounds__(kMatchWarps * 32) simota_match_kernel(const SimParams p) {
    __shared__ int s_anchor[kMatchWarps][kMaxBoth];
    __shared__ float s_cost[kMatchWarps][kMaxBoth];
    __shared__ float s_terms[kMatchWarps][96];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * kMatchWarps + warp;
    const int G = p.meta[b * 4 + 0], Nc = p.meta[b * 4 + 1];
    if (g >= G || Nc == 0) return;
    const float *L = p.labels + ((size_t)b * p.Lmax + g) * 5;
    const int gc = (int)L[0];  // .to(int64) truncates (:89)
    const float gx = L[1], gy = L[2], gw = L[3], gh = L[4];
    const int *ca = p.cand_anchor + (size_t)b * p.A;
    const float4 *cb = p.cand_box + (size_t)b * p.A;
    unsigned *scount = p.sel_count + (size_t)b * p.A;
    unsigned *sming = p.sel_ming + (size_t)b * p.A;
    // ATen picks a 64-wide block for sum(-1) when the [G,Nc] output has fewer than 16 elements (and C >= 64)
    const bool wide = (long long)G * Nc < 16 && p.C >= 64;

    // ---- top-10 IoU values (sorted descending across lanes 0..31)
    float top = -1.f, thresh = -1.f;
    for (int n0 = 0; n0 < Nc; n0 += 32) {
        const int n = n0 + lane;
        const float v = n < Nc ? pair_iou(gx, gy, gw, gh, cb[n]) : -2.f;
        unsigned m = __ballot_sync(0xffffffffu, v > thresh);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const float x = __shfl_sync(0xffffffffu, v, j);
            const float up = __shfl_up_sync(0xffffffffu, top, 1);
            if (top < x) top = (lane == 0 || up >= x) ? x : up;
        }
        thresh = __shfl_sync(0xffffffffu, top, 9);
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
    int k = max((int)__shfl_sync(0xffffffffu, v, 0), 1);

    if (!(k < Nc - 1)) {  // quirk Q3 (:343): the GT takes EVERY candidate
        for (int n = lane; n < Nc; n += 32) {
            atomicAdd(&scount[ca[n]], 1u);
            atomicMin(&sming[ca[n]], (unsigned)g);
        }
        return;
    }

    // ---- anchors both in-box and in-centre, in ascending anchor order
    const short *rect = p.rect + ((size_t)b * p.Lmax + g) * p.n_levels * 8;
    int nb = 0;
    for (int l = 0; l < p.n_levels; ++l) {
        const short *r8 = rect + l * 8;
        const int x0 = max(r8[0], r8[4]), x1 = min(r8[1], r8[5]);
        const int y0 = max(r8[2], r8[6]), y1 = min(r8[3], r8[7]);
        if (x0 > x1 || y0 > y1) continue;
        const int wx = x1 - x0 + 1, cells = wx * (y1 - y0 + 1);
        for (int i = lane; i < cells; i += 32)
            if (nb + i < kMaxBoth) s_anchor[warp][nb + i] = p.off[l] + (y0 + i / wx) * p.w[l] + x0 + i % wx;
        nb = min(nb + cells, kMaxBoth);
    }
    __syncwarp();
    for (int i = 0; i < nb; ++i) {
        const float c = pair_cost(p, b, s_anchor[warp][i], gx, gy, gw, gh, gc, true, wide, lane, s_terms[warp]);
        if (lane == 0) s_cost[warp][i] = c;
    }
    __syncwarp();

    // ---- k smallest (cost, anchor); ties -> lowest anchor index (stable sort, :342)
    const int take = min(k, nb);
    for (int r = 0; r < take; ++r) {
        unsigned long long best = ~0ull;
        for (int i = lane; i < nb; i += 32) {
            const float c = s_cost[warp][i];
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
            const int a = s_anchor[warp][i];
            atomicAdd(&scount[a], 1u);
            atomicMin(&sming[a], (unsigned)g);
            s_cost[warp][i] = __int_as_float(0x7fc00000);
        }
        __syncwarp();
    }
    if (k <= nb) return;

    // ---- tiny GT: fewer in-both anchors than k.  The remaining picks come from the candidates whose
    // cost carries +1e5 (quantised to 1/128, T8): smallest (cost, anchor) over all other candidates.
    const int need = k - nb;  // <= 10
    unsigned long long mine = ~0ull;  // lanes 0..need-1 hold the `need` smallest keys, ascending
    for (int n = 0; n < Nc; ++n) {
        const int a = ca[n];
        int l, x, y;
        anchor_cell(p, a, l, x, y);
        if (in_both(rect + l * 8, x, y)) continue;
        float c = pair_cost(p, b, a, gx, gy, gw, gh, gc, false, wide, lane, s_terms[warp]);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long key = ((unsigned long long)float_ordered(c) << 32) | (unsigned)a;
        const unsigned long long upk = __shfl_up_sync(0xffffffffu, mine, 1);
        if (key < mine) mine = (lane == 0 || upk <= key) ? key : upk;
    }
    if (lane < need && mine != ~0ull) {
        const int a = (int)(mine & 0xffffffffu);
        atomicAdd(&scount[a], 1u);
        atomicMin(&sming[a], (unsigned)g);
    }
}

// ---- K3 ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kFinThreads) simota_finalize_kernel(const SimParams p) {
    __shared__ int s_nconf, s_nfg;
    __shared__ float s_T[kFinThreads / 32][96];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.meta[b * 4 + 0], Nc = p.meta[b * 4 + 1];
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    const unsigned *scount = p.sel_count + (size_t)b * p.A;
    const unsigned *sming = p.sel_ming + (size_t)b * p.A;
    int *conflicts = p.cand_anchor + (size_t)b * p.A;  // candidate list is dead by now
    uint8_t *FG = p.fg_mask + (size_t)b * p.A;
    int32_t *MG = p.matched_gt + (size_t)b * p.A;
    float *MI = p.matched_iou + (size_t)b * p.A;
    if (tid == 0) { s_nconf = 0; s_nfg = 0; }
    __syncthreads();
    int nfg = 0;
    for (int a = tid; a < p.A; a += kFinThreads) {
        const unsigned c = scount[a];
        if (c == 0) {
            FG[a] = 0; MG[a] = -1; MI[a] = 0.f;
        } else {
            ++nfg;
            if (c == 1) {
                const int g = (int)sming[a];
                const float *gr = L + 5 * g;
                FG[a] = 1; MG[a] = g;
                MI[a] = pair_iou(gr[1], gr[2], gr[3], gr[4], load_box(p.preds + ((size_t)b * p.A + a) * p.ch));  // :367
            } else {
                conflicts[atomicAdd(&s_nconf, 1)] = a;
            }
        }
    }
    if (nfg) atomicAdd(&s_nfg, nfg);
    __syncthreads();
    const int nconf = s_nconf;
    const bool wide = (long long)G * Nc < 16 && p.C >= 64;
    // anchors claimed by several GTs: argmin of the cost column over ALL GT rows, first minimum (:352-356)
    for (int ci = warp; ci < nconf; ci += kFinThreads / 32) {
        const int a = conflicts[ci];
        const float *row = p.preds + ((size_t)b * p.A + a) * p.ch;
        const float4 pb = load_box(row);
        int l, x, y;
        anchor_cell(p, a, l, x, y);
        float *T = s_T[warp];
        const bool fast = p.C >= 32;
        if (fast) {
            // class cost for every possible GT class from ONE reduction: butterfly over the all-negative
            // leaves keeps, per lane, the sibling sums of its path; swapping the positive leaf re-adds them.
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
            for (int c = p.C + lane; c < 96; c += 32) T[c] = s;  // class id outside [0,C): no positive leaf
            __syncwarp();
        }
        unsigned long long best = ~0ull;
        if (fast) {
            for (int g = lane; g < G; g += 32) {
                const float *gr = L + 5 * g;
                const short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
                const float iou = pair_iou(gr[1], gr[2], gr[3], gr[4], pb);
                const float liou = -logf(iou + 1e-8f);
                const float cost = (T[min(max((int)gr[0], 0), 95)] + 3.0f * liou) + (in_both(r8, x, y) ? 0.0f : 100000.0f);
                const unsigned long long key = ((unsigned long long)float_ordered(cost) << 32) | (unsigned)g;
                best = key < best ? key : best;
            }
        } else {
            for (int g = 0; g < G; ++g) {
                const float *gr = L + 5 * g;
                const short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
                const float cost = pair_cost(p, b, a, gr[1], gr[2], gr[3], gr[4], (int)gr[0], in_both(r8, x, y), wide,
                                             lane, T);
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
            FG[a] = 1; MG[a] = g;
            MI[a] = pair_iou(gr[1], gr[2], gr[3], gr[4], pb);
        }
        __syncwarp();
    }
    if (tid == 0) p.num_fg[b] = s_nfg;  // :358
}

static size_t sim_ws_layout(int B, int A, int Lmax, int n_levels, SimParams *p, unsigned char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_meta = take((size_t)B * 4 * sizeof(int));
    const size_t o_ca = take((size_t)B * A * sizeof(int));
    const size_t o_cb = take((size_t)B * A * sizeof(float4));
    const size_t o_sc = take((size_t)B * A * sizeof(unsigned));
    const size_t o_sm = take((size_t)B * A * sizeof(unsigned));
    const size_t o_rect = take((size_t)B * Lmax * n_levels * 8 * sizeof(short));
    if (p) {
        p->meta = reinterpret_cast<int *>(base + o_meta);
        p->cand_anchor = reinterpret_cast<int *>(base + o_ca);
        p->cand_box = reinterpret_cast<float4 *>(base + o_cb);
        p->sel_count = reinterpret_cast<unsigned *>(base + o_sc);
        p->sel_ming = reinterpret_cast<unsigned *>(base + o_sm);
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
    const size_t bm = (size_t)((A + 31) / 32) * sizeof(unsigned);
    PLYOLO_REQUIRE(bm <= 160 * 1024, "A=%d too large for the candidate bitmap", A);
    if (bm > 40 * 1024) cudaFuncSetAttribute(simota_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm);
    simota_prep_kernel<<<B, kPrepThreads, bm, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_prep_kernel");
    simota_match_kernel<<<dim3((Lmax + kMatchWarps - 1) / kMatchWarps, B), kMatchWarps * 32, 0, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_match_kernel");
    simota_finalize_kernel<<<B, kFinThreads, 0, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_finalize_kernel");
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
