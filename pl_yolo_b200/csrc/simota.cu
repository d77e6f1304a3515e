// simota.cu — SimOTA label assignment for a whole batch (yolox_loss.py:43-118, :231-370;
// iou_loss.py:391-414).  Three launches, no host synchronisation, no [G,Nc,80] temporaries.
//
//  K1 simota_prep_kernel      4 CTAs per image
//     GT count (:43), closed-form geometry prior: for every (GT, level) the in-box and in-centre
//     anchors are axis-aligned cell rectangles, found with the reference's own fp32 comparisons
//     (edge rounded first, then the delta, :249-307) and rasterised into a shared-memory bitmap
//     fg = union of all rectangles (:310).  Every CTA builds the whole bitmap (cheap) and takes a
//     quarter of the anchors: background defaults of the outputs, ordered compaction of the
//     candidates (candidate n <-> n-th set bit, :79-82) and the gather of their corner boxes / areas
//     (16 of the 340 bytes of each prediction row) for the IoU sweep.
//  K2a simota_sweep_kernel    4 warps per GT
//     top-10 IoU values (:336-340) of every GT -> dynamic k; groups of 32 candidates whose union box misses the GT
//     are skipped (their IoU is 0); per-lane top-6 lists, exact fallback when they cannot prove themselves.
//  K2b simota_match_kernel    one CTA (16 warps) per 8 GTs of an image (dealt by prep: balanced by #in-both anchors)
//     1. cost (:84-108) only for the <= 25 * levels anchors that are both in-box and in-centre (every
//        other cost carries +1e5, so the k smallest live there unless the GT is tiny), and of those
//        only the pairs that can still be among the k smallest: a rigorous tight lower bound (exact positive BCE
//        leaf + IoU term + hardware-approximation negative leaves minus their error bound) prunes the expensive
//        80-class sweep (one warp per pair, ATen's CUDA reduce order: lane t adds classes t, t+32, t+64, then a
//        halving tree) to about k + 1 pairs per GT;
//     2. k smallest (cost, anchor) per GT -> per-anchor match count / tentative match by atomics;
//        anchors claimed twice are resolved right away by the argmin of the cost over ALL GTs
//        (:352-356, quirk Q4);
//     3. the image's last CTA to finish patches the resolved matches over the tentative ones and
//        publishes num_fg (:357-369).
#include <cfloat>

#include "common.cuh"

namespace plyolo {

constexpr int kPrepThreads = 512;
#ifndef PLYOLO_PREP_SPLIT
#define PLYOLO_PREP_SPLIT 4
#endif
constexpr int kPrepSplit = PLYOLO_PREP_SPLIT;  // CTAs per image in the prep kernel
constexpr int kGtPerCta = 8;      // GTs per CTA in the match kernel
constexpr int kMatchWarps = 16;  // two per GT during the IoU sweep
constexpr int kMatchThreads = kMatchWarps * 32;
constexpr int kSweepSub = 4;      // IoU sweep: warps per GT (each takes every 4th candidate group the GT overlaps)
constexpr int kSweepGts = 2;      // GTs of one image per CTA
constexpr int kSweepThreads = kSweepGts * kSweepSub * 32;
constexpr int kSweepChunk = 1024;  // candidate groups listed per pass (all of them at 640^2)
constexpr int kMaxBoth = 36 * PLYOLO_MAX_LEVELS;  // 5x5 centre cells per level (6x6 if an edge rounds outward)
constexpr int kPrefetchRows = 2;  // L2 prefetch of the in-both anchors' prediction rows: 0 none, 1 first line, 2 whole row
constexpr bool kTightAll = true;  // tight lower bounds for every pair up front (else: only for pairs the cheap bound cannot prune)
constexpr int kExtraEval = 1;    // exact costs evaluated beyond k in the first round (tightens the pruning bound)
constexpr int kMaxConf = 128;     // conflicts one CTA can create (<= 11 per GT)
constexpr int kTinyWindow = 2048; // candidates per pass of the tiny-GT search

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
    int *meta;            // [B,8]  G, Nc, -, #matched anchors, #finished match CTAs, #conflict anchors
    int *conf_list;       // [B,A]  anchors claimed by more than one GT (unordered)
    int *cand_anchor;     // [B,A]  candidate n -> anchor
    float4 *cand_box;     // [B,A]  corners (cx-w/2, cy-h/2, cx+w/2, cy+h/2) of candidate n
    float *cand_area;     // [B,A]  w*h of candidate n
    unsigned *sel_count;  // [B,A]  number of GTs that claimed the anchor
    int *res_g;           // [B,A]  conflict resolution: argmin GT ...
    float *res_iou;       // [B,A]  ... and its IoU
    short *rect;          // [B,Lmax,n_levels,8] in-box x0,x1,y0,y1 | in-centre x0,x1,y0,y1 (inclusive)
    float4 *grp_box;      // [B,A/32+1] union box of every group of 32 consecutive candidates
    int *dyn_k;           // [B,Lmax] dynamic k of every GT (:336-340)
    int *gt_perm;         // [B,Lmax+8] match CTA c handles the GTs gt_perm[8c .. 8c+7] (-1 = none): balanced by #in-both anchors
    int force_exact;      // debug: the IoU sweep takes its exact warp-wide list for every GT
    long long *prof;      // debug: [B][gridDim.x][16] phase timestamps of the match kernel, or null
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

// one prediction row as the lanes of a warp hold it: class logits lane + 32 j, objectness, box
struct RawRow {
    float x[3];
    float obj;
    float4 box;
};
__device__ __forceinline__ void load_row(const float *row, const int C, const int lane, RawRow &r) {
    r.obj = __ldg(row + 4);
#pragma unroll
    for (int j = 0; j < 3; ++j) r.x[j] = (lane + 32 * j < C) ? __ldg(row + 5 + lane + 32 * j) : 0.f;
    r.box = make_float4(__ldg(row), __ldg(row + 1), __ldg(row + 2), __ldg(row + 3));
}

// BCE leaves of one prediction row, lanes over classes (yolox_loss.py:94-101).
__device__ __forceinline__ void lane_terms(const RawRow &r, const int C, const int lane, LaneTerms &t) {
    const float so = sigmoid_ref(r.obj);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
            const float p = sqrtf(sigmoid_ref(r.x[j]) * so);
            t.p[j] = p;
            t.neg[j] = -fmaxf(log1pf(-p), -100.f);  // ATen BCE, target 0: (0-1)*max(log1p(-p),-100)
        } else {
            t.p[j] = 0.f;
            t.neg[j] = 0.f;
        }
    }
}
__device__ __forceinline__ float pos_term(const float p) { return -fmaxf(logf(p), -100.f); }  // target 1

// One negative BCE leaf -max(log1p(-p), -100), p = sqrt(sigmoid(x) * so) = rsqrt(1/so + e^-x / so), from the
// hardware approximations (ex2, rsqrt; lg2 only when some p > 0.3), within kLeafErr of the exact fp32 leaf or below it:
// the relative error of p stays below 2e-6, so for p <= 0.9 (|d leaf / d p| <= 10) the leaf of the approximate p is
// within 2e-5 of the exact one; below 0.3 the truncated series p + p^2/2 + p^3/3 <= -log(1 - p) replaces the
// logarithm; above 0.9 the true leaf exceeds -log(0.1 + 1e-5) > 2.3 and 2 is returned.
// inv_so = 1 / so, lso = log2(inv_so) (inf is fine: p = 0).  Explicit fma: this is a bound, not reference arithmetic.
constexpr float kLeafErr = 5.0e-5f;
// p of one leaf (branch-free: the leaves of a row are independent chains)
__device__ __forceinline__ float leaf_p_approx(const float x, const float inv_so, const float lso) {
    float e, pr;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(__fmaf_rn(x, -1.4426950408889634f, lso)));  // e^-x / so
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(pr) : "f"(inv_so + e));
    return pr;
}
__device__ __forceinline__ float leaf_series(const float pr) {  // p + p^2/2 + p^3/3 <= -log(1 - p)
    return __fmaf_rn(pr * pr, __fmaf_rn(pr, 0.3333333f, 0.5f), pr);
}
__device__ __forceinline__ float leaf_general(const float pr) {  // any p: series, logarithm or the constant
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f - pr));
    const float n = pr > 0.3f ? l * -0.6931471805599453f : leaf_series(pr);
    return pr > 0.9f ? 2.0f : n;
}

// ATen's strided accumulators for one lane: wide == false: 32 lanes ((e0+e1)+e2); wide == true: the
// 64-lane block used when there are fewer than 16 outputs, folded to 32 lanes ((e0+e2)+(e1+0)).
__device__ __forceinline__ float lane_combine(const float e0, const float e1, const float e2, const bool wide) {
    return wide ? ((e0 + e2) + e1) : ((e0 + e1) + e2);
}

// Sum of the C BCE leaves for GT class `gc` in ATen's CUDA reduce order; result valid in lane 0.
// C >= 32 path (registers only); smaller C goes through `terms` in shared memory.
__device__ __forceinline__ float cls_cost(const LaneTerms &t, const int C, const int gc, const int lane,
                                          const bool wide, float *terms /* per-warp [96] */) {
    float e[3] = {t.neg[0], t.neg[1], t.neg[2]};
    if (lane == (gc & 31) && gc >= 0 && gc < 96) {  // only the lane that owns the GT class evaluates the log
        const int j = gc >> 5;
        const float pp = j == 0 ? t.p[0] : (j == 1 ? t.p[1] : t.p[2]);
        const float pt = pos_term(pp);
        if (j == 0) e[0] = pt; else if (j == 1) e[1] = pt; else e[2] = pt;
    }
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

__global__ void __launch_bounds__(kPrepThreads) simota_prep_kernel(const SimParams p) {
    extern __shared__ unsigned prep_smem[];  // bitmap [nwords] | word_base [nwords + 1] | this CTA's candidates [A/4 + 128] | in-both count per GT [Lmax]
    __shared__ int s_G, s_warp[kPrepThreads / 32];
    const int q = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_launch_dependents();  // the sweep grid may be scheduled behind this one (it waits for our completion itself)
    const int nwords = (p.A + 31) >> 5;
    unsigned *bm_fg = prep_smem;
    int *word_base = reinterpret_cast<int *>(prep_smem + nwords);
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    if (tid == 0) s_G = 0;
    for (int i = tid; i < nwords; i += kPrepThreads) bm_fg[i] = 0u;
    // background defaults for this CTA's quarter of the anchors (:57-62); the matching overwrites the matched ones
    const int a_per = (p.A + kPrepSplit - 1) / kPrepSplit;
    const int a_lo = q * a_per, a_hi = min(p.A, a_lo + a_per);
    for (int a = a_lo + tid; a < a_hi; a += kPrepThreads) {
        p.sel_count[(size_t)b * p.A + a] = 0u;
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

    int *s_nb = reinterpret_cast<int *>(prep_smem) + 2 * nwords + 1 + (p.A / kPrepSplit + 128);  // [Lmax] (CTA 0 only)
    if (q == 0) {
        for (int g = tid; g < p.Lmax; g += kPrepThreads) s_nb[g] = 0;
        __syncthreads();
    }
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
        if (q == 0) {
            short *r8 = p.rect + (((size_t)b * p.Lmax + g) * p.n_levels + l) * 8;
            r8[0] = (short)bx0; r8[1] = (short)bx1; r8[2] = (short)by0; r8[3] = (short)by1;
            r8[4] = (short)cx0; r8[5] = (short)cx1; r8[6] = (short)cy0; r8[7] = (short)cy1;
        }
        if (q == 0) {  // in-box AND in-centre cells of this level: the match kernel's cost pairs of the GT
            const int ix0 = max(bx0, cx0), ix1 = min(bx1, cx1), iy0 = max(by0, cy0), iy1 = min(by1, cy1);
            if (ix0 <= ix1 && iy0 <= iy1) atomicAdd(&s_nb[g], (ix1 - ix0 + 1) * (iy1 - iy0 + 1));
        }
        const int o = p.off[l];
        if (bx0 <= bx1)
            for (int y = by0; y <= by1; ++y) set_bits(bm_fg, o + y * W + bx0, o + y * W + bx1);
        if (cx0 <= cx1)
            for (int y = cy0; y <= cy1; ++y) set_bits(bm_fg, o + y * W + cx0, o + y * W + cx1);
    }
    __syncthreads();

    if (q == 0) {
        // Load balance of the match kernel: its CTA c takes the GTs gt_perm[8c .. 8c+7].  GTs sorted by their
        // number of in-both anchors (descending, ties by index) are dealt to the ceil(G/8) CTAs in snake order.
        const int ncta = (G + kGtPerCta - 1) / kGtPerCta;
        int *perm = p.gt_perm + (size_t)b * (p.Lmax + kGtPerCta);
        for (int i = tid; i < ncta * kGtPerCta; i += kPrepThreads) perm[i] = -1;
        __syncthreads();
        for (int g = tid; g < G; g += kPrepThreads) {
            const int nbg = s_nb[g];
            int r = 0;
            for (int h = 0; h < G; ++h) {
                const int nbh = s_nb[h];
                r += (nbh > nbg || (nbh == nbg && h < g)) ? 1 : 0;
            }
            const int round = r / ncta, pos = r - round * ncta;
            const int c = (round & 1) ? ncta - 1 - pos : pos;
            perm[c * kGtPerCta + round] = g;
        }
    }

    // exclusive prefix of the word popcounts: candidate n <-> n-th set bit of fg (:79-82)
    int run = 0;  // same value in every thread
    for (int w0 = 0; w0 < nwords; w0 += kPrepThreads) {
        const int w = w0 + tid;
        unsigned m = w < nwords ? bm_fg[w] : 0u;
        if (w == nwords - 1 && (p.A & 31)) m &= (1u << (p.A & 31)) - 1u;
        const int c = __popc(m);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        int base = run, tot = 0;
#pragma unroll
        for (int k = 0; k < kPrepThreads / 32; ++k) {
            if (k < warp) base += s_warp[k];
            tot += s_warp[k];
        }
        if (w < nwords) word_base[w] = base + inc - c;
        run += tot;
        __syncthreads();
    }
    const int Nc = run;
    // This CTA's share of the candidates: whole groups of 32 consecutive candidates.  The words holding them are
    // expanded into the candidate list (global, for the match kernel) and a shared copy for the gather below,
    // so no CTA depends on another CTA's part of the list.
    float4 *cb = p.cand_box + (size_t)b * p.A;
    float *car = p.cand_area + (size_t)b * p.A;
    float4 *gb = p.grp_box + (size_t)b * (p.A / 32 + 1);
    int *ca = p.cand_anchor + (size_t)b * p.A;
    int *my_list = word_base + nwords + 1;  // [n1 - n0]
    const int ngrp = (Nc + 31) >> 5;
    const int gp = (ngrp + kPrepSplit - 1) / kPrepSplit;
    const int gq_lo = min(q * gp, ngrp), gq_hi = min(gq_lo + gp, ngrp);
    const int n0 = gq_lo * 32, n1 = min(gq_hi * 32, Nc);
    auto word_of = [&](const int n) {
        int lo = 0, hi = nwords;  // word_base[lo] <= n < word_base[hi]  (word_base[nwords] == Nc)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (word_base[mid] <= n) lo = mid; else hi = mid;
        }
        return lo;
    };
    if (n0 < n1) {
        const int w_first = word_of(n0), w_last = word_of(n1 - 1);
        for (int w = w_first + tid; w <= w_last; w += kPrepThreads) {
            unsigned m = bm_fg[w];
            if (w == nwords - 1 && (p.A & 31)) m &= (1u << (p.A & 31)) - 1u;
            int n = word_base[w];
            while (m) {
                const int bit = __ffs(m) - 1;
                m &= m - 1;
                if (n >= n0 && n < n1) {
                    my_list[n - n0] = (w << 5) + bit;
                    ca[n] = (w << 5) + bit;
                }
                ++n;
            }
        }
    }
    __syncthreads();
    // gather of the candidates' boxes (16 of the 340 bytes of each prediction row), one warp per group, two groups
    // in flight; the group's union box lets the IoU sweep skip groups a GT cannot overlap
    for (int gq0 = gq_lo + 2 * warp; gq0 < gq_hi; gq0 += 2 * (kPrepThreads / 32)) {
        float4 pb[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int n = (gq0 + u) * 32 + lane;
            ok[u] = gq0 + u < gq_hi && n < Nc;
            if (ok[u]) pb[u] = load_box(p.preds + ((size_t)b * p.A + my_list[n - n0]) * p.ch);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (gq0 + u >= gq_hi) break;  // warp-uniform
            const int n = (gq0 + u) * 32 + lane;
            float4 c = make_float4(3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f);
            if (ok[u]) {
                // the candidate-side operands of bboxes_iou(xyxy=False) (iou_loss.py:400-410)
                c = make_float4(pb[u].x - pb[u].z / 2, pb[u].y - pb[u].w / 2, pb[u].x + pb[u].z / 2, pb[u].y + pb[u].w / 2);
                cb[n] = c;
                car[n] = pb[u].z * pb[u].w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c.x = fminf(c.x, __shfl_xor_sync(0xffffffffu, c.x, o));
                c.y = fminf(c.y, __shfl_xor_sync(0xffffffffu, c.y, o));
                c.z = fmaxf(c.z, __shfl_xor_sync(0xffffffffu, c.z, o));
                c.w = fmaxf(c.w, __shfl_xor_sync(0xffffffffu, c.w, o));
            }
            if (lane == 0) gb[gq0 + u] = c;
        }
    }
    if (q == 0 && tid == 0) {
        p.meta[b * 8 + 0] = G;
        p.meta[b * 8 + 1] = Nc;
        p.meta[b * 8 + 3] = 0;
        p.meta[b * 8 + 4] = 0;
        p.meta[b * 8 + 5] = 0;
        p.num_gt[b] = G;
        p.num_fg[b] = 0;
    }
}

// cost of (GT, anchor row) exactly as yolox_loss.py:84-108; warp-cooperative, valid in lane 0; *iou_out = IoU
__device__ __forceinline__ float pair_cost_row(const SimParams &p, const RawRow &r, const float gx, const float gy,
                                               const float gw, const float gh, const int gc, const bool both,
                                               const bool wide, const int lane, float *terms, float *iou_out) {
    LaneTerms t;
    lane_terms(r, p.C, lane, t);
    const float lcls = cls_cost(t, p.C, gc, lane, wide, terms);
    const float iou = pair_iou(gx, gy, gw, gh, r.box);
    const float liou = -logf(iou + 1e-8f);                          // :86
    *iou_out = iou;
    return (lcls + 3.0f * liou) + (both ? 0.0f : 100000.0f);        // :104-108
}
__device__ __forceinline__ float pair_cost(const SimParams &p, const int b, const int a, const float gx,
                                           const float gy, const float gw, const float gh, const int gc,
                                           const bool both, const bool wide, const int lane, float *terms,
                                           float *iou_out) {
    RawRow r;
    load_row(p.preds + ((size_t)b * p.A + a) * p.ch, p.C, lane, r);
    return pair_cost_row(p, r, gx, gy, gw, gh, gc, both, wide, lane, terms, iou_out);
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

// anchor a was claimed by several GTs: argmin of the cost column over ALL GT rows, first minimum (:352-356).
// Warp-cooperative.  The class costs of the anchor for every possible GT class come from one reduction in
// ATen's order: a butterfly over the all-negative leaves leaves each lane the sibling sums of its path, and
// swapping in the positive leaf re-adds them (C >= 32); smaller class counts evaluate every pair.
__device__ void resolve_conflict(const SimParams &p, const int b, const int a, const int G, const bool wide, float *T) {
    const int lane = threadIdx.x & 31;
    const float *L = p.labels + (size_t)b * p.Lmax * 5;
    const float *row = p.preds + ((size_t)b * p.A + a) * p.ch;
    const float4 pb = load_box(row);
    int l, x, y;
    anchor_cell(p, a, l, x, y);
    const bool fast = p.C >= 32;
    if (fast) {
        RawRow rr;
        load_row(row, p.C, lane, rr);
        LaneTerms t;
        lane_terms(rr, p.C, lane, t);
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
            float iou;
            const float cost = pair_cost(p, b, a, gr[1], gr[2], gr[3], gr[4], (int)gr[0], in_both(r8, x, y), wide, lane, T, &iou);
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

#ifndef PLYOLO_SWEEP_OCC
#define PLYOLO_SWEEP_OCC 4  // resident sweep CTAs per SM the register allocation is sized for (5 / 6: 48 / 40 registers with
                            // spills, measured 89.0 / 90.9 us per SimOTA call against 88.3: no gain)
#endif
// ---- K2a: IoU sweep ---------------------------------------------------------------------------
// kSweepSub warps per GT (kSweepGts GTs of one image per CTA).  Every warp tests the union boxes of the image's
// candidate groups 32 at a time and visits its share of the groups the GT overlaps (all other pairs have
// IoU 0; bboxes_iou with xyxy=False, iou_loss.py:400-414).  Every lane keeps the 4 largest IoUs of ITS
// candidates (branch-free insertion, no shuffles in the loop; the division only runs when a pair can enter
// the lane's list); the 10 largest of the image (:336-340, values only) are then extracted from the lane lists.
// That is exact unless a lane that saw more than 4 overlapping candidates still holds 4 values above the
// 10th largest — then (rare) the GT is redone with the exact warp-wide list.  dynamic k = clamp(int(sum of the
// top min(10, Nc)), 1) with ATen's reduce tree goes straight to the match kernel.
// Exact warp-wide list of the largest IoUs above `floor_v` (lane i = i-th largest; the list starts filled with
// floor_v, so fewer than 10 larger values leave copies of floor_v behind them).  floor_v = 0 is the plain top-10.
__device__ __noinline__ float sweep_exact(const float4 *cb, const float *car, const float4 *gb, const int Nc,
                                          const float g_x1, const float g_y1, const float g_x2, const float g_y2,
                                          const float area_a, const float floor_v) {
    const int lane = threadIdx.x & 31;
    const int ngrp = (Nc + 31) >> 5;
    // a pair that does not overlap has IoU (+/-)0 and can never displace anything
    float top = floor_v, thresh = floor_v;
    for (int q0 = 0; q0 < ngrp; q0 += 32) {
        bool h = false;
        if (q0 + lane < ngrp) {
            const float4 u = __ldg(gb + q0 + lane);
            h = fmaxf(g_x1, u.x) < fminf(g_x2, u.z) && fmaxf(g_y1, u.y) < fminf(g_y2, u.w);
        }
        unsigned hit = __ballot_sync(0xffffffffu, h);
        while (hit) {  // eight groups per trip: their loads are in flight together
            constexpr int U = 8;
            bool ok[U];
            float4 cc[U];
            float ar[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int n = (q0 + (hit ? __ffs(hit) - 1 : 0)) * 32 + lane;
                ok[u] = hit != 0u && n < Nc;
                hit &= hit - 1;  // 0 stays 0
                cc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                ar[u] = 0.f;
                if (ok[u]) { cc[u] = __ldg(cb + n); ar[u] = __ldg(car + n); }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float v = 0.f;
                const float tlx = fmaxf(g_x1, cc[u].x), tly = fmaxf(g_y1, cc[u].y);
                const float brx = fminf(g_x2, cc[u].z), bry = fminf(g_y2, cc[u].w);
                if (ok[u] && tlx < brx && tly < bry) {  // en == 1
                    const float area_i = (brx - tlx) * (bry - tly);
                    const float den = (area_a + ar[u]) - area_i;
                    if (area_i >= (thresh * den) * 0.999f) v = area_i / den;
                }
                unsigned m = __ballot_sync(0xffffffffu, v > thresh);
                if (m) {
                    while (m) {
                        const int j = __ffs(m) - 1;
                        m &= m - 1;
                        top_insert(top, __shfl_sync(0xffffffffu, v, j), lane);
                    }
                    thresh = __shfl_sync(0xffffffffu, top, 9);
                }
            }
        }
    }
    return top;
}

// barrier of the kSweepSub warps that share GT `gl` of the CTA
__device__ __forceinline__ void gt_barrier(const int gl) {
    static_assert(kSweepGts == 2, "one named barrier per GT of the CTA");
    if (gl == 0) asm volatile("bar.sync 1, %0;" ::"n"(kSweepSub * 32) : "memory");
    else asm volatile("bar.sync 2, %0;" ::"n"(kSweepSub * 32) : "memory");
}

// branch-free insertion of a (a >= 0, or 0 for "nothing") into the descending list t[0] >= ... >= t[kLaneTop-1]
constexpr int kLaneTop = 6;  // values kept per lane
__device__ __forceinline__ void lane_insert(float (&t)[kLaneTop], float a) {
#pragma unroll
    for (int i = 0; i < kLaneTop - 1; ++i) {
        const float hi = fmaxf(t[i], a);
        a = fminf(t[i], a);
        t[i] = hi;
    }
    t[kLaneTop - 1] = fmaxf(t[kLaneTop - 1], a);
}

__global__ void __launch_bounds__(kSweepThreads, PLYOLO_SWEEP_OCC) simota_sweep_kernel(const SimParams p) {
    __shared__ float s_list[kSweepGts][kSweepSub - 1][kLaneTop][32];
    __shared__ int s_seen[kSweepGts][kSweepSub - 1][32];
    __shared__ unsigned short s_hit[kSweepGts][kSweepChunk];
    __shared__ int s_nhit[kSweepGts];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // programmatic dependent launch: everything below reads the prep kernel's results; once they are complete the match
    // grid may start too — its first phases (in-both lists, row prefetch, exact terms, tight bounds) only need prep's
    // results, which are then guaranteed visible, and run while this grid drains
    pdl_wait();
    pdl_launch_dependents();
    const int gl = warp / kSweepSub, ws = warp % kSweepSub;  // GT of the CTA, part of the GT's hit groups
    const int g = blockIdx.x * kSweepGts + gl;
    const int G = p.meta[b * 8 + 0], Nc = p.meta[b * 8 + 1];
    if (g >= G || Nc == 0) return;  // uniform for the kSweepSub warps of a GT
    const float *L = p.labels + ((size_t)b * p.Lmax + g) * 5;
    const float gx = L[1], gy = L[2], gw = L[3], gh = L[4];
    const float g_x1 = gx - gw / 2, g_y1 = gy - gh / 2, g_x2 = gx + gw / 2, g_y2 = gy + gh / 2;
    const float area_a = gw * gh;
    const float4 *cb = p.cand_box + (size_t)b * p.A;
    const float *car = p.cand_area + (size_t)b * p.A;
    const float4 *gb = p.grp_box + (size_t)b * (p.A / 32 + 1);
    const int ngrp = (Nc + 31) >> 5;

    float t[kLaneTop];  // the lane's largest IoUs, descending (lists start at +0)
#pragma unroll
    for (int i = 0; i < kLaneTop; ++i) t[i] = 0.f;
    int seen = 0;                                  // overlapping candidates of this lane
    auto visit = [&](const bool valid, const float4 cc, const float ar) {
        const float tlx = fmaxf(g_x1, cc.x), tly = fmaxf(g_y1, cc.y);
        const float brx = fminf(g_x2, cc.z), bry = fminf(g_y2, cc.w);
        if (valid && tlx < brx && tly < bry) {  // en == 1
            ++seen;
            const float area_i = (brx - tlx) * (bry - tly);
            const float den = (area_a + ar) - area_i;
            if (area_i >= (t[kLaneTop - 1] * den) * 0.999f) {
                const float q = area_i / den;
                lane_insert(t, q > t[kLaneTop - 1] ? q : 0.f);  // NaN / not above the lane's last: no-op
            }
        }
    };
    // Chunks of kSweepChunk groups: the GT's warps list the groups whose union box the GT overlaps (any order:
    // only the IoU values matter), then walk the list together, 4 groups per warp and trip in flight.
    for (int c0 = 0; c0 < ngrp; c0 += kSweepChunk) {
        if (threadIdx.x % (kSweepSub * 32) == 0) s_nhit[gl] = 0;
        gt_barrier(gl);
        const int c1 = min(ngrp, c0 + kSweepChunk);
        for (int q0 = c0 + ws * 32; q0 < c1; q0 += kSweepSub * 32) {
            bool h = false;
            if (q0 + lane < c1) {
                const float4 u = __ldg(gb + q0 + lane);
                h = fmaxf(g_x1, u.x) < fminf(g_x2, u.z) && fmaxf(g_y1, u.y) < fminf(g_y2, u.w);
            }
            const unsigned hit = __ballot_sync(0xffffffffu, h);
            if (hit) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_nhit[gl], __popc(hit));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (h) s_hit[gl][base + __popc(hit & ((1u << lane) - 1u))] = (unsigned short)(q0 + lane - c0);
            }
        }
        gt_barrier(gl);
        const int nhit = s_nhit[gl];
        for (int i0 = ws * 4; i0 < nhit; i0 += kSweepSub * 4) {
            bool ok[4];
            float4 cc[4];
            float ar[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ok[u] = i0 + u < nhit;
                const int n = (c0 + (ok[u] ? (int)s_hit[gl][i0 + u] : 0)) * 32 + lane;
                ok[u] = ok[u] && n < Nc;
                cc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                ar[u] = 0.f;
                if (ok[u]) { cc[u] = __ldg(cb + n); ar[u] = __ldg(car + n); }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u >= nhit) break;  // warp-uniform
                visit(ok[u], cc[u], ar[u]);
            }
        }
        if (c1 < ngrp) gt_barrier(gl);  // the list is rebuilt by the next pass
    }
    // ---- the GT's first warp collects the other warps' lane lists (same lane = same candidates modulo 32)
    if (ws > 0) {
#pragma unroll
        for (int i = 0; i < kLaneTop; ++i) s_list[gl][ws - 1][i][lane] = t[i];
        s_seen[gl][ws - 1][lane] = seen;
    }
    gt_barrier(gl);
    if (ws > 0) return;
#pragma unroll
    for (int w = 0; w < kSweepSub - 1; ++w) {
#pragma unroll
        for (int u = 0; u < kLaneTop; ++u) lane_insert(t, s_list[gl][w][u][lane]);
        seen += s_seen[gl][w][lane];
    }
    // ---- the 10 largest over all lanes: ten rounds of (warp max, pop it from the first lane holding it)
    const float t_last = t[kLaneTop - 1];
    float top = 0.f;  // lane i < 10: i-th largest
    float m = 0.f;
    for (int r = 0; r < 10; ++r) {
        m = t[0];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == r) top = m;
        const unsigned who = __ballot_sync(0xffffffffu, t[0] == m);
        if (lane == __ffs(who) - 1) {
#pragma unroll
            for (int i = 0; i < kLaneTop - 1; ++i) t[i] = t[i + 1];
            t[kLaneTop - 1] = -1.f;
        }
    }
    // A lane that dropped candidates (saw more than it keeps) and whose last value still beats the 10th may have
    // dropped a top-10 value.  Every candidate value above m (the 10th largest found) belongs to the true top 10,
    // and at least 10 values are >= m: the exact list of the values above m, padded with m, is the answer.
    const bool unsafe = seen > kLaneTop && t_last > m;
    if (p.force_exact) top = sweep_exact(cb, car, gb, Nc, g_x1, g_y1, g_x2, g_y2, area_a, 0.f);
    else if (__any_sync(0xffffffffu, unsafe)) top = sweep_exact(cb, car, gb, Nc, g_x1, g_y1, g_x2, g_y2, area_a, m);
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
    if (lane == 0) p.dyn_k[(size_t)b * p.Lmax + g] = max((int)v, 1);
}

// ---- K2 ------------------------------------------------------------------------------------
struct MatchShared {
    int anchor[kGtPerCta][kMaxBoth];
    float cost[kGtPerCta][kMaxBoth];
    float iou[kGtPerCta][kMaxBoth];
    float liou3[kGtPerCta][kMaxBoth];  // 3 * L_iou of the pair (exact)
    float pos[kGtPerCta][kMaxBoth];    // positive BCE leaf of the pair (exact)
    float so[kGtPerCta][kMaxBoth];     // 1 / sigmoid(obj) of the pair's anchor
    float lso[kGtPerCta][kMaxBoth];    // log2 of it
    float terms[kMatchWarps][96];
    int conf[kMaxConf];
    int k[kGtPerCta], nb[kGtPerCta], gcls[kGtPerCta], gidx[kGtPerCta];
    int nconf, last, n_eval, n_cand;
    unsigned ucut[kGtPerCta];                      // ordered(U): the k-th smallest exact cost after the first round
    unsigned short clist[kGtPerCta * kMaxBoth];    // pairs whose cheap lower bound does not exceed U
    unsigned short elist[kGtPerCta * kMaxBoth];    // (GT, anchor slot) pairs whose exact cost is wanted
    unsigned char state[kGtPerCta][kMaxBoth];      // 0 = lower bound only, 1 = queued, 2 = exact cost known
    // tiny GTs (fewer in-both anchors than k): CTA-wide search over the candidates whose cost carries +1e5
    int tiny[kGtPerCta], ntiny, tn;
    unsigned long long wmin[kMatchWarps];
    unsigned long long tpick[16];                  // the `need` smallest lower-bound keys, then their exact cost keys
    unsigned short tlist[kTinyWindow];             // candidates of the window whose lower bound does not exceed U
    unsigned long long tkey[kTinyWindow];          // their exact (cost, anchor) keys
};

__global__ void __launch_bounds__(kMatchThreads) simota_match_kernel(const SimParams p) {
    extern __shared__ __align__(16) unsigned char match_smem[];
    MatchShared &sh = *reinterpret_cast<MatchShared *>(match_smem);
    const int b = blockIdx.y, g0 = blockIdx.x * kGtPerCta;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = p.meta[b * 8 + 0], Nc = p.meta[b * 8 + 1];
    if (g0 >= G || Nc == 0) return;
    const int n_active = (G + kGtPerCta - 1) / kGtPerCta;  // CTAs of this image that reach the end
    const int gi = warp & (kGtPerCta - 1), half = warp / kGtPerCta;  // IoU sweep: GT and candidate half of this warp
    const int *perm = p.gt_perm + (size_t)b * (p.Lmax + kGtPerCta) + g0;  // this CTA's GTs (balanced by prep)
    const int g = __ldg(perm + gi);
    const bool live = g >= 0;
    const float *L = p.labels + ((size_t)b * p.Lmax + (live ? g : 0)) * 5;
    const int gc = (int)L[0];  // .to(int64) truncates (:89)
    const float gx = L[1], gy = L[2], gw = L[3], gh = L[4];
    const int *ca = p.cand_anchor + (size_t)b * p.A;
    // ATen picks a 64-wide block for sum(-1) when the [G,Nc] output has fewer than 16 elements (and C >= 64)
    const bool wide = (long long)G * Nc < 16 && p.C >= 64;
    if (tid == 0) { sh.nconf = 0; sh.ntiny = 0; }
    long long *prof = p.prof ? p.prof + ((size_t)b * gridDim.x + blockIdx.x) * 16 : nullptr;
#define SPROF(slot) do { if (prof) { __syncthreads(); if (tid == 0) prof[slot] = clock64(); } } while (0)
    SPROF(0);

    // ---- 0. anchors both in-box and in-centre of every GT of the CTA (ascending anchor order; closed form from
    // the rectangles) and an L2 prefetch of their prediction rows: the IoU sweep below hides the DRAM latency
    if (half == 0) {
        int nb = 0;
        if (live) {
            const short *rect = p.rect + ((size_t)b * p.Lmax + g) * p.n_levels * 8;
            for (int l = 0; l < p.n_levels; ++l) {
                const short *r8 = rect + l * 8;
                const int x0 = max(r8[0], r8[4]), x1 = min(r8[1], r8[5]);
                const int y0 = max(r8[2], r8[6]), y1 = min(r8[3], r8[7]);
                if (x0 > x1 || y0 > y1) continue;
                const int wx = x1 - x0 + 1, cells = wx * (y1 - y0 + 1);
                for (int i = lane; i < cells; i += 32) {
                    if (nb + i < kMaxBoth) {
                        const int a = p.off[l] + (y0 + i / wx) * p.w[l] + x0 + i % wx;
                        sh.anchor[gi][nb + i] = a;
                        const char *row = reinterpret_cast<const char *>(p.preds + ((size_t)b * p.A + a) * p.ch);
                        if (kPrefetchRows >= 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
                        if (kPrefetchRows >= 2) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 128));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 256));
                            if (p.ch * 4 > 384 - 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + p.ch * 4 - 4));
                        }
                    }
                }
                nb = min(nb + cells, kMaxBoth);
            }
        }
        if (lane == 0) sh.nb[gi] = nb;
    }

    SPROF(1);
    // ---- 1. the GT's class / row for the pair loops; its dynamic k (:336-340) comes from the IoU sweep kernel, which may
    // still be running (programmatic dependent launch): it is read after the bounds below, behind griddepcontrol.wait
    SPROF(2);
    if (half == 0 && lane == 0) {
        sh.gcls[gi] = gc;
        sh.gidx[gi] = live ? g : 0;
    }
    __syncthreads();

    // ---- 2. cost of the in-both anchors (:84-108).  Only the k smallest costs of a GT matter, and
    //   cost = fl(fl(cls + 3 L_iou) + 0),  cls = ATen-ordered sum of 80 non-negative BCE leaves, one of them
    //   pos = -max(log p[gt class], -100):  every fp32 addition of non-negatives is monotone, so
    //   lb = fl(fl(pos + S + 3 L_iou) + 0) <= cost for any S <= the sum of the other 79 leaves.
    // (a) lb for every pair with S from fast-math leaves minus their error bound (tight: ~1e-3);  (b) exact cost — one warp per pair, the
    // expensive 80-class sweep — for the k pairs of smallest lb per GT;  (c) with U = the largest of those
    // exact costs, exact cost for every pair with lb <= U: all others cost more than k pairs already do.
    // The selection (3.) then runs over the exactly known costs only.
    SPROF(3);
    int pbase[kGtPerCta + 1];
    pbase[0] = 0;
#pragma unroll
    for (int q = 0; q < kGtPerCta; ++q) pbase[q + 1] = pbase[q] + sh.nb[q];
    const int npairs = pbase[kGtPerCta];
    // (a1) one thread per pair: the exact terms (positive leaf, IoU, 3 L_iou) from 6 floats of the row
    for (int t = tid; t < npairs; t += kMatchThreads) {
        int q = 0, base = 0;
#pragma unroll
        for (int u = 1; u < kGtPerCta; ++u)
            if (t >= pbase[u]) { q = u; base = pbase[u]; }
        const int i = t - base;
        const float *row = p.preds + ((size_t)b * p.A + sh.anchor[q][i]) * p.ch;
        const float *Lq = p.labels + ((size_t)b * p.Lmax + sh.gidx[q]) * 5;
        const int qc = (int)Lq[0];
        const float so = sigmoid_ref(__ldg(row + 4));
        float pos = 0.f;
        if (qc >= 0 && qc < p.C) pos = pos_term(sqrtf(sigmoid_ref(__ldg(row + 5 + qc)) * so));
        const float iou = pair_iou(Lq[1], Lq[2], Lq[3], Lq[4], load_box(row));
        const float liou = -logf(iou + 1e-8f);
        sh.pos[q][i] = pos;
        sh.liou3[q][i] = 3.0f * liou;
        sh.cost[q][i] = (pos + 3.0f * liou) + 0.0f;  // lb0: the negative leaves are >= 0
        sh.so[q][i] = 1.0f / so;
        sh.lso[q][i] = log2f(1.0f / so);
        sh.iou[q][i] = iou;
        sh.state[q][i] = 0;
        if (kTightAll) sh.clist[t] = (unsigned short)(q * kMaxBoth + i);
    }
    // Tight lower bound of the listed pairs sh.clist[0, ncand) — eight lanes per pair, four pairs per warp at a
    // time, every class logit of the row loaded up front:
    // S = guaranteed lower bound of the negative leaves' sum (leaf_p_approx + leaf_series minus kLeafErr per leaf),
    // lb = fl(fl((pos + S) (1 - 4e-6)) + 3 L_iou) — 4e-6 covers the fp32 roundings of the reference's 80-leaf
    // tree sum (<= 8 half-ulps) and of this accumulation.  filter: pairs with lb <= U are queued for the exact
    // evaluation, the others leave the selection; !filter: lb replaces the cheap bound.
    auto tight_pass = [&](const int ncand, const bool filter) {
        constexpr int NJ = (PLYOLO_MAX_CLASSES + 7) / 8;
        const int sub = lane & 7;
        for (int t0 = warp * 4; t0 < ncand; t0 += kMatchWarps * 4) {
            const int t = t0 + (lane >> 3);
            const bool valid = t < ncand;
            const int e = valid ? sh.clist[t] : 0;
            const int q = e / kMaxBoth, i = e % kMaxBoth;
            const float *row = p.preds + ((size_t)b * p.A + (valid ? sh.anchor[q][i] : 0)) * p.ch + 5;
            float x[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) x[j] = __ldg(row + min(sub + 8 * j, p.C - 1));  // unconditional: all in flight at once
            const int qc = sh.gcls[q];
            const float inv_so = sh.so[q][i], lso = sh.lso[q][i];
            float sneg = 0.f, pmax = 0.f;
            int nleaf = 0;
            float pr[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                pr[j] = 0.f;
                if (8 * j < p.C) {  // uniform
                    const int c = sub + 8 * j;
                    const bool use = c < p.C && c != qc;
                    const float v = leaf_p_approx(x[j], inv_so, lso);
                    pr[j] = use ? v : 0.f;  // p = 0: leaf 0 in either formula
                    nleaf += use ? 1 : 0;
                    pmax = fmaxf(pmax, pr[j]);
                }
            }
            if (!__any_sync(0xffffffffu, pmax > 0.3f)) {  // the usual case: every leaf by the series, no branches
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (8 * j < p.C) sneg += leaf_series(pr[j]);
            } else {
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    if (8 * j < p.C) sneg += leaf_general(pr[j]);
            }
            sneg = sneg - (float)nleaf * kLeafErr;  // every exact leaf is >= max(approximation - kLeafErr, 0)
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) sneg += __shfl_xor_sync(0xffffffffu, sneg, o);
            if (!(sneg >= 0.f)) sneg = 0.f;  // also NaN
            if (valid && sub == 0) {
                const float cl = (sh.pos[q][i] + sneg) * 0.999996f;
                const float lb = (cl + sh.liou3[q][i]) + 0.0f;
                if (!filter) {
                    sh.cost[q][i] = lb;
                } else if (float_ordered(lb) <= sh.ucut[q]) {
                    sh.state[q][i] = 1;
                    sh.elist[atomicAdd(&sh.n_eval, 1)] = (unsigned short)e;
                } else {
                    sh.cost[q][i] = __int_as_float(0x7fc00000);  // cannot be among the k smallest: out of the selection
                }
            }
        }
    };
    if (kTightAll) {
        __syncthreads();
        tight_pass(npairs, false);
    }
    pdl_wait();  // the IoU sweep has completed: dynamic k of every GT
    if (half == 0 && lane == 0) {
        int k = 0;
        if (live) {
            k = __ldcg(p.dyn_k + (size_t)b * p.Lmax + g);
            if (!(k < Nc - 1)) sh.nb[gi] = 0;  // quirk Q3: every candidate is taken, no cost needed
        }
        sh.k[gi] = k;
    }
    if (tid == 0) { sh.n_eval = 0; sh.n_cand = 0; }
    __syncthreads();
    // exact cost of every queued pair: the warps walk the list with stride 16, loading the NEXT pair's
    // prediction row (three coalesced requests) before evaluating the current one
    auto evaluate_queued = [&]() {
        const int ne = sh.n_eval;
        RawRow cur, nxt;
        int q = 0, i = 0;
        if (warp < ne) {
            q = sh.elist[warp] / kMaxBoth; i = sh.elist[warp] % kMaxBoth;
            load_row(p.preds + ((size_t)b * p.A + sh.anchor[q][i]) * p.ch, p.C, lane, cur);
        }
        for (int t = warp; t < ne; t += kMatchWarps) {
            int qn = 0, in_ = 0;
            const bool more = t + kMatchWarps < ne;
            if (more) {
                qn = sh.elist[t + kMatchWarps] / kMaxBoth; in_ = sh.elist[t + kMatchWarps] % kMaxBoth;
                load_row(p.preds + ((size_t)b * p.A + sh.anchor[qn][in_]) * p.ch, p.C, lane, nxt);
            }
            const float *Lq = p.labels + ((size_t)b * p.Lmax + sh.gidx[q]) * 5;
            float iou;
            const float c = pair_cost_row(p, cur, Lq[1], Lq[2], Lq[3], Lq[4], (int)Lq[0], true, wide, lane, sh.terms[warp], &iou);
            if (lane == 0) { sh.cost[q][i] = c; sh.state[q][i] = 2; }
            if (more) { cur = nxt; q = qn; i = in_; }
        }
    };
    SPROF(4);
    // (b) queue the k smallest lower bounds of every GT (ties -> lower slot, like the final selection)
    if (half == 0 && sh.nb[gi] > 0) {
        const int nb = sh.nb[gi], take = min(sh.k[gi] + kExtraEval, nb);  // a few more than k: a tighter U below
        for (int r = 0; r < take; ++r) {
            unsigned long long best = ~0ull;
            for (int i = lane; i < nb; i += 32)
                if (sh.state[gi][i] == 0) {
                    const unsigned long long key = ((unsigned long long)float_ordered(sh.cost[gi][i]) << 32) | (unsigned)i;
                    best = key < best ? key : best;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other < best ? other : best;
            }
            __syncwarp();  // every lane has read the states of this round
            if (lane == 0) {
                const int i = (int)(best & 0xffffffffu);
                sh.state[gi][i] = 1;
                sh.elist[atomicAdd(&sh.n_eval, 1)] = (unsigned short)(gi * kMaxBoth + i);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    SPROF(5);
    evaluate_queued();
    __syncthreads();
    SPROF(6);
    if (prof && tid == 0) prof[12] = sh.n_eval;
    if (tid == 0) sh.n_eval = 0;
    __syncthreads();
    // (c) everything whose lower bound does not exceed the largest exact cost found so far
    if (half == 0 && sh.nb[gi] > 0) {
        const int nb = sh.nb[gi], kk = min(sh.k[gi], nb);
        // U = the kk-th smallest exact cost known so far (>= the true kk-th smallest cost of the GT)
        unsigned umax = 0u, floor_ = 0u;
        bool first = true;
        for (int r = 0; r < kk; ++r) {
            unsigned long long best = ~0ull;  // smallest (cost, slot) above the previous pick
            for (int i = lane; i < nb; i += 32)
                if (sh.state[gi][i] == 2) {
                    const unsigned long long key = ((unsigned long long)float_ordered(sh.cost[gi][i]) << 32) | (unsigned)i;
                    const unsigned long long prev = ((unsigned long long)umax << 32) | floor_;
                    if ((first || key > prev) && key < best) best = key;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
                best = other < best ? other : best;
            }
            umax = (unsigned)(best >> 32);
            floor_ = (unsigned)(best & 0xffffffffu);
            first = false;
        }
        if (lane == 0) sh.ucut[gi] = umax;
        for (int i = lane; i < nb; i += 32)
            if (sh.state[gi][i] == 0) {
                if (float_ordered(sh.cost[gi][i]) <= umax) {
                    if (kTightAll) {
                        sh.state[gi][i] = 1;
                        sh.elist[atomicAdd(&sh.n_eval, 1)] = (unsigned short)(gi * kMaxBoth + i);
                    } else {
                        sh.clist[atomicAdd(&sh.n_cand, 1)] = (unsigned short)(gi * kMaxBoth + i);
                    }
                } else {
                    sh.cost[gi][i] = __int_as_float(0x7fc00000);  // cannot be among the k smallest: out of the selection
                }
            }
    }
    __syncthreads();
    if (!kTightAll) tight_pass(sh.n_cand, true);
    __syncthreads();
    SPROF(7);
    evaluate_queued();
    __syncthreads();
    SPROF(8);
    if (prof && tid == 0) { prof[13] = sh.n_eval; prof[14] = npairs; }

    // ---- 3. per GT (first 8 warps): k smallest (cost, anchor); ties -> lowest anchor index (stable sort, :342)
    auto push_conflict = [&](const int a) {
        const int i = atomicAdd(&sh.nconf, 1);
        if (i < kMaxConf) sh.conf[i] = a;
    };
    if (half == 0 && live) {
        const int k = sh.k[gi], nb = sh.nb[gi];
        if (!(k < Nc - 1)) {  // quirk Q3 (:343): the GT takes EVERY candidate (Nc <= 11 here, k <= 10)
            for (int n = lane; n < Nc; n += 32) {
                const float iou = pair_iou(gx, gy, gw, gh, load_box(p.preds + ((size_t)b * p.A + ca[n]) * p.ch));
                if (claim(p, b, ca[n], g, iou)) push_conflict(ca[n]);
            }
        } else {
            const int take = min(k, nb);
            for (int r = 0; r < take; ++r) {
                unsigned long long best = ~0ull;
                for (int i = lane; i < nb; i += 32) {
                    const float c = sh.cost[gi][i];
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
                    if (claim(p, b, sh.anchor[gi][i], g, sh.iou[gi][i])) push_conflict(sh.anchor[gi][i]);
                    sh.cost[gi][i] = __int_as_float(0x7fc00000);
                }
                __syncwarp();
            }
            // tiny GT: fewer in-both anchors than k — the remaining picks are searched by the whole CTA below
            if (k > nb && lane == 0) sh.tiny[atomicAdd(&sh.ntiny, 1)] = gi;
        }
    }

    // ---- 3b. tiny GTs: the remaining k - nb picks come from the candidates whose cost carries +1e5 (quantised to
    // 1/128, T8): the smallest (cost, anchor) over all candidates that are not in-both.  Whole CTA per GT:
    //   lb(n) = fl(3 L_iou + 1e5) <= cost(n) (the class cost is >= 0, fp32 addition is monotone) from the candidate's
    //   corner box alone;  the `need` smallest lb -> their exact costs -> U = the largest of them, an upper bound of
    //   the need-th smallest cost;  every candidate with lb <= U gets its exact 80-class cost (one warp each), the
    //   need smallest exact keys win.  Exact for any input; cheap because only boxes that overlap the GT about as
    //   well as the best ones can have lb <= U.
    __syncthreads();
    for (int job = 0; job < sh.ntiny; ++job) {
        const int q = sh.tiny[job], nbq = sh.nb[q], need = sh.k[q] - nbq, gq = sh.gidx[q];  // need <= 10
        const float *Lq = p.labels + ((size_t)b * p.Lmax + gq) * 5;
        const int qc = (int)Lq[0];
        const float qx = Lq[1], qy = Lq[2], qw = Lq[3], qh = Lq[4];
        const float q_x1 = qx - qw / 2, q_y1 = qy - qh / 2, q_x2 = qx + qw / 2, q_y2 = qy + qh / 2, q_area = qw * qh;
        const float4 *cb = p.cand_box + (size_t)b * p.A;
        const float *car = p.cand_area + (size_t)b * p.A;
        // lower-bound key of candidate n; false: the candidate is in-both (its cost was handled above)
        auto lb_key = [&](const int n, unsigned long long &key) -> bool {
            const int a = __ldg(ca + n);
            for (int i = 0; i < nbq; ++i)
                if (sh.anchor[q][i] == a) return false;
            const float4 c = __ldg(cb + n);  // the candidate-side operands of bboxes_iou, as the prep kernel formed them
            const float tlx = fmaxf(q_x1, c.x), tly = fmaxf(q_y1, c.y), brx = fminf(q_x2, c.z), bry = fminf(q_y2, c.w);
            const float en = (tlx < brx ? 1.f : 0.f) * (tly < bry ? 1.f : 0.f);
            const float area_i = ((brx - tlx) * (bry - tly)) * en;
            const float iou = area_i / ((q_area + __ldg(car + n)) - area_i);
            const float liou = -logf(iou + 1e-8f);
            const float lb = (0.0f + 3.0f * liou) + 100000.0f;
            key = ((unsigned long long)float_ordered(lb) << 32) | (unsigned)a;
            return true;
        };
        auto block_min = [&](unsigned long long v) -> unsigned long long {  // every thread gets the CTA-wide minimum
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
                v = other < v ? other : v;
            }
            if (lane == 0) sh.wmin[warp] = v;
            __syncthreads();
            v = sh.wmin[0];
#pragma unroll
            for (int w = 1; w < kMatchWarps; ++w) v = sh.wmin[w] < v ? sh.wmin[w] : v;
            __syncthreads();
            return v;
        };
        // (a) the `need` smallest lower-bound keys, one CTA-wide minimum per round
        unsigned long long prev = 0ull;
        for (int r = 0; r < need; ++r) {
            unsigned long long best = ~0ull;
            for (int n = tid; n < Nc; n += kMatchThreads) {
                unsigned long long key;
                if (lb_key(n, key) && (r == 0 || key > prev) && key < best) best = key;
            }
            prev = block_min(best);
            if (tid == 0) sh.tpick[r] = prev;
        }
        __syncthreads();
        // (b) their exact costs (one warp each) -> U
        if (warp < need && sh.tpick[warp] != ~0ull) {
            float iou;
            const float c = pair_cost(p, b, (int)(sh.tpick[warp] & 0xffffffffu), qx, qy, qw, qh, qc, false, wide, lane,
                                      sh.terms[warp], &iou);
            if (lane == 0) sh.tpick[warp] = (unsigned long long)float_ordered(c) << 32;
        }
        __syncthreads();
        unsigned ucut = 0u;
        for (int r = 0; r < need; ++r)
            if (sh.tpick[r] != ~0ull) ucut = max(ucut, (unsigned)(sh.tpick[r] >> 32));
        // (c) windows of the candidate list: exact cost of everything with lb <= U, warp 0 keeps the need smallest keys
        unsigned long long mine = ~0ull;  // warp 0: lanes 0..need-1 hold the smallest keys so far, ascending
        for (int w0 = 0; w0 < Nc; w0 += kTinyWindow) {
            if (tid == 0) sh.tn = 0;
            __syncthreads();
            for (int n = w0 + tid; n < min(Nc, w0 + kTinyWindow); n += kMatchThreads) {
                unsigned long long key;
                if (lb_key(n, key) && (unsigned)(key >> 32) <= ucut) sh.tlist[atomicAdd(&sh.tn, 1)] = (unsigned short)(n - w0);
            }
            __syncthreads();
            const int tn = sh.tn;
            for (int t = warp; t < tn; t += kMatchWarps) {
                const int a = __ldg(ca + w0 + sh.tlist[t]);
                float iou;
                const float c = pair_cost(p, b, a, qx, qy, qw, qh, qc, false, wide, lane, sh.terms[warp], &iou);
                if (lane == 0) sh.tkey[t] = ((unsigned long long)float_ordered(c) << 32) | (unsigned)a;
            }
            __syncthreads();
            if (warp == 0) {
                for (int t = 0; t < tn; ++t) {
                    const unsigned long long key = sh.tkey[t];
                    const unsigned long long upk = __shfl_up_sync(0xffffffffu, mine, 1);
                    if (key < mine) mine = (lane == 0 || upk <= key) ? key : upk;
                }
            }
        }
        if (warp == 0 && lane < need && mine != ~0ull) {
            const int a = (int)(mine & 0xffffffffu);
            const float iou = pair_iou(qx, qy, qw, qh, load_box(p.preds + ((size_t)b * p.A + a) * p.ch));
            if (claim(p, b, a, gq, iou)) push_conflict(a);
        }
        __syncthreads();
    }

    // ---- resolve the conflicts this CTA created (one warp each)
    __syncthreads();
    SPROF(9);
    const int nc = min(sh.nconf, kMaxConf);
    for (int i = warp; i < nc; i += kMatchWarps) resolve_conflict(p, b, sh.conf[i], G, wide, sh.terms[warp]);

    SPROF(10);
    // ---- 4. the last CTA of the image to finish patches the resolved matches over the tentative ones
    __threadfence();
    __syncthreads();
    if (tid == 0) sh.last = (atomicAdd(&p.meta[b * 8 + 4], 1) == n_active - 1) ? 1 : 0;
    __syncthreads();
    if (sh.last) {
        __threadfence();
        const int nconf = __ldcg(p.meta + b * 8 + 5);
        for (int i = tid; i < nconf; i += kMatchThreads) {
            const int a = __ldcg(p.conf_list + (size_t)b * p.A + i);
            p.matched_gt[(size_t)b * p.A + a] = __ldcg(p.res_g + (size_t)b * p.A + a);
            p.matched_iou[(size_t)b * p.A + a] = __ldcg(p.res_iou + (size_t)b * p.A + a);
        }
        if (tid == 0) p.num_fg[b] = __ldcg(p.meta + b * 8 + 3);  // :358
    }
    SPROF(11);
#undef SPROF
}

static thread_local long long *g_sim_prof = nullptr;
static thread_local int g_sim_force_exact = 0;

static size_t sim_ws_layout(int B, int A, int Lmax, int n_levels, SimParams *p, unsigned char *base) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_meta = take((size_t)B * 8 * sizeof(int));
    const size_t o_sl = take((size_t)B * A * sizeof(int));
    const size_t o_ri = take((size_t)B * A * sizeof(float));
    const size_t o_ca = take((size_t)B * A * sizeof(int));
    const size_t o_cb = take((size_t)B * A * sizeof(float4));
    const size_t o_car = take((size_t)B * A * sizeof(float));
    const size_t o_sc = take((size_t)B * A * sizeof(unsigned));
    const size_t o_sm = take((size_t)B * A * sizeof(unsigned));
    const size_t o_rect = take((size_t)B * Lmax * n_levels * 8 * sizeof(short));
    const size_t o_gb = take((size_t)B * (A / 32 + 1) * sizeof(float4));
    const size_t o_k = take((size_t)B * Lmax * sizeof(int));
    const size_t o_perm = take((size_t)B * (Lmax + kGtPerCta) * sizeof(int));
    if (p) {
        p->grp_box = reinterpret_cast<float4 *>(base + o_gb);
        p->dyn_k = reinterpret_cast<int *>(base + o_k);
        p->gt_perm = reinterpret_cast<int *>(base + o_perm);
        p->meta = reinterpret_cast<int *>(base + o_meta);
        p->conf_list = reinterpret_cast<int *>(base + o_sl);
        p->res_iou = reinterpret_cast<float *>(base + o_ri);
        p->cand_anchor = reinterpret_cast<int *>(base + o_ca);
        p->cand_box = reinterpret_cast<float4 *>(base + o_cb);
        p->cand_area = reinterpret_cast<float *>(base + o_car);
        p->sel_count = reinterpret_cast<unsigned *>(base + o_sc);
        p->res_g = reinterpret_cast<int *>(base + o_sm);
        p->rect = reinterpret_cast<short *>(base + o_rect);
    }
    return off;
}

}  // namespace plyolo

// debug hook (not part of include/plyolo.h): device buffer [B][ceil(Lmax/8)][16] of int64 for the match kernel's
// phase timestamps (adds barriers: timing only), null switches it off
extern "C" void plyolo_debug_simota_profile(void *device_buf) { plyolo::g_sim_prof = static_cast<long long *>(device_buf); }
// debug hook: the IoU sweep uses its exact warp-wide top-10 list for every GT (normally only when the lane lists
// cannot prove themselves exact) — lets the tests pin both routes
extern "C" void plyolo_debug_simota_force_exact(int on) { plyolo::g_sim_force_exact = on; }

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
    p.prof = g_sim_prof;
    p.force_exact = g_sim_force_exact;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bm = (2 * (size_t)((A + 31) / 32) + 1 + (size_t)(A / kPrepSplit + 128) + (size_t)Lmax) * sizeof(unsigned);
    PLYOLO_REQUIRE(bm <= 160 * 1024, "A=%d too large for the candidate bitmap", A);
    if (first_use_on_device(2)) {  // fixed maxima, once per device
        cudaFuncSetAttribute(simota_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(simota_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MatchShared));
    }
    record_stage_event(0, st);
    simota_prep_kernel<<<dim3(kPrepSplit, B), kPrepThreads, bm, st>>>(p);
    PLYOLO_CHECK_LAUNCH("simota_prep_kernel");
    const bool pdl = pdl_enabled();
    if (launch_ex(simota_sweep_kernel, dim3((Lmax + kSweepGts - 1) / kSweepGts, B), dim3(kSweepThreads), 0, st, pdl, p) != cudaSuccess) {
        set_error("simota_sweep_kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    count_launch();
    record_stage_event(1, st);
    if (launch_ex(simota_match_kernel, dim3((Lmax + kGtPerCta - 1) / kGtPerCta, B), dim3(kMatchThreads), sizeof(MatchShared), st, pdl,
                  p) != cudaSuccess) {
        set_error("simota_match_kernel: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    count_launch();
    record_stage_event(2, st);
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
