// nms.cuh — class-aware NMS of one image by one CTA (torchvision.ops.batched_nms / nms semantics:
// tv:ops/boxes.py:51-120, torchvision::nms) + the tail of postprocess (postprocess.py:43-46).
// Included by postprocess.cu after CandWs / the tile constants are defined.
//
// Input: the score stage's candidates of the image, per tile of 128 anchors in anchor order
// (CandWs).  Output: the first max_det kept boxes in (score desc, anchor asc) order.
//
// FAST PATH (class-aware, <= kFastCap candidates).  The greedy sweep only couples boxes whose IoU
// test fires.  With the coordinate trick (boxes + class * (max_coord + 1)) boxes of different classes
// can only meet when one of them has x1 and y1 below -1, so unless such a pair actually suppresses
// (checked exactly, below) the sweep decomposes by class:
//   1. one pass over the candidates: 64-bit keys (class | ~score | slot), class histogram, max
//      coordinate, the list of "cross" boxes (x1 < -0.5 and y1 < -0.5);
//   2. counting scatter of the keys into class segments (order inside a segment is irrelevant);
//   3. exact cross-class check: far-corner candidates against the cross boxes, torchvision's arithmetic;
//      a hit sends the image to the general path;
//   4. one warp per class (largest first, dynamic): bitonic sort of the segment (registers for <= 32
//      keys, shared memory above), gather + offset the boxes, greedy sweep in chunks of 32 against the
//      class's kept list (early exit when the whole chunk is dead);
//   5. kept keys (class stripped) compacted, block-sorted, first max_det gathered to the output.
// GENERAL PATH (class-agnostic, > kFastCap candidates, interleaving class ranges, or a cross-class
// hit): block-wide bitonic sort of all keys, then greedy rounds of 256 candidates against the kept
// list with a suppression bit-matrix among the round's survivors; exits at max_det keeps.
#pragma once

#include <cfloat>

#include <cooperative_groups.h>

#include "common.cuh"

namespace plyolo {

namespace cg = cooperative_groups;

// torchvision's IoU test; a = kept (higher-scored, "row") box, b = later ("column") box.
// Exactly `inter / union > thr` with torchvision's roundings, but the IEEE division only runs inside a
// +-1e-6 relative band around the threshold: outside it the correctly rounded quotient provably lies
// on the same side as the (cheap) product test.  No overlap -> quotient 0 -> never above thr >= 0.
__device__ __forceinline__ bool suppresses(const float4 a, const float4 b, const int flavor, const float thr_f,
                                           const double thr_d) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    const float inter = w * h;
    if (!(inter > 0.f) && thr_f >= 0.f) return false;
    const float Sa = (a.z - a.x) * (a.w - a.y);
    float u;
    if (!(flavor & PLYOLO_IOU_NOFMA)) {
        // torchvision 0.26 nms_kernel.cu as compiled for sm_100: Sb is contracted into the sum
        // (0 / 20000 near-threshold pairs differ on B200; the un-fused form flips 352 of them)
        u = __fmaf_rn(b.z - b.x, b.w - b.y, Sa) - inter;
    } else {
        const float Sb = (b.z - b.x) * (b.w - b.y);
        u = (Sa + Sb) - inter;
    }
    if (u > 0.f && thr_f > 0.f && u < 1e30f && inter > 1e-30f) {
        const float cut = thr_f * u;
        if (inter > cut * 1.000001f) return true;
        if (inter < cut * 0.999999f) return false;
    }
    const float iou = inter / u;
    return (flavor & PLYOLO_THR_F64) ? ((double)iou > thr_d) : (iou > thr_f);
}

// inverse of float_ordered
__device__ __forceinline__ float ordered_float(const uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kMaxSortCap = 16384;
constexpr int kRound = 256;                 // candidates per round of the general path
constexpr int kSub = kNmsThreads / kRound;  // threads per candidate
constexpr int kCrossBit = 0x100;            // class word flag: this box must be tested against every class
constexpr int kMaxClasses = 128;            // class field of the key: 7 bits

// ---- sort keys ----------------------------------------------------------------------------------
// bits [63:57] class (0 in the class-less keys of the general path), [56:25] ~ordered(score),
// [24:0] candidate slot (tile * 128 + position: ascending slot == anchor order).  Ascending key order
// is (class,) score descending, ties -> lower slot == what a stable descending sort by score yields.
constexpr int kSlotBits = 25;
constexpr unsigned long long kSlotMask = (1ull << kSlotBits) - 1ull;
constexpr unsigned long long kOrderMask = (1ull << 57) - 1ull;  // score + slot: the global order

__device__ __forceinline__ int key_slot(const unsigned long long k) { return (int)(k & kSlotMask); }
__device__ __forceinline__ int key_class(const unsigned long long k) { return (int)(k >> 57); }

// Dynamic shared memory of nms_image (bytes, 16-byte aligned pieces):
//   general path: keys[sort_cap] u64 | kept list + round buffers (nms_general_bytes)
//   fast path   : keys[fast_cap] u64 | kept_box[fast_cap] float4 (first the staging keys) | box_r[fast_cap] float4
//   then        : pref[NT + 1] int | keep_bits[fast_cap / 32 + 1] u32
__host__ __device__ inline size_t nms_general_bytes(const int max_det) {
    return ((size_t)max_det * 24 + (size_t)kRound * 24 + (size_t)kRound * (kRound / 32) * 4 + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t nms_core_bytes(const int sort_cap, const int fast_cap, const int max_det) {
    const size_t gen_b = (size_t)sort_cap * 8 + nms_general_bytes(max_det), fast_b = (size_t)fast_cap * 40;
    return gen_b > fast_b ? gen_b : fast_b;
}
__host__ __device__ inline size_t nms_smem_bytes(const int sort_cap, const int fast_cap, const int max_det, const int NT) {
    return nms_core_bytes(sort_cap, fast_cap, max_det) + (size_t)(NT + 1) * sizeof(int) +
           (size_t)((fast_cap >> 5) + 1) * sizeof(unsigned);
}

struct NmsParams {
    int B, NT, max_nms, max_det, flavor, agnostic, sort_cap, fast_cap;
    int variant;      // PLYOLO_NMS_*: 0 = postprocess.py / batched_nms; YOLOv3 / YOLOv5 decoder call sites otherwise (general path only)
    float fixed_span; // > 0: class offset = class * fixed_span (yolov5_decoder.py:70) instead of max_coordinate + 1
    int only_image;   // nms_general_kernel launched from the device for ONE image (>= 0), else -1: one CTA per image
    unsigned general_smem;  // dynamic shared memory of nms_general_kernel (for the device-side launch)
    int all_general;  // nms_general_kernel: every image takes the general path (no class-split kernel ran)
    int device_launch;  // the class-split kernel launches the general path itself (0: the host launched it behind it)
    int wait_tiles;   // nms_fast_kernel: spin on the image's scored-tile counter (the score kernel may still be running)
    float thr_f;
    double thr_d;
    CandWs ws;
    float *dets;
    int32_t *counts;
    int32_t *keep_idx;
    // multi-GPU evaluation: the image's rows and count are ALSO stored straight into the other ranks' gathered
    // buffers (peer memory over NVLink / NVSwitch) — the detection all-gather happens inside the NMS kernels
    int n_peers;
    float *peer_dets[PLYOLO_MAX_PEERS];      // each already offset to this rank's block [B, max_det, 6]
    int32_t *peer_counts[PLYOLO_MAX_PEERS];  // each already offset to this rank's block [B]
    long long *prof;  // debug: [B][16] phase timestamps (clock64) or null
};

// one output row / the image's count, to the caller's buffers and to every peer's
__device__ __forceinline__ void store_row6(const NmsParams &p, const int b, const int row, const float2 r0, const float2 r1,
                                           const float2 r2) {
    const size_t o = ((size_t)b * p.max_det + row) * 6;
    float2 *d = reinterpret_cast<float2 *>(p.dets + o);
    d[0] = r0; d[1] = r1; d[2] = r2;
    for (int r = 0; r < p.n_peers; ++r) {
        float2 *q = reinterpret_cast<float2 *>(p.peer_dets[r] + o);
        q[0] = r0; q[1] = r1; q[2] = r2;
    }
}
// the caller's buffers only (the class-split kernel ships the finished image block to the peers in one coalesced pass)
__device__ __forceinline__ void store_row6_local(const NmsParams &p, const int b, const int row, const float2 r0, const float2 r1,
                                                 const float2 r2) {
    float2 *d = reinterpret_cast<float2 *>(p.dets + ((size_t)b * p.max_det + row) * 6);
    d[0] = r0; d[1] = r1; d[2] = r2;
}
__device__ __forceinline__ void store_count(const NmsParams &p, const int b, const int n) {
    p.counts[b] = n;
    for (int r = 0; r < p.n_peers; ++r) p.peer_counts[r][b] = n;
}

#define NMS_PROF(slot)                                                                   \
    do {                                                                                 \
        if (p.prof && threadIdx.x == 0) p.prof[(size_t)b * 16 + (slot)] = clock64();     \
    } while (0)

// Block-wide bitonic sort, ascending, n_pad = power of two >= 64.  Steps with partner distance <= 32
// run in registers (each warp owns 64 consecutive keys, two per lane, exchanged by shuffles); only
// distances >= 64 go through shared memory with a block barrier.
__device__ void block_sort(unsigned long long *keys, const int n_pad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    auto local_steps = [&](const int k_lo, const int k_hi, const int j_hi) {
        for (int chunk = warp; chunk < (n_pad >> 6); chunk += kNmsWarps) {
            const int i0 = (chunk << 6) + lane;
            unsigned long long a = keys[i0], c = keys[i0 + 32];
            for (int k = k_lo; k <= k_hi; k <<= 1) {
                const bool up = (i0 & k) == 0;
                for (int j = min(k >> 1, j_hi); j > 0; j >>= 1) {
                    if (j == 32) {
                        if ((a > c) == up) { const unsigned long long t = a; a = c; c = t; }
                    } else {
                        const bool lower = (lane & j) == 0;
                        const bool upa = (i0 & k) == 0, upc = ((i0 + 32) & k) == 0;
                        const unsigned long long oa = __shfl_xor_sync(0xffffffffu, a, j);
                        const unsigned long long oc = __shfl_xor_sync(0xffffffffu, c, j);
                        a = (lower == upa) ? (a < oa ? a : oa) : (a > oa ? a : oa);
                        c = (lower == upc) ? (c < oc ? c : oc) : (c > oc ? c : oc);
                    }
                }
            }
            keys[i0] = a;
            keys[i0 + 32] = c;
        }
    };
    local_steps(2, 64, 32);
    __syncthreads();
    for (int k = 128; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j >= 64; j >>= 1) {
            for (int i = tid; i < n_pad; i += kNmsThreads) {
                const int q = i ^ j;
                if (q > i) {
                    const unsigned long long x = keys[i], y = keys[q];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[q] = x; }
                }
            }
            __syncthreads();
        }
        local_steps(k, k, 32);
        __syncthreads();
    }
}

// ---- warp-level sort of one class segment ---------------------------------------------------------
// Bitonic network whose merges all run ascending (the first step of a merge mirrors the upper half:
// partner = e ^ (k - 1), the rest are e ^ j), so +inf padding up to a power of two stays at the end and
// is never stored.  A block of 32 * E consecutive keys lives in registers, E per lane (element
// e = lane * E + r): steps with partner distance < E are register moves, the rest are shuffles.
__device__ __forceinline__ void ce_u64(unsigned long long &lo, unsigned long long &hi) {
    const unsigned long long a = lo, c = hi;
    const bool sw = a > c;
    lo = sw ? c : a;
    hi = sw ? a : c;
}

// Steps whose partner sits in another lane take the lane distance at run time (one compact loop body,
// friendly to the instruction cache: a fully unrolled network is instruction-fetch bound); steps inside
// a lane are unrolled with compile-time register indices.
template <int E, int K>
__device__ __forceinline__ void reg_flip_inlane(unsigned long long (&v)[E]) {  // K <= E
#pragma unroll
    for (int r = 0; r < E; ++r) {
        constexpr int m = K - 1;
        if ((r ^ m) > r) ce_u64(v[r], v[r ^ m]);
    }
}
template <int E, int J>
__device__ __forceinline__ void reg_xor_inlane(unsigned long long (&v)[E]) {  // J < E
#pragma unroll
    for (int r = 0; r < E; ++r)
        if ((r & J) == 0) ce_u64(v[r], v[r | J]);
}
template <int E, int J>
__device__ __forceinline__ void reg_xor_inlane_chain(unsigned long long (&v)[E]) {  // J, J/2, ..., 1
    if constexpr (J >= 1) {
        reg_xor_inlane<E, J>(v);
        reg_xor_inlane_chain<E, J / 2>(v);
    }
}
// stages K, 2K, ..., E entirely inside the lanes
template <int E, int K>
__device__ __forceinline__ void reg_stages_inlane(unsigned long long (&v)[E]) {
    if constexpr (K <= E) {
        reg_flip_inlane<E, K>(v);
        reg_xor_inlane_chain<E, K / 4>(v);
        reg_stages_inlane<E, 2 * K>(v);
    }
}
// mirror step of merge stage k = E * kl (kl >= 2): partner lane ^ (kl - 1), register E - 1 - r
template <int E>
__device__ __forceinline__ void reg_flip_lanes(unsigned long long (&v)[E], const int lane, const int kl) {
    const int lm = kl - 1;
    const bool lower = (lane & (kl >> 1)) == 0;  // top flipped bit clear: this element is the lower one
#pragma unroll
    for (int r = 0; r < (E + 1) / 2; ++r) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, v[E - 1 - r], lm);
        const unsigned long long o2 = __shfl_xor_sync(0xffffffffu, v[r], lm);
        v[r] = ((v[r] < o1) == lower) ? v[r] : o1;  // keys are distinct: keep the smaller one iff this is the lower slot
        if (E - 1 - r != r) v[E - 1 - r] = ((v[E - 1 - r] < o2) == lower) ? v[E - 1 - r] : o2;
    }
}
// xor step at element distance E * lj: partner lane ^ lj, same register
template <int E>
__device__ __forceinline__ void reg_xor_lanes(unsigned long long (&v)[E], const int lane, const int lj) {
    const bool lower = (lane & lj) == 0;
#pragma unroll
    for (int r = 0; r < E; ++r) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v[r], lj);
        v[r] = ((v[r] < o) == lower) ? v[r] : o;
    }
}
// the xor steps of one merge over the whole register block, from lane distance lj0 down to element distance 1
template <int E>
__device__ __forceinline__ void reg_merge_tail(unsigned long long (&v)[E], const int lane, const int lj0) {
#pragma unroll 1
    for (int lj = lj0; lj >= 1; lj >>= 1) reg_xor_lanes<E>(v, lane, lj);
    reg_xor_inlane_chain<E, E / 2>(v);
}

template <int E>
__device__ __forceinline__ void reg_load(unsigned long long (&v)[E], const unsigned long long *k, const int n, const int lane) {
#pragma unroll
    for (int r = 0; r < E; ++r) v[r] = (lane * E + r < n) ? k[lane * E + r] : ~0ull;
}
template <int E>
__device__ __forceinline__ void reg_store(const unsigned long long (&v)[E], unsigned long long *k, const int n, const int lane) {
#pragma unroll
    for (int r = 0; r < E; ++r)
        if (lane * E + r < n) k[lane * E + r] = v[r];
}

// complete sort of k[0, n), n <= 32 * E
template <int E>
__device__ __forceinline__ void warp_sort_regs(unsigned long long *k, const int n) {
    const int lane = threadIdx.x & 31;
    unsigned long long v[E];
    reg_load<E>(v, k, n, lane);
    reg_stages_inlane<E, 2>(v);
#pragma unroll 1
    for (int kl = 2; kl <= 32; kl <<= 1) {  // merge stages k = E * kl
        reg_flip_lanes<E>(v, lane, kl);
        reg_merge_tail<E>(v, lane, kl >> 2);
    }
    reg_store<E>(v, k, n, lane);
    __syncwarp();
}

constexpr int kSortE = 8;  // keys per lane of the largest register block (256 keys)

// Ascending sort of k[0, n) (shared memory) by one warp, any n.
__device__ __forceinline__ void warp_sort(unsigned long long *k, const int n) {
    const int lane = threadIdx.x & 31;
    if (n <= 1) return;
    if (n <= 32) return warp_sort_regs<1>(k, n);
    if (n <= 64) return warp_sort_regs<2>(k, n);
    if (n <= 128) return warp_sort_regs<4>(k, n);
    if (n <= 256) return warp_sort_regs<8>(k, n);
    constexpr int E = kSortE, BLK = 32 * E;
    // blocks of BLK keys sorted in registers; merge steps with partner distance >= BLK through shared memory
    int P = BLK;
    while (P < n) P <<= 1;
    for (int b0 = 0; b0 < n; b0 += BLK) warp_sort_regs<E>(k + b0, min(BLK, n - b0));
    auto ce = [&](const int i, const int q) {
        if (q < n) {
            const unsigned long long x = k[i], y = k[q];
            if (x > y) { k[i] = y; k[q] = x; }
        }
    };
    for (int kk = 2 * BLK; kk <= P; kk <<= 1) {
        const int h = kk >> 1;
        for (int t = lane; t < (P >> 1); t += 32) {
            const int base = (t / h) * kk, o = t & (h - 1);
            if (base + o >= n) break;
            ce(base + o, base + kk - 1 - o);
        }
        __syncwarp();
        for (int j = h >> 1; j >= BLK; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                if (i >= n) break;
                ce(i, i | j);
            }
            __syncwarp();
        }
        for (int b0 = 0; b0 < n; b0 += BLK) {
            unsigned long long v[E];
            reg_load<E>(v, k + b0, n - b0, lane);
            reg_merge_tail<E>(v, lane, 16);  // lane distance 16 == element distance BLK / 2
            reg_store<E>(v, k + b0, n - b0, lane);
        }
        __syncwarp();
    }
}

// Greedy NMS of one class segment [s, e) of the class-sorted keys by one warp.  The segment is in
// score order; box_r holds the image's boxes by candidate rank (the key's low bits), `off` is the
// class offset of the coordinate trick.  Kept boxes are appended to kept_box[s + k] so the test
// against earlier keeps is a broadcast read.  Stops at max_det keeps: later boxes of the class cannot
// reach the output.
__device__ __forceinline__ void warp_class_nms(const unsigned long long *keys, const float4 *box_r, float4 *kept_box,
                                               unsigned *keep_bits, const int s, const int e, const float off,
                                               const int max_det, const int flavor, const float thr_f,
                                               const double thr_d) {
    const int lane = threadIdx.x & 31;
    int K = 0;
    for (int c0 = s; c0 < e && K < max_det; c0 += 32) {
        const int i = c0 + lane;
        const bool valid = i < e;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            bx = box_r[key_slot(keys[i])];
            bx.x = bx.x + off; bx.y = bx.y + off; bx.z = bx.z + off; bx.w = bx.w + off;  // tv:ops/boxes.py:101
        }
        bool dead = !valid;
        for (int k = 0; k < K; k += 4) {
            if (__all_sync(0xffffffffu, dead)) break;  // dense clusters die against the first keeps
            // 4 independent tests per trip (entries past K are stale boxes of this segment or zeros: masked)
            const float4 k0 = kept_box[s + k], k1 = kept_box[s + k + 1], k2 = kept_box[s + k + 2], k3 = kept_box[s + k + 3];
            const bool s0 = suppresses(k0, bx, flavor, thr_f, thr_d);
            const bool s1 = k + 1 < K && suppresses(k1, bx, flavor, thr_f, thr_d);
            const bool s2 = k + 2 < K && suppresses(k2, bx, flavor, thr_f, thr_d);
            const bool s3 = k + 3 < K && suppresses(k3, bx, flavor, thr_f, thr_d);
            dead = dead || s0 || s1 || s2 || s3;
        }
        unsigned alive = __ballot_sync(0xffffffffu, !dead);
        unsigned keepm = 0u;
        int room = max_det - K;
        while (alive && room > 0) {
            const int j = __ffs(alive) - 1;  // lowest surviving lane == best remaining score: kept
            keepm |= 1u << j;
            alive &= ~(1u << j);
            --room;
            if (!alive) break;
            const float4 jb = make_float4(__shfl_sync(0xffffffffu, bx.x, j), __shfl_sync(0xffffffffu, bx.y, j),
                                          __shfl_sync(0xffffffffu, bx.z, j), __shfl_sync(0xffffffffu, bx.w, j));
            const bool sup = ((alive >> lane) & 1u) && suppresses(jb, bx, flavor, thr_f, thr_d);
            alive &= ~__ballot_sync(0xffffffffu, sup);
        }
        if ((keepm >> lane) & 1u) kept_box[s + K + __popc(keepm & ((1u << lane) - 1u))] = bx;
        if (lane == 0 && keepm) {
            const int w = c0 >> 5, sh = c0 & 31;
            atomicOr(&keep_bits[w], keepm << sh);
            if (sh && (keepm >> (32 - sh))) atomicOr(&keep_bits[w + 1], keepm >> (32 - sh));
        }
        K += __popc(keepm);
        __syncwarp();
    }
}

// writes the image's output rows (postprocess.py:43-46): score order, zero padded to max_det
template <typename SlotOf>
__device__ __forceinline__ void write_dets(const NmsParams &p, const int b, const size_t slot0, const int nkept,
                                           SlotOf slot_of) {
    if (p.variant != PLYOLO_NMS_YOLOX) {
        // sibling decoders: rows (x1,y1,x2,y2, obj, conf / best class score, class) — yolov3_decoder.py:88, yolov5_decoder.py:59
        for (int i = threadIdx.x; i < p.max_det; i += kNmsThreads) {
            float *d = p.dets + ((size_t)b * p.max_det + i) * 7;
            if (i < nkept) {
                const int slot = slot_of(i);
                const float4 bx = p.ws.box[slot0 + slot];
                const int meta = p.ws.meta[slot0 + slot];
                const float sc = p.ws.score[slot0 + slot], ax = p.ws.aux[slot0 + slot];
                d[0] = bx.x; d[1] = bx.y; d[2] = bx.z; d[3] = bx.w;
                d[4] = p.variant == PLYOLO_NMS_YOLOV3 ? ax : sc;   // objectness
                d[5] = p.variant == PLYOLO_NMS_YOLOV3 ? sc : ax;   // conf (v3) / best class score (v5)
                d[6] = (float)(meta >> 24);
                if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = meta & 0xffffff;
            } else {
                for (int q = 0; q < 7; ++q) d[q] = 0.f;
                if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
            }
        }
        if (threadIdx.x == 0) p.counts[b] = nkept;
        return;
    }
    for (int i = threadIdx.x; i < p.max_det; i += kNmsThreads) {
        if (i < nkept) {
            const int slot = slot_of(i);
            const float4 bx = p.ws.box[slot0 + slot];
            const int meta = p.ws.meta[slot0 + slot];
            store_row6(p, b, i, make_float2(bx.x, bx.y), make_float2(bx.z, bx.w), make_float2(p.ws.score[slot0 + slot], (float)(meta >> 24)));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = meta & 0xffffff;
        } else {
            store_row6(p, b, i, make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
        }
    }
    if (threadIdx.x == 0) store_count(p, b, nkept);
}

// The whole NMS of image b by the calling CTA (kNmsThreads threads, all of them must call).
// smem_raw: nms_smem_bytes(...) bytes of 16-byte aligned dynamic shared memory:
// (layout: see nms_core_bytes)
__device__ __forceinline__ void nms_image(const NmsParams &p, const int b, unsigned char *smem_raw) {
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);  // [sort_cap] / fast: [fast_cap]
    unsigned char *region = smem_raw + (size_t)p.sort_cap * 8;                    // general path buffers
    float4 *kept_fast = reinterpret_cast<float4 *>(smem_raw + (size_t)p.fast_cap * 8);   // [fast_cap]
    float4 *box_r = reinterpret_cast<float4 *>(smem_raw + (size_t)p.fast_cap * 24);      // [fast_cap]
    int *pref = reinterpret_cast<int *>(smem_raw + nms_core_bytes(p.sort_cap, p.fast_cap, p.max_det));  // [NT+1]
    unsigned *keep_bits = reinterpret_cast<unsigned *>(pref + p.NT + 1);                    // [fast_cap/32 + 1]
    __shared__ int s_wcnt[kNmsWarps];
    __shared__ float red[kNmsWarps];
    __shared__ int s_nkept, s_total, s_ncross, s_fallback, s_next;
    __shared__ unsigned s_minx, s_miny;
    __shared__ int cls_cnt[kMaxClasses], seg_begin[kMaxClasses], seg_cursor[kMaxClasses], cls_order[kMaxClasses];
    __shared__ int word_base[kFastCap / 32 + 1];
    __shared__ float4 cross_box[kMaxCross];
    __shared__ unsigned long long cross_key[kMaxCross];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = p.NT;
    const int *tcount = p.ws.tile_count + (size_t)b * NT;
    const size_t slot0 = (size_t)b * NT * kPpTile;
    NMS_PROF(0);

    // ---- exclusive prefix of the tile counts (warp 0, segmented)
    if (warp == 0) {
        const int seg = (NT + 31) / 32;
        const int lo = min(lane * seg, NT), hi = min(lo + seg, NT);
        int s = 0;
        for (int i = lo; i < hi; ++i) s += tcount[i];
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        int run = inc - s;
        for (int i = lo; i < hi; ++i) { pref[i] = run; run += tcount[i]; }
        if (lane == 31) { pref[NT] = inc; s_total = inc; }
        if (lane == 0) { s_nkept = 0; s_ncross = 0; s_fallback = 0; s_next = 0; s_minx = 0xffffffffu; s_miny = 0xffffffffu; }
    }
    if (tid < kMaxClasses) cls_cnt[tid] = 0;
    __syncthreads();
    // postprocess.py:24-25 keeps the first max_nms candidates in anchor order; the YOLOv3 / YOLOv5 decoders keep the
    // max_nms BEST ones (argsort, yolov3_decoder.py:98-100, yolov5_decoder.py:66-67): all candidates are sorted and the
    // sweep stops after max_nms of them
    const bool by_score = p.variant != PLYOLO_NMS_YOLOX;
    const int Nk = by_score ? s_total : min(s_total, p.max_nms);
    if (Nk > p.sort_cap) {  // (sibling call sites only) more candidates than one CTA can sort: reported, not guessed
        if (tid == 0) p.counts[b] = -1;
        return;
    }
    NMS_PROF(1);

    // batched_nms branch (tv:ops/boxes.py:80): per-class loop vs coordinate trick; the sibling decoders call plain nms
    const bool per_class = !by_score && !p.agnostic && 4 * (long long)Nk > ((p.flavor & PLYOLO_NMS_RULE_CPU) ? 4000 : 100000);
    const bool use_off = !p.agnostic && !per_class;
    bool fast = !by_score && !p.agnostic && Nk <= p.fast_cap;
    unsigned long long *stage = reinterpret_cast<unsigned long long *>(kept_fast);  // fast path: keys in candidate order

    // ---- one pass over the candidates (rank r = position in anchor order): key, max coordinate
    // (tv:ops/boxes.py:99 boxes.max()), class histogram and cross boxes for the fast path
    auto build_keys = [&](const bool with_class, unsigned long long *dst) -> float {
        float mx = -FLT_MAX;
        for (int r = tid; r < Nk; r += kNmsThreads) {
            int lo = 0, hi = NT;  // pref[lo] <= r < pref[hi]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pref[mid] <= r) lo = mid; else hi = mid;
            }
            const int idx = lo * kPpTile + (r - pref[lo]);
            const float sc = p.ws.score[slot0 + idx];
            const float4 bx = p.ws.box[slot0 + idx];
            // low bits: the slot (general path) or the candidate rank (fast path: indexes box_r) — same order
            unsigned long long key = ((unsigned long long)(~float_ordered(sc)) << kSlotBits) | (unsigned)(with_class ? r : idx);
            mx = fmaxf(mx, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
            if (with_class) {
                box_r[r] = bx;
                const int cl = p.ws.meta[slot0 + idx] >> 24;
                key |= (unsigned long long)cl << 57;
                atomicAdd(&cls_cnt[cl], 1);
                // A box of a HIGHER class can reach back into a lower class's offset range only if both its
                // x1 and y1 lie below -1 (+- rounding); everything else never overlaps another class.
                if (use_off && bx.x < -0.5f && bx.y < -0.5f) {
                    const int ci = atomicAdd(&s_ncross, 1);
                    if (ci < kMaxCross) { cross_box[ci] = bx; cross_key[ci] = key; }
                    atomicMin(&s_minx, float_ordered(bx.x));
                    atomicMin(&s_miny, float_ordered(bx.y));
                }
            }
            dst[r] = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) red[warp] = mx;
        __syncthreads();
        mx = red[0];
#pragma unroll
        for (int w = 1; w < kNmsWarps; ++w) mx = fmaxf(mx, red[w]);
        return mx;
    };
    const float max_coord1 = build_keys(fast, fast ? stage : keys) + 1.0f;  // max_coordinate + 1 (tv:ops/boxes.py:100)
    const float span = p.fixed_span > 0.f ? p.fixed_span : max_coord1;       // yolov5_decoder.py:70: class * 4096
    // the same-class shortcut needs offsets that dwarf their own rounding error (ulp(C*span) << 0.5) and class ranges
    // wide enough for every box (always true for max_coordinate + 1)
    const bool filter_ok = span > 0.f && span * 256.f < 4.0e6f && max_coord1 <= span;
    NMS_PROF(2);

    if (fast && (!use_off || filter_ok) && s_ncross <= kMaxCross) {
        // =========================== fast path: class-segmented NMS =================================
        // ---- class segments (exclusive prefix of the histogram) and the largest-first class order
        if (warp == 0) {
            int c4[4], s = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) { c4[u] = cls_cnt[lane * 4 + u]; s += c4[u]; }
            int inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            int run = inc - s;
#pragma unroll
            for (int u = 0; u < 4; ++u) { seg_begin[lane * 4 + u] = run; seg_cursor[lane * 4 + u] = run; run += c4[u]; }
        } else if (tid >= 32 && tid < 32 + kMaxClasses) {
            const int c = tid - 32, n = cls_cnt[c];
            int rank = 0;
            for (int o = 0; o < kMaxClasses; ++o) {
                const int m = cls_cnt[o];
                rank += (m > n || (m == n && o < c)) ? 1 : 0;
            }
            cls_order[rank] = c;
        }
        for (int i = tid; i < (p.fast_cap >> 5) + 1; i += kNmsThreads) keep_bits[i] = 0u;
        __syncthreads();
        // ---- counting scatter into the class segments + exact cross-class check
        const int ncross = s_ncross;
        const float lim_x = ncross ? ordered_float(s_minx) + span - 1.0f : 0.f;
        const float lim_y = ncross ? ordered_float(s_miny) + span - 1.0f : 0.f;
        for (int r = tid; r < Nk; r += kNmsThreads) {
            const unsigned long long ky = stage[r];
            const int cy = key_class(ky);
            keys[atomicAdd(&seg_cursor[cy], 1)] = ky;
            if (ncross) {
                // only boxes within |min x1| x |min y1| of the far corner can be reached by a cross box
                float4 y = box_r[r];
                if (!(y.z > lim_x && y.w > lim_y)) continue;
                const float offy = (float)cy * span;
                y.x = y.x + offy; y.y = y.y + offy; y.z = y.z + offy; y.w = y.w + offy;
                for (int q = 0; q < ncross; ++q) {
                    const unsigned long long kx = cross_key[q];
                    const int cx = key_class(kx);
                    if (cx == cy) continue;
                    float4 x = cross_box[q];
                    const float offx = (float)cx * span;
                    x.x = x.x + offx; x.y = x.y + offx; x.z = x.z + offx; x.w = x.w + offx;
                    if (!(x.x < y.z && x.y < y.w && y.x < x.z && y.y < x.w)) continue;  // no overlap: quotient 0
                    const bool x_first = (kx & kOrderMask) < (ky & kOrderMask);
                    if (x_first ? suppresses(x, y, p.flavor, p.thr_f, p.thr_d) : suppresses(y, x, p.flavor, p.thr_f, p.thr_d))
                        s_fallback = 1;
                }
            }
        }
        __syncthreads();  // `stage` is dead from here on: its memory becomes the kept lists
        NMS_PROF(3);
        if (!s_fallback) {
            // ---- one warp per class, largest first: sort the segment, sweep it
            for (;;) {
                int oi = 0;
                if (lane == 0) oi = atomicAdd(&s_next, 1);
                oi = __shfl_sync(0xffffffffu, oi, 0);
                if (oi >= kMaxClasses) break;
                const int c = cls_order[oi];
                const int n = cls_cnt[c];
                if (n == 0) break;
                const int s = seg_begin[c];
                const long long t0 = p.prof ? clock64() : 0;
                warp_sort(keys + s, n);
                const long long t1 = p.prof ? clock64() : 0;
                const float off = use_off ? (float)c * span : 0.f;  // tv:ops/boxes.py:100 (its own rounding)
                warp_class_nms(keys, box_r, kept_fast, keep_bits, s, s + n, off, p.max_det, p.flavor, p.thr_f, p.thr_d);
                if (p.prof && lane == 0) {
                    const long long t2 = clock64();
                    atomicMax(reinterpret_cast<unsigned long long *>(p.prof + (size_t)b * 16 + 14), ((unsigned long long)(t2 - t1) << 32) | (unsigned)n);
                    atomicMax(reinterpret_cast<unsigned long long *>(p.prof + (size_t)b * 16 + 15), ((unsigned long long)(t1 - t0) << 32) | (unsigned)n);
                }
            }
            if (p.prof && lane == 0) atomicMax(reinterpret_cast<unsigned long long *>(p.prof + (size_t)b * 16 + 13), (unsigned long long)clock64());
            __syncthreads();
            NMS_PROF(4);
            // ---- kept keys, class stripped, compacted (box_r is dead: its memory now holds keys2)
            unsigned long long *keys2 = reinterpret_cast<unsigned long long *>(box_r);
            const int nwords = (Nk + 31) >> 5;
            if (warp == 0) {
                int c4[4], s = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int w = lane * 4 + u;
                    c4[u] = w < nwords ? __popc(keep_bits[w]) : 0;
                    s += c4[u];
                }
                int inc = s;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                int run = inc - s;
#pragma unroll
                for (int u = 0; u < 4; ++u) { word_base[lane * 4 + u] = run; run += c4[u]; }
                if (lane == 31) word_base[kFastCap / 32] = inc;
            }
            __syncthreads();
            NMS_PROF(8);
            const int Kt = word_base[kFastCap / 32];
            for (int i = tid; i < Nk; i += kNmsThreads) {
                const unsigned wbits = keep_bits[i >> 5];
                if ((wbits >> (i & 31)) & 1u)
                    keys2[word_base[i >> 5] + __popc(wbits & ((1u << (i & 31)) - 1u))] = keys[i] & kOrderMask;
            }
            int n2 = 64;
            while (n2 < Kt) n2 <<= 1;
            for (int i = Kt + tid; i < n2; i += kNmsThreads) keys2[i] = ~0ull;
            __syncthreads();
            NMS_PROF(5);
            block_sort(keys2, n2);
            NMS_PROF(6);
            write_dets(p, b, slot0, min(Kt, p.max_det), [&](const int i) {
                const int r = key_slot(keys2[i]);  // candidate rank -> slot
                int lo = 0, hi = NT;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (pref[mid] <= r) lo = mid; else hi = mid;
                }
                return lo * kPpTile + (r - pref[lo]);
            });
            NMS_PROF(7);
            if (p.prof && tid == 0) { p.prof[(size_t)b * 16 + 10] = Nk; p.prof[(size_t)b * 16 + 11] = Kt; p.prof[(size_t)b * 16 + 12] = ncross; }
            return;
        }
    }
    if (fast) {
        // interleaving class ranges, too many cross boxes, or a pair of different classes suppresses:
        // redo the image with class-less keys and the exact global sweep
        __syncthreads();
        build_keys(false, keys);
    }

    // =============================== general path: global greedy sweep ==============================
    int n_pad = 64;
    while (n_pad < Nk) n_pad <<= 1;
    for (int i = Nk + tid; i < n_pad; i += kNmsThreads) keys[i] = ~0ull;
    __syncthreads();
    block_sort(keys, n_pad);
    const int Nlim = by_score ? min(Nk, p.max_nms) : Nk;  // score-sorted truncation of the sibling call sites

    float4 *kept_box = reinterpret_cast<float4 *>(region);                 // [max_det]
    float4 *cbox = kept_box + p.max_det;                                   // [kRound]
    int *kept_cls = reinterpret_cast<int *>(cbox + kRound);                // [max_det]
    int *kept_slot = kept_cls + p.max_det;                                 // [max_det]
    int *ccls = kept_slot + p.max_det;                                     // [kRound]
    int *cidx = ccls + kRound;                                             // [kRound]
    unsigned *cmask = reinterpret_cast<unsigned *>(cidx + kRound);         // [kRound * kRound / 32]

    // ---- greedy NMS in rounds of kRound candidates (score order), kSub threads per candidate:
    //   (A) test every candidate of the round against the kept list (<= max_det boxes in shared memory);
    //   (B) compact the survivors (usually a small fraction: dense clusters die against earlier keeps) and
    //       build the suppression bit-matrix among survivors only;
    //   (C) warp 0 sweeps the survivors sequentially over the remaining bits (ffs), appends the keeps,
    // and the loop exits as soon as max_det boxes are kept (output order == score order == sweep order).
    for (int base = 0; base < Nlim; base += kRound) {
        const int nkept = s_nkept;
        if (nkept >= p.max_det) break;
        const int nch = min(kRound, Nlim - base);
        const int ci = tid / kSub, sub = tid % kSub;
        // (A) every thread of a candidate fetches the same record (one broadcast request per candidate)
        bool sup = false;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        int cl = 0;
        if (ci < nch) {
            const int slot = key_slot(keys[base + ci]);
            bx = p.ws.box[slot0 + slot];
            cl = p.ws.meta[slot0 + slot] >> 24;
            if (use_off) {
                if (!filter_ok || (bx.x < -0.5f && bx.y < -0.5f)) cl |= kCrossBit;
                const float off = (float)(cl & 0xff) * span;  // tv:ops/boxes.py:100-101 (separate roundings)
                bx.x = bx.x + off; bx.y = bx.y + off; bx.z = bx.z + off; bx.w = bx.w + off;
            } else if (!per_class) {
                cl |= kCrossBit;  // class-agnostic: every pair is tested
            }
            for (int k = sub; k < nkept; k += kSub) {
                const int kc = kept_cls[k];
                if (kc != cl && !((kc | cl) & kCrossBit)) continue;  // different class, neither can cross
                sup |= suppresses(kept_box[k], bx, p.flavor, p.thr_f, p.thr_d);
            }
        }
#pragma unroll
        for (int o = 1; o < kSub; o <<= 1) sup |= __shfl_xor_sync(0xffffffffu, sup ? 1 : 0, o) != 0;
        const bool alive = ci < nch && !sup;
        // survivors of each warp (lanes with sub == 0 speak for their candidate)
        const unsigned am = __ballot_sync(0xffffffffu, alive && sub == 0);
        if (lane == 0) s_wcnt[warp] = __popc(am);
        __syncthreads();
        int abase = 0, n_al = 0;
#pragma unroll
        for (int w = 0; w < kNmsWarps; ++w) {
            if (w < warp) abase += s_wcnt[w];
            n_al += s_wcnt[w];
        }
        if (alive && sub == 0) {
            const int r = abase + __popc(am & ((1u << lane) - 1u));
            cbox[r] = bx;
            ccls[r] = cl;
            cidx[r] = ci;
        }
        __syncthreads();
        // (B) row r of the survivor matrix: bit c set <=> survivor r (if kept) suppresses the later survivor c
        {
            const int r = tid / kSub;
            float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);
            int rc = 0;
            if (r < n_al) { rb = cbox[r]; rc = ccls[r]; }
#pragma unroll
            for (int w = 0; w < kRound / 32; ++w) {
                unsigned m = 0u;
                if (r < n_al && (w << 5) + 31 > r) {
                    const int c_hi = min(n_al, (w + 1) << 5);
                    for (int c = (w << 5) + sub; c < c_hi; c += kSub) {
                        const int cc = ccls[c];
                        if (c > r && (cc == rc || ((cc | rc) & kCrossBit)) &&
                            suppresses(rb, cbox[c], p.flavor, p.thr_f, p.thr_d))
                            m |= 1u << (c & 31);
                    }
                }
#pragma unroll
                for (int o = 1; o < kSub; o <<= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
                if (r < n_al && sub == 0) cmask[r * (kRound / 32) + w] = m;
            }
        }
        __syncthreads();
        if (warp == 0) {
            // (C) lane l < kRound/32 owns word l of the removed / kept bit-vectors over survivor indices
            unsigned removed_w = 0, keep_w = 0;
            if (lane < kRound / 32) {
                const int lo = lane << 5;
                removed_w = n_al >= lo + 32 ? 0u : (n_al <= lo ? 0xffffffffu : (0xffffffffu << (n_al - lo)));
            }
            int nk = nkept;
            for (int wd = 0; wd < kRound / 32 && nk < p.max_det; ++wd) {
                while (nk < p.max_det) {
                    const unsigned avail = ~__shfl_sync(0xffffffffu, removed_w, wd);
                    if (!avail) break;
                    const int bit = __ffs(avail) - 1;
                    const int r = (wd << 5) + bit;
                    if (lane == wd) { keep_w |= 1u << bit; removed_w |= 1u << bit; }
                    if (lane < kRound / 32) removed_w |= cmask[r * (kRound / 32) + lane];
                    ++nk;
                }
            }
            // append the keeps in order
            int pos = nkept;
            for (int wd = 0; wd < kRound / 32; ++wd) {
                const unsigned kw = __shfl_sync(0xffffffffu, keep_w, wd);
                if ((kw >> lane) & 1u) {
                    const int dst = pos + __popc(kw & ((1u << lane) - 1u));
                    const int r = (wd << 5) + lane;
                    kept_box[dst] = cbox[r];
                    kept_cls[dst] = ccls[r];
                    kept_slot[dst] = key_slot(keys[base + cidx[r]]);
                }
                pos += __popc(kw);
            }
            if (lane == 0) s_nkept = nk;
        }
        __syncthreads();
    }
    write_dets(p, b, slot0, s_nkept, [&](const int i) { return kept_slot[i]; });
}

// ---- team sort: 8 warps (256 threads, named barrier `bar_id`) sort one large class segment -----------
__device__ __forceinline__ void team_barrier(const int bar_id) {
    asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
}

// Ascending sort of k[0, n) in shared memory by a team of 8 warps: blocks of 32 * E keys are sorted in
// registers (one block per warp at a time), merge steps with partner distance >= a block go through
// shared memory, the rest of each merge again in registers.  Same ascending-only network as warp_sort.
template <int E>
__device__ __forceinline__ void team_sort_e(unsigned long long *k, const int n, const int tw, const int bar_id) {
    constexpr int BLK = 32 * E;
    const int lane = threadIdx.x & 31, tt = tw * 32 + lane;  // thread index inside the team
    int P = BLK;
    while (P < n) P <<= 1;
    for (int b0 = tw * BLK; b0 < n; b0 += 8 * BLK) warp_sort_regs<E>(k + b0, min(BLK, n - b0));
    team_barrier(bar_id);
    auto ce = [&](const int i, const int q) {
        if (q < n) {
            const unsigned long long x = k[i], y = k[q];
            if (x > y) { k[i] = y; k[q] = x; }
        }
    };
    for (int kk = 2 * BLK; kk <= P; kk <<= 1) {
        const int h = kk >> 1;
        for (int t = tt; t < (P >> 1); t += 256) {
            const int base = (t / h) * kk, o = t & (h - 1);
            if (base + o >= n) break;
            ce(base + o, base + kk - 1 - o);
        }
        team_barrier(bar_id);
        for (int j = h >> 1; j >= BLK; j >>= 1) {
            for (int t = tt; t < (P >> 1); t += 256) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                if (i >= n) break;
                ce(i, i | j);
            }
            team_barrier(bar_id);
        }
        for (int b0 = tw * BLK; b0 < n; b0 += 8 * BLK) {
            unsigned long long v[E];
            reg_load<E>(v, k + b0, n - b0, lane);
            reg_merge_tail<E>(v, lane, 16);  // lane distance 16 == element distance BLK / 2
            reg_store<E>(v, k + b0, n - b0, lane);
        }
        team_barrier(bar_id);
    }
}

__device__ __forceinline__ void team_sort(unsigned long long *k, const int n, const int tw, const int bar_id) {
    if (n <= 512) team_sort_e<2>(k, n, tw, bar_id);
    else if (n <= 1024) team_sort_e<4>(k, n, tw, bar_id);
    else team_sort_e<8>(k, n, tw, bar_id);
}

constexpr int kTeamMin = 192;  // classes with more candidates are sorted by a team of 8 warps

// ---- general path: one CTA per image that needs it ------------------------------------------------------
// An image is redone here, exactly, when the class-split kernel (nms_fast.cuh) finds that it cannot handle it
// (max_nms truncation, a class group above its capacity, a cross-class pair that suppresses, list overflows).
//  * Plain stream launches: that kernel launches this one FROM THE DEVICE for the one image (CUDA dynamic parallelism,
//    fire-and-forget: the images that need it run side by side, and the stream does not move on before they have
//    finished) — nothing at all is launched for the normal step.
//  * Under stream capture (CUDA graphs) the host launches it behind the class-split kernel, one CTA per image, each
//    leaving at once unless its image was flagged: fire-and-forget children of a graph's kernel node are NOT ordered
//    before the rest of the graph / the next replay (measured: a replayed step "finished" in 60 us and its rows arrived
//    several replays later), so device-side launches are not used there.
//  * The host also launches it when the call as a whole cannot use the class split (`all_general`: class-agnostic NMS,
//    the YOLOv3 / YOLOv5 call sites, very large max_det, anchors beyond the fast key layout).
__global__ void __launch_bounds__(kNmsThreads, 1) nms_general_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char nms_smem[];
    // programmatic dependent launch: the CTAs may be scheduled while the previous kernel drains; everything below
    // needs its results
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int b = p.only_image >= 0 ? p.only_image : (int)blockIdx.x;
    if (p.only_image < 0 && !p.all_general && __ldcg(&p.ws.ctr[b * kImgCtr + kCtrGeneral]) == 0) return;
    nms_image(p, b, nms_smem);
}

}  // namespace plyolo
