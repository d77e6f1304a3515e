// nms_fast.cuh — class-aware NMS of one image by a cluster of kGroups small CTAs, built to run UNDER the score
// kernel: launched with programmatic dependent launch, 512 threads / ~51 KB of shared memory / 48 registers, so
// that one CTA fits on every SM next to the persistent score CTA; every cluster waits for its image's "all tiles
// scored" counter (release / acquire at GPU scope) and then works out of L2 while the score kernel keeps
// streaming the later images from HBM.  Semantics: torchvision.ops.batched_nms (tv:ops/boxes.py:51-120,
// torchvision::nms) + the tail of postprocess (models/evaluators/postprocess.py:43-46).
//
// Per (image, class group) CTA (class & 3 == group; the score kernel bucketed key + box per group):
//   1. load the bucket (keys + boxes, coalesced; issued together with the counters) into registers / shared memory;
//   2. kFastRounds rounds of "max first": the best surviving box of a class that is not a round winner yet is
//      certainly kept by the greedy sweep (everything above it is a winner that did not suppress it, or dead), so
//      whatever it suppresses is dead for certain — the CTA finds the per-class best with two native 32-bit
//      shared-memory minima and every thread tests its candidates against it.  Dense clusters (the bulk of the
//      candidates) disappear here, fully in parallel; cross boxes that could suppress across classes are checked
//      exactly on the way (a hit -> general path);
//   3. counting scatter of the survivors into class segments; one warp per class (largest first): bitonic sort in
//      registers of keys that carry the box index, then the greedy sweep of the class in chunks of 32, keeps
//      appended to the group's kept list;
//   4. cluster merge through distributed shared memory: every CTA ranks its kept keys among the kept keys of all
//      groups by counting (keys are distinct: rank == output row) and writes its rows straight from shared memory.
// Everything the class split cannot do exactly (class-agnostic, max_nms truncation, a group above kFastCapG,
// a cross-class pair that suppresses, list overflows) raises the image's general-path flag; nms_general_kernel
// (nms.cuh: nms_image) then redoes that image exactly.
#pragma once

#include "nms.cuh"

namespace plyolo {

constexpr int kGC = 32;                               // class slots per group: class c -> slot c / kGroups
#ifndef PLYOLO_NMS_ROUNDS
#define PLYOLO_NMS_ROUNDS 4
#endif
constexpr int kFastRounds = PLYOLO_NMS_ROUNDS;        // max-first rounds before the sort
constexpr int kFastCross = 128;                       // cross boxes that can suppress across classes (after the filters)
constexpr int kAllKeys = 960;                         // kept keys of all groups in the merge
constexpr int kKeys2Cap = 640;                        // kept keys of one group (more -> general path)
constexpr int kIdxBits = 11;                          // box index inside the group (cap <= 2048)
constexpr int kFastAnchorBits = 21;                   // anchor index must fit (A <= 2 097 152)
constexpr unsigned long long kIdxMask = (1ull << kIdxBits) - 1ull;
constexpr size_t kFastUnion = (size_t)(kAllKeys + 64) * 8;  // cross list | kept index lists | merged key list
static_assert(kFastUnion >= (size_t)1536 * 2 && kFastUnion >= (size_t)kFastCross * 24, "union region too small");
constexpr int kFastMaxDet = kKeys2Cap - 32;  // max_det the fast kernel supports

// One image = one cluster of kGroups class-group CTAs: 512 threads x 3 candidates (1536 per group), 48 registers.
// (An 8-group layout — 256 threads x 4, clusters of 8 — was measured in round 2: bit-identical, but the phases are
// latency- rather than issue-bound, so an image's NMS was not shorter (17.8 vs 15.6 us after its last tile), and a
// cluster of 8 needs 8 free SMs of one GPC beside the score CTAs: cfg2 69.7 vs 45.0 us.  Removed.)
constexpr int kFastThreads = 512;
constexpr int kFastPer = 3;                           // candidates per thread
constexpr int kFastCapG = kFastThreads * kFastPer;    // candidates per (image, class group) the kernel stages
constexpr int kFastRegs = 48;
constexpr int kXPer = kMaxCross / kFastThreads;       // cross-list entries per thread
static_assert(kFastCapG <= kBucketCap && kFastCapG <= (1 << kIdxBits), "layout assumptions");
// dynamic shared memory: skey[cap] u64 | sbox[cap] float4 | scls[cap] u8 | union | keys2[kKeys2Cap] u64 | kidx2 u16
constexpr size_t kFastSmemBytes = (size_t)kFastCapG * 25 + kFastUnion + (size_t)kKeys2Cap * 10;

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// sort key of the fast kernel: [63:32] ~ordered(score) | [31:11] anchor | [10:0] box index in the group.
// Ascending = score descending, ties -> lower anchor; the global order key is key >> kIdxBits.
__device__ __forceinline__ int fkey_idx(const unsigned long long k) { return (int)(k & kIdxMask); }

__device__ __forceinline__ float4 shift_box(const float4 b, const float off) {  // tv:ops/boxes.py:101 (its own rounding)
    return make_float4(b.x + off, b.y + off, b.z + off, b.w + off);
}

// hands image b to the general path: device-side fire-and-forget launch of one CTA of nms_general_kernel (see nms.cuh)
__device__ __forceinline__ void launch_general_for(const NmsParams &p, const int b) {
    NmsParams q = p;
    q.only_image = b;
    q.all_general = 1;
    q.prof = nullptr;
    q.wait_tiles = 0;
    nms_general_kernel<<<1, kNmsThreads, p.general_smem, cudaStreamFireAndForget>>>(q);
}

// Register budget of the co-residency (per SM sub-partition: 16384 registers): the score CTA puts at most 5 of its 19 warps on
// one sub-partition, this CTA 4 of its 16: 5 * 32 * kScoreRegs + 4 * 32 * kFastRegs <= 16384.
__global__ void __cluster_dims__(kGroups, 1, 1) __maxnreg__(kFastRegs) nms_fast_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char fsm[];
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(fsm);                       // [cap] class segments
    float4 *sbox = reinterpret_cast<float4 *>(fsm + (size_t)kFastCapG * 8);                        // [cap] by box index, un-offset
    unsigned char *scls = fsm + (size_t)kFastCapG * 24;                                            // [cap] class by box index
    unsigned char *uni = fsm + (size_t)kFastCapG * 25;
    float4 *x_box = reinterpret_cast<float4 *>(uni);                                               // [kFastCross] (steps 1-2)
    unsigned long long *x_key = reinterpret_cast<unsigned long long *>(uni + (size_t)kFastCross * 16);
    unsigned short *klist = reinterpret_cast<unsigned short *>(uni);                               // [cap] (step 3)
    unsigned long long *allk = reinterpret_cast<unsigned long long *>(uni);                        // [kAllKeys + 64] (step 4)
    unsigned long long *keys2 = reinterpret_cast<unsigned long long *>(uni + kFastUnion);          // [kKeys2Cap]
    unsigned short *kidx2 = reinterpret_cast<unsigned short *>(uni + kFastUnion + (size_t)kKeys2Cap * 8);
    // per round: best (~score) of the class among the live non-winners, then (anchor | box index) among the ties
    __shared__ unsigned c_best_hi[kFastRounds][kGC], c_best_lo[kFastRounds][kGC];
    __shared__ int g_cnt[kGC], g_begin[kGC], g_cursor[kGC], g_order[kGC];
    __shared__ int g_next, g_fallback, g_nk2, g_nx;
    // cross-box scratch of the first round, in the (still unused) sort-key region: per class the min x1 / y1 of its
    // cross boxes (ordered uint), per class of the group the far-corner limits
    unsigned *x_minx = reinterpret_cast<unsigned *>(skey), *x_miny = x_minx + kMaxClasses;
    float *c_limx = reinterpret_cast<float *>(x_miny + kMaxClasses), *c_limy = c_limx + kGC;
    __shared__ int c_kpub, c_fallback;  // read by the other CTAs of the cluster

    const int g = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = p.NT;
    int *ctr = p.ws.ctr + b * kImgCtr;
    long long *prof = p.prof ? p.prof + ((size_t)b * kGroups + g) * 16 : nullptr;
#define FPROF(slot) do { if (prof && tid == 0) prof[slot] = clock64(); } while (0)
    FPROF(0);

    // ---- wait until every tile of the image has been scored (the score kernel may still be running)
    if (tid < kGC) {
#pragma unroll
        for (int r = 0; r < kFastRounds; ++r) { c_best_hi[r][tid] = 0xffffffffu; c_best_lo[r][tid] = 0xffffffffu; }
        g_cnt[tid] = 0;
    }
    for (int i = tid; i < kMaxClasses; i += kFastThreads) { x_minx[i] = 0xffffffffu; x_miny[i] = 0xffffffffu; }
    if (tid == 0) {
        g_next = 0; g_fallback = 0; g_nk2 = 0; g_nx = 0;
        if (p.wait_tiles)
            while (ld_acquire_gpu(&ctr[kCtrDone]) < NT) __nanosleep(64);
    }
    __syncthreads();
    FPROF(1);
    if (prof && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); prof[13] = (long long)gt; }
    // the last image's cluster keeps the grid alive until the score grid has completed and flushed: whatever follows
    // this kernel in the stream is then ordered after both kernels
    const bool last_image = b == (int)gridDim.y - 1;
#define FAST_EXIT() do { if (last_image) asm volatile("griddepcontrol.wait;" ::: "memory"); return; } while (0)

    // the bucket is read before its fill count is known (entries past the count are stale bytes of the caller-owned
    // workspace, never used): one L2 round trip for counters, keys and boxes instead of two
    const unsigned long long *bucket = p.ws.gkey + (size_t)b * kBucketImg + g * kBucketCap;
    const float4 *bbox = p.ws.gbox + (size_t)b * kBucketImg + g * kBucketCap;
    unsigned long long nk[kFastPer];
    int cl[kFastPer];
    float4 xb_spec[kXPer];
    unsigned long long xk_spec[kXPer];
    static_assert(kXPer * kFastThreads == kMaxCross, "the cross list is read kXPer entries per thread");
    {
        unsigned long long key[kFastPer];
        float4 bx[kFastPer];
#pragma unroll
        for (int k = 0; k < kFastPer; ++k) {
            key[k] = __ldcg(bucket + tid + k * kFastThreads);
            bx[k] = __ldcg(bbox + tid + k * kFastThreads);
        }
#pragma unroll
        for (int e = 0; e < kXPer; ++e) {
            xb_spec[e] = __ldcg(p.ws.xbox + (size_t)b * kMaxCross + tid + e * kFastThreads);
            xk_spec[e] = __ldcg(p.ws.xkey + (size_t)b * kMaxCross + tid + e * kFastThreads);
        }
#pragma unroll
        for (int k = 0; k < kFastPer; ++k) {  // meaningless past the fill count: never used there
            const int i = tid + k * kFastThreads;
            cl[k] = key_class(key[k]);
            nk[k] = ((key[k] >> kSlotBits) << 32) | ((key[k] & kSlotMask) << kIdxBits) | (unsigned)i;
            sbox[i] = bx[k];
            scls[i] = (unsigned char)cl[k];
        }
    }
    int total = 0, gmax = 0, n = 0;
#pragma unroll
    for (int q = 0; q < kGroups; ++q) {
        const int c = __ldcg(&ctr[q]);
        total += c;
        gmax = max(gmax, c);
        if (q == g) n = c;
    }
    const int xc = __ldcg(&ctr[kCtrCross]);
    const float max_x2 = ordered_float((unsigned)__ldcg(&ctr[kCtrMaxX2])), max_y2 = ordered_float((unsigned)__ldcg(&ctr[kCtrMaxY2]));
    const float span = ordered_float((unsigned)__ldcg(&ctr[kCtrMaxCoord])) + 1.0f;  // max_coordinate + 1 (tv:ops/boxes.py:100)
    const bool per_class = 4 * (long long)total > ((p.flavor & PLYOLO_NMS_RULE_CPU) ? 4000 : 100000);  // tv:ops/boxes.py:80
    const bool use_off = !per_class;
    const bool filter_ok = span > 0.f && span * 256.f < 4.0e6f;
    if (total == 0) {  // no candidate: postprocess.py:26-27 leaves None (count 0, zero rows)
        if (g == 0) {
            for (int i = tid; i < p.max_det; i += kFastThreads) {
                store_row6(p, b, i, make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f));
                if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
            }
            if (tid == 0) store_count(p, b, 0);
        }
        FAST_EXIT();
    }
    const bool fast = total <= p.max_nms && gmax <= kFastCapG && (!use_off || (filter_ok && xc <= kMaxCross));
    if (!fast) {
        if (g == 0 && tid == 0) { ctr[kCtrGeneral] = 1; if (p.device_launch) launch_general_for(p, b); }
        FAST_EXIT();
    }

    // ---- 1b. cross boxes (x1, y1 < -0.5: the only boxes that can reach into a LOWER class's offset range).
    // A cross box x (class k) and a box y of a class c < k intersect, after the offsets, in at most
    //   wmax = (max x2 - span + 1) - x.x1  by  hmax = (max y2 - span + 1) - x.y1      (k - c >= 1; rounding < 1),
    // and IoU > thr needs inter > thr * union >= thr * area(x): a cross box with wmax * hmax < thr * area(x) can
    // neither suppress nor be suppressed across classes.  Real boxes never pass (most of the box would have to lie
    // beyond the far corner of every other box), so the list is normally empty and the check below costs nothing.
    const bool xcheck = use_off && xc > 0;
    if (xcheck) {
#pragma unroll
        for (int e = 0; e < kXPer; ++e) {
            if (tid + e * kFastThreads >= xc) continue;
            const float4 x = xb_spec[e];
            const float wmax = (max_x2 - span + 1.0f) - x.x;
            const float hmax = (max_y2 - span + 1.0f) - x.y;
            const float a_lo = fmaxf((x.z - x.x) - 1.0f, 0.f) * fmaxf((x.w - x.y) - 1.0f, 0.f) * 0.99999f;
            const bool drop = !(wmax > 0.f && hmax > 0.f) || (p.thr_f > 0.f && (wmax * hmax) * 1.00001f < p.thr_f * a_lo);
            if (!drop) {
                const unsigned long long kx = xk_spec[e];
                const int xi = atomicAdd(&g_nx, 1);
                if (xi < kFastCross) {
                    atomicMin(&x_minx[key_class(kx)], float_ordered(x.x));
                    atomicMin(&x_miny[key_class(kx)], float_ordered(x.y));
                    x_box[xi] = shift_box(x, (float)key_class(kx) * span);
                    x_key[xi] = kx;
                }
            }
        }
    }

    // ---- 2. max-first rounds.  live bit k: candidate k of this thread may still be kept; win bit k: it won a round
    unsigned live = 0u, win = 0u;
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
        if (tid + k * kFastThreads < n) live |= 1u << k;
    // (Measured on B200: pre-reducing the per-class minima inside the warp — match.any + redux, or segmented shuffle
    // minima over runs of equal classes — costs more than the same-address shared-memory atomics it saves.)
    int nx = 0;
#pragma unroll 1
    for (int r = 0; r < kFastRounds; ++r) {
#pragma unroll
        for (int k = 0; k < kFastPer; ++k)
            if (((live & ~win) >> k) & 1u) atomicMin(&c_best_hi[r][cl[k] / kGroups], (unsigned)(nk[k] >> 32));
        __syncthreads();
        // (64-bit shared-memory atomics are CAS loops: the ties of the best score settle a second 32-bit minimum)
#pragma unroll
        for (int k = 0; k < kFastPer; ++k)
            if ((((live & ~win) >> k) & 1u) && (unsigned)(nk[k] >> 32) == c_best_hi[r][cl[k] / kGroups])
                atomicMin(&c_best_lo[r][cl[k] / kGroups], (unsigned)nk[k]);
        if (r == 0) {
            nx = min(g_nx, kFastCross);
            if (tid == 0 && g_nx > kFastCross) g_fallback = 1;  // (never seen) more such cross boxes than the list holds
            if (nx > 0) {
                // A box of class c meets a cross box of a class k > c only if its x2 / y2 exceed (min x1 / y1 of that
                // class's cross boxes) + (k - c) * span (rounding: < 1): per-class limits, 16 threads per class
                const int cs = tid >> 4, part = tid & 15, c = cs * kGroups + g;
                float lx = 3.0e38f, ly = 3.0e38f;
                for (int k = c + 1 + part; k < kMaxClasses; k += 16) {
                    const unsigned ox = x_minx[k];
                    if (ox != 0xffffffffu) {
                        const float d = (float)(k - c) * span - 1.0f;
                        lx = fminf(lx, ordered_float(ox) + d);
                        ly = fminf(ly, ordered_float(x_miny[k]) + d);
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
                    lx = fminf(lx, __shfl_xor_sync(0xffffffffu, lx, o));
                    ly = fminf(ly, __shfl_xor_sync(0xffffffffu, ly, o));
                }
                if (part == 0) { c_limx[cs] = lx; c_limy[cs] = ly; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kFastPer; ++k) {
            if (!((live >> k) & 1u)) continue;
            const int i = tid + k * kFastThreads, c = cl[k];
            const float off = use_off ? (float)c * span : 0.f;  // tv:ops/boxes.py:100
            const float4 y0 = sbox[i];
            const float4 y = shift_box(y0, off);
            if (r == 0 && nx > 0 && y0.z > c_limx[c / kGroups] && y0.w > c_limy[c / kGroups]) {
                // exact cross-class check (rare): any suppressing pair sends the image to the general path
                const unsigned long long oy = nk[k] >> kIdxBits;  // (~score, anchor)
                for (int q = 0; q < nx; ++q) {
                    const unsigned long long kx = x_key[q];
                    if (key_class(kx) <= c) continue;  // the pair is found from the lower class's side
                    const float4 x = x_box[q];
                    if (!(x.x < y.z && x.y < y.w && y.x < x.z && y.y < x.w)) continue;  // no overlap: quotient 0
                    const unsigned long long ox = (((kx & kOrderMask) >> kSlotBits) << kFastAnchorBits) | (kx & kSlotMask);
                    if (ox < oy ? suppresses(x, y, p.flavor, p.thr_f, p.thr_d) : suppresses(y, x, p.flavor, p.thr_f, p.thr_d))
                        g_fallback = 1;
                }
            }
            if ((win >> k) & 1u) continue;
            const unsigned lo = c_best_lo[r][c / kGroups];
            if (lo == 0xffffffffu) continue;  // the class has no live non-winner left
            const int ti = (int)(lo & (unsigned)kIdxMask);
            if (ti == i) { win |= 1u << k; continue; }
            if (suppresses(shift_box(sbox[ti], off), y, p.flavor, p.thr_f, p.thr_d)) live &= ~(1u << k);
        }
    }
    FPROF(2);
    // histogram of the survivors
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
        if ((live >> k) & 1u) atomicAdd(&g_cnt[cl[k] / kGroups], 1);
    __syncthreads();
    FPROF(3);

    // ---- 3. class segments of the survivors (exclusive prefix), largest-first class order
    if (warp == 0) {
        const int c = g_cnt[lane];
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        g_begin[lane] = inc - c;
        g_cursor[lane] = inc - c;
    } else if (warp == 1) {
        const int c = g_cnt[lane];
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            const int m = __shfl_sync(0xffffffffu, c, o);
            rank += (m > c || (m == c && o < lane)) ? 1 : 0;
        }
        g_order[rank] = lane;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
        if ((live >> k) & 1u) skey[atomicAdd(&g_cursor[cl[k] / kGroups], 1)] = nk[k];
    __syncthreads();
    FPROF(4);

    // one warp per class, largest first: sort the segment (the keys carry the box index: nothing to gather), then the
    // greedy sweep in chunks of 32 against the class's kept list (box indices, broadcast reads)
    for (;;) {
        int oi = 0;
        if (lane == 0) oi = atomicAdd(&g_next, 1);
        oi = __shfl_sync(0xffffffffu, oi, 0);
        if (oi >= kGC) break;
        const int cs = g_order[oi], nc = g_cnt[cs];
        if (nc == 0) break;
        const int s = g_begin[cs];
        const long long ts0 = prof ? clock64() : 0;
        warp_sort(skey + s, nc);
        if (prof && tid == 0) prof[15] += clock64() - ts0;
        const float off = use_off ? (float)(cs * kGroups + g) * span : 0.f;
        int K = 0;
        for (int c0 = 0; c0 < nc && K < p.max_det; c0 += 32) {
            const bool valid = c0 + lane < nc;
            const unsigned long long ky = valid ? skey[s + c0 + lane] : ~0ull;
            const int bi = fkey_idx(ky);
            const float4 bx = valid ? shift_box(sbox[bi], off) : make_float4(0.f, 0.f, 0.f, 0.f);
            bool dead = !valid;
            for (int k = 0; k < K; k += 4) {
                if (__all_sync(0xffffffffu, dead)) break;  // dense clusters die against the first keeps
                // 4 independent tests per trip (entries past K: clamped and masked)
                const float4 k0 = shift_box(sbox[klist[s + k]], off);
                const float4 k1 = shift_box(sbox[klist[s + min(k + 1, K - 1)]], off);
                const float4 k2 = shift_box(sbox[klist[s + min(k + 2, K - 1)]], off);
                const float4 k3 = shift_box(sbox[klist[s + min(k + 3, K - 1)]], off);
                const bool s0 = suppresses(k0, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s1 = k + 1 < K && suppresses(k1, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s2 = k + 2 < K && suppresses(k2, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s3 = k + 3 < K && suppresses(k3, bx, p.flavor, p.thr_f, p.thr_d);
                dead = dead || s0 || s1 || s2 || s3;
            }
            // settle the chunk (lanes ascending == score descending; stops at max_det keeps: later boxes of the class
            // cannot reach the output)
            unsigned alive = __ballot_sync(0xffffffffu, !dead);
            const int room = p.max_det - K;
            if (p.thr_f >= 0.f) {
                // Only a box that intersects a later live box can suppress anything inside the chunk.  A geometric
                // pre-pass (independent shuffles, no division) finds, per lane, the lower live lanes it intersects;
                // the serial part then only visits the lanes somebody intersects — after the max-first rounds the
                // survivors rarely touch each other, so most chunks settle without a single iteration.
                const int cnt = min(32, nc - c0);
                unsigned ovl = 0u;
                for (int j = 0; j < cnt - 1; ++j) {
                    const float jx = __shfl_sync(0xffffffffu, bx.x, j), jy = __shfl_sync(0xffffffffu, bx.y, j);
                    const float jz = __shfl_sync(0xffffffffu, bx.z, j), jw = __shfl_sync(0xffffffffu, bx.w, j);
                    const bool hit = lane > j && fmaxf(jx, bx.x) < fminf(jz, bx.z) && fmaxf(jy, bx.y) < fminf(jw, bx.w);
                    ovl |= hit ? (1u << j) : 0u;
                }
                ovl = dead ? 0u : (ovl & alive);
                unsigned rem = __reduce_or_sync(0xffffffffu, ovl);  // live lanes that a later live lane intersects
                while (rem) {
                    const int jl = __ffs(rem) - 1;
                    rem &= rem - 1u;
                    if (!((alive >> jl) & 1u)) continue;  // suppressed meanwhile: suppresses nobody
                    const float4 jb = make_float4(__shfl_sync(0xffffffffu, bx.x, jl), __shfl_sync(0xffffffffu, bx.y, jl),
                                                  __shfl_sync(0xffffffffu, bx.z, jl), __shfl_sync(0xffffffffu, bx.w, jl));
                    const bool sup = ((alive >> lane) & 1u) && ((ovl >> jl) & 1u) && suppresses(jb, bx, p.flavor, p.thr_f, p.thr_d);
                    alive &= ~__ballot_sync(0xffffffffu, sup);
                }
            } else {
                // negative threshold: boxes that do not even touch suppress each other — the plain serial sweep
                unsigned left = alive, keep = 0u;
                while (left) {
                    const int jl = __ffs(left) - 1;
                    keep |= 1u << jl;
                    left &= ~(1u << jl);
                    const float4 jb = make_float4(__shfl_sync(0xffffffffu, bx.x, jl), __shfl_sync(0xffffffffu, bx.y, jl),
                                                  __shfl_sync(0xffffffffu, bx.z, jl), __shfl_sync(0xffffffffu, bx.w, jl));
                    const bool sup = ((left >> lane) & 1u) && suppresses(jb, bx, p.flavor, p.thr_f, p.thr_d);
                    left &= ~__ballot_sync(0xffffffffu, sup);
                }
                alive = keep;
            }
            unsigned keepm = alive;  // whoever is still alive is kept, best first
            if (__popc(keepm) > room) keepm &= (1u << __fns(keepm, 0, room + 1)) - 1u;
            const int nkeep = __popc(keepm), rk = __popc(keepm & ((1u << lane) - 1u));
            int base2 = 0;
            if (lane == 0 && nkeep) base2 = atomicAdd(&g_nk2, nkeep);
            base2 = __shfl_sync(0xffffffffu, base2, 0);
            if ((keepm >> lane) & 1u) {
                klist[s + K + rk] = (unsigned short)bi;
                if (base2 + rk < kKeys2Cap) {  // the group's kept keys (any order) and their boxes
                    keys2[base2 + rk] = ky >> kIdxBits;
                    kidx2[base2 + rk] = (unsigned short)bi;
                }
            }
            K += nkeep;
            __syncwarp();
        }
    }
    __syncthreads();
    FPROF(5);

    // ---- 4. merge across the image's kGroups CTAs through distributed shared memory
    const int Kg = g_nk2;
    const int Kpub = min(Kg, p.max_det);
    if (Kg > p.max_det && Kg <= kKeys2Cap) {
        // rare: only the group's first max_det keeps in global order can reach the output — rank by counting
        unsigned long long *tmp = allk;  // the kept index lists are dead
        unsigned short *tmpi = reinterpret_cast<unsigned short *>(skey);  // so are the sort keys
        for (int i = tid; i < Kg; i += kFastThreads) {
            const unsigned long long key = keys2[i];
            int rank = 0;
            for (int j = 0; j < Kg; ++j) rank += keys2[j] < key ? 1 : 0;
            if (rank < p.max_det) { tmp[rank] = key; tmpi[rank] = kidx2[i]; }
        }
        __syncthreads();
        for (int i = tid; i < p.max_det; i += kFastThreads) { keys2[i] = tmp[i]; kidx2[i] = tmpi[i]; }
        __syncthreads();
    }
    if (tid == 0) { c_kpub = Kpub; c_fallback = (g_fallback || Kg > kKeys2Cap) ? 1 : 0; }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    FPROF(6);
    int kq[kGroups], total_k = 0, fb = 0;
#pragma unroll
    for (int q = 0; q < kGroups; ++q) {
        kq[q] = *cluster.map_shared_rank(&c_kpub, q);
        fb |= *cluster.map_shared_rank(&c_fallback, q);
        total_k += kq[q];
    }
    if (total_k > kAllKeys) fb = 1;
    if (!fb) {
        int base = 0;
#pragma unroll
        for (int q = 0; q < kGroups; ++q) {
            const unsigned long long *src = cluster.map_shared_rank(keys2, q);
            for (int i = tid; i < kq[q]; i += kFastThreads) allk[base + i] = src[i];
            base += kq[q];
        }
        for (int i = total_k + tid; i < ((total_k + 63) & ~63); i += kFastThreads) allk[i] = ~0ull;
    }
    cluster.sync();  // nobody reads another CTA's shared memory past this point (also a CTA barrier)
    FPROF(7);
    if (prof && tid == 0) { prof[10] = n; prof[11] = Kg; prof[14] = g_begin[kGC - 1] + g_cnt[kGC - 1]; }
    if (fb) {  // a pair of different classes suppresses (or a list overflowed): the exact global sweep redoes the image
        if (g == 0 && tid == 0) { ctr[kCtrGeneral] = 1; if (p.device_launch) launch_general_for(p, b); }
        FAST_EXIT();
    }
    const int nkept = min(total_k, p.max_det);
    // rank of every own key = number of smaller keys in all lists (padded with +inf to a multiple of 64):
    // 4 threads per key (128 keys per pass: normally one pass), 16 independent compares per trip; the row is written
    // straight from shared memory (box by index, score and anchor from the key, class by index)
    constexpr int kSubT = 4;
    const int total_pad = (total_k + 63) & ~63;
    for (int i0 = 0; i0 < Kpub; i0 += kFastThreads / kSubT) {
        const int i = i0 + tid / kSubT, sub = tid % kSubT;
        const unsigned long long key = i < Kpub ? keys2[i] : 0ull;
        int rank = 0;
        if (i < Kpub) {
            for (int j = sub; j < total_pad; j += 16 * kSubT) {
#pragma unroll
                for (int u = 0; u < 16; ++u) rank += allk[j + u * kSubT] < key ? 1 : 0;
            }
        }
#pragma unroll
        for (int o = 1; o < kSubT; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (i < Kpub && sub == 0 && rank < nkept) {
            const int bi = kidx2[i];
            const float4 bx = sbox[bi];
            const float sc = ordered_float(~(unsigned)(key >> kFastAnchorBits));  // the key holds the score bits
            store_row6_local(p, b, rank, make_float2(bx.x, bx.y), make_float2(bx.z, bx.w), make_float2(sc, (float)scls[bi]));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + rank] = (int)(key & ((1ull << kFastAnchorBits) - 1ull));
        }
    }
    if (g == 0) {
        for (int i = nkept + tid; i < p.max_det; i += kFastThreads) {
            store_row6_local(p, b, i, make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
        }
        if (tid == 0) store_count(p, b, nkept);
    }
    if (p.n_peers > 0) {
        // Multi-GPU exchange: the image's finished block [max_det, 6] goes to every peer's gathered buffer over NVLink.
        // Row-by-row stores from the ranking threads (24 B at scattered rows, three partial sectors each) cost 1.3 us per
        // peer and image — with 7 peers more than the NMS itself; instead the cluster waits for its rows to be complete
        // (release / acquire at cluster scope), and every CTA copies a quarter of the block to every peer with fully
        // coalesced 8-byte stores (256 B per warp instruction) read back from L2.
        cluster.sync();
        const int chunks = p.max_det * 3;  // float2 chunks of the block
        const int per = (chunks + kGroups - 1) / kGroups, c0 = g * per, c1 = min(chunks, c0 + per);
        const float2 *src = reinterpret_cast<const float2 *>(p.dets + (size_t)b * p.max_det * 6);
        for (int w = tid; w < (c1 - c0) * p.n_peers; w += kFastThreads) {
            const int r = w / (c1 - c0), c = c0 + (w - r * (c1 - c0));
            reinterpret_cast<float2 *>(p.peer_dets[r] + (size_t)b * p.max_det * 6)[c] = __ldcg(src + c);
        }
    }
    FPROF(8);
    if (prof && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt)); prof[15] += 0; prof[12] = (long long)gt; }
    FAST_EXIT();
#undef FPROF
#undef FAST_EXIT
}

}  // namespace plyolo
