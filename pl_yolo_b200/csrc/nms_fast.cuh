// nms_fast.cuh — class-aware NMS of one image by a cluster of kGroups small CTAs, built to run UNDER the score
// kernel: launched with programmatic dependent launch, 512 threads / ~51 KB of shared memory / <= 64 registers, so
// that one CTA fits on every SM next to the persistent score CTA; every cluster waits for its image's "all tiles
// scored" counter (release / acquire at GPU scope) and then works out of L2 while the score kernel keeps
// streaming the later images from HBM.  Semantics: torchvision.ops.batched_nms (tv:ops/boxes.py:51-120,
// torchvision::nms) + the tail of postprocess (models/evaluators/postprocess.py:43-46).
//
// Per (image, class group) CTA (class & 3 == group; the score kernel bucketed key + box per group):
//   1. load the bucket (keys + boxes, coalesced) into registers / shared memory, class offset applied
//      (tv:ops/boxes.py:100-101, separate roundings); the best key of every class by atomicMin;
//   2. pre-kill: the best box of a class is always kept by the greedy sweep, so every candidate it suppresses is
//      dead for certain and never enters the sort (dense clusters lose most of their members here); exact
//      cross-class check of the boxes near the far corner against the cross list (a hit -> general path);
//   3. counting scatter of the survivors into class segments, per-class bitonic sorts (registers; teams of 8
//      warps for large classes) of keys that carry the box index, so no gather follows the sort;
//   4. greedy sweep, one work item per chunk of 32 boxes in (class, chunk) order: a chunk tests its boxes against
//      the keeps of the earlier chunks of its class as they become final (release / acquire flags in shared
//      memory), settles its own boxes, appends its keeps to the group's kept-key list;
//   5. cluster merge through distributed shared memory: every CTA ranks its kept keys among the kept keys of all
//      groups by counting (keys are distinct: rank == output row) and writes its rows.
// Everything the class split cannot do exactly (class-agnostic, max_nms truncation, a group above kFastCapG,
// too many cross boxes, a cross-class pair that suppresses, more than max_det keeps in one group) raises the image's
// general-path flag; nms_general_kernel (nms.cuh: nms_image) then redoes that image exactly.
#pragma once

#include "nms.cuh"

namespace plyolo {

constexpr int kFastThreads = 512;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kFastPer = 3;                           // candidates per thread
constexpr int kFastCapG = kFastThreads * kFastPer;    // 1536 candidates per (image, class group)
constexpr int kGC = kMaxClasses / kGroups;            // class slots per group: class c -> slot c / kGroups
constexpr int kFastCross = 256;                       // cross boxes that can reach another class's range (after the filter)
constexpr int kAllKeys = 1024;                        // kept keys of all groups in the merge
constexpr int kKeys2Cap = 896;                        // kept keys of one group (more -> general path)
constexpr int kIdxBits = 11;                          // box index inside the group (kFastCapG <= 2048)
constexpr int kFastSlotBits = 21;                     // candidate slot (tile * 128 + position) must fit
constexpr unsigned long long kIdxMask = (1ull << kIdxBits) - 1ull;
constexpr size_t kFastUnion = (size_t)(kAllKeys + 64) * 8;  // cross list | kept index lists | merged key list
static_assert(kFastUnion >= (size_t)kFastCapG * 2 && kFastUnion >= (size_t)kFastCross * 24, "union region too small");
static_assert(kFastCapG <= (1 << kIdxBits) && kGroups == 4 && kGC == 32, "layout assumptions");

// dynamic shared memory: skey[cap] u64 | sbox[cap] float4 | union | keys2[kKeys2Cap] u64  (52 KB; + 1.5 KB static:
// fits on an SM beside the score CTA's 171.5 KB)
constexpr size_t kFastSmemBytes = (size_t)kFastCapG * 24 + kFastUnion + (size_t)kKeys2Cap * 8;
constexpr int kFastMaxDet = kKeys2Cap - 32;  // max_det the fast kernel supports

__device__ __forceinline__ int ld_acquire_gpu(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int lds_acquire_cta(const int *p) {
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_release_cta(int *p, const int v) {
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// sort key of the fast kernel: [63:32] ~ordered(score) | [31:11] slot | [10:0] box index in the group.
// Ascending = score descending, ties -> lower slot (anchor order); the global order key is key >> kIdxBits.
__device__ __forceinline__ int fkey_idx(const unsigned long long k) { return (int)(k & kIdxMask); }

// Register budget of the co-residency (per SM sub-partition: 16384 registers): the score CTA puts 5 of its 18 warps on
// one sub-partition, this CTA 4 of its 16: 5 * 32 * kScoreRegs + 4 * 32 * kFastRegs <= 16384.
#ifndef PLYOLO_NMS_REGS
#define PLYOLO_NMS_REGS 48
#endif
constexpr int kFastRegs = PLYOLO_NMS_REGS;

__global__ void __cluster_dims__(kGroups, 1, 1) __maxnreg__(kFastRegs) nms_fast_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char fsm[];
    unsigned long long *skey = reinterpret_cast<unsigned long long *>(fsm);                       // [cap] class segments
    float4 *sbox = reinterpret_cast<float4 *>(fsm + (size_t)kFastCapG * 8);                        // [cap] by box index
    unsigned char *uni = fsm + (size_t)kFastCapG * 24;
    float4 *x_box = reinterpret_cast<float4 *>(uni);                                               // [kFastCross] (steps 1-2)
    unsigned long long *x_key = reinterpret_cast<unsigned long long *>(uni + (size_t)kFastCross * 16);
    unsigned short *klist = reinterpret_cast<unsigned short *>(uni);                               // [cap] (step 4)
    unsigned long long *allk = reinterpret_cast<unsigned long long *>(uni);                        // [kAllKeys + 64] (step 5)
    unsigned long long *keys2 = reinterpret_cast<unsigned long long *>(uni + kFastUnion);          // [kKeys2Cap]
    __shared__ unsigned c_best_hi[kGC], c_best_lo[kGC];  // best (~score) of the class, then (slot | box index) among the ties
    __shared__ int g_cnt[kGC], g_begin[kGC], g_cursor[kGC], g_order[kGC], g_cum[kGC + 1], g_fin[kGC];
    __shared__ int g_kcum[kFastCapG / 32 + kGC];
    __shared__ int g_next, g_next2, g_fallback, g_nbig, g_nk2, g_nx;
    __shared__ unsigned x_minx[kMaxClasses], x_miny[kMaxClasses];  // per class: min x1 / y1 of its cross boxes (ordered uint)
    __shared__ float c_limx[kGC], c_limy[kGC];                     // per class of the group: far-corner limits
    __shared__ int c_kpub, c_fallback;  // read by the other CTAs of the cluster

    const int g = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = p.NT;
    const size_t slot0 = (size_t)b * NT * kPpTile;
    int *ctr = p.ws.ctr + b * kImgCtr;
    long long *prof = p.prof ? p.prof + ((size_t)b * kGroups + g) * 16 : nullptr;
#define FPROF(slot) do { if (prof && tid == 0) prof[slot] = clock64(); } while (0)
    FPROF(0);

    // ---- wait until every tile of the image has been scored (the score kernel may still be running)
    if (tid < kGC) { c_best_hi[tid] = 0xffffffffu; c_best_lo[tid] = 0xffffffffu; g_cnt[tid] = 0; g_fin[tid] = 0; }
    if (tid < kMaxClasses) { x_minx[tid] = 0xffffffffu; x_miny[tid] = 0xffffffffu; }
    if (tid == 0) {
        g_next = 0; g_next2 = 0; g_fallback = 0; g_nbig = 0; g_nk2 = 0; g_nx = 0;
        if (p.wait_tiles)
            while (ld_acquire_gpu(&ctr[kCtrDone]) < NT) __nanosleep(64);
    }
    __syncthreads();
    FPROF(1);
    // the last image's cluster keeps the grid alive until the score grid has completed and flushed: whatever follows
    // this kernel in the stream is then ordered after both kernels
    const bool last_image = b == (int)gridDim.y - 1;
#define FAST_EXIT() do { if (last_image) asm volatile("griddepcontrol.wait;" ::: "memory"); return; } while (0)

    // the bucket is read before its fill count is known (entries past the count are stale bytes of the caller-owned
    // workspace, never used): one L2 round trip for counters, keys and boxes instead of two
    const unsigned long long *bucket = p.ws.gkey + ((size_t)b * kGroups + g) * kFastCapG;
    const float4 *bbox = p.ws.gbox + ((size_t)b * kGroups + g) * kFastCapG;
    unsigned long long key[kFastPer];
    float4 bx[kFastPer];
#pragma unroll
    for (int k = 0; k < kFastPer; ++k) {
        key[k] = __ldcg(bucket + tid + k * kFastThreads);
        bx[k] = __ldcg(bbox + tid + k * kFastThreads);
    }
    int total = 0, gmax = 0, n = 0;
#pragma unroll
    for (int q = 0; q < kGroups; ++q) {
        const int c = __ldcg(&ctr[q]);
        total += c;
        gmax = max(gmax, c);
        if (q == g) n = c;
    }
    const int xc = __ldcg(&ctr[kGroups + 1]);
    const float span = ordered_float((unsigned)__ldcg(&ctr[kGroups])) + 1.0f;  // max_coordinate + 1 (tv:ops/boxes.py:100)
    const bool per_class = 4 * (long long)total > ((p.flavor & PLYOLO_NMS_RULE_CPU) ? 4000 : 100000);  // tv:ops/boxes.py:80
    const bool use_off = !per_class;
    const bool filter_ok = span > 0.f && span * 256.f < 4.0e6f;
    if (total == 0) {  // no candidate: postprocess.py:26-27 leaves None (count 0, zero rows)
        if (g == 0) {
            for (int i = tid; i < p.max_det; i += kFastThreads) {
                float2 *d = reinterpret_cast<float2 *>(p.dets + ((size_t)b * p.max_det + i) * 6);
                d[0] = make_float2(0.f, 0.f); d[1] = make_float2(0.f, 0.f); d[2] = make_float2(0.f, 0.f);
                if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
            }
            if (tid == 0) p.counts[b] = 0;
        }
        FAST_EXIT();
    }
    const bool fast = total <= p.max_nms && gmax <= kFastCapG && (!use_off || (filter_ok && xc <= kMaxCross));
    if (!fast) {
        if (g == 0 && tid == 0) ctr[kCtrGeneral] = 1;
        FAST_EXIT();
    }

    // ---- 1. load: keys + boxes of the bucket (thread t owns candidates t, t + 512, t + 1024), cross list
    const bool xcheck = use_off && xc > 0;
    unsigned long long nk[kFastPer];
    float bz[kFastPer], bw[kFastPer];
    int cl[kFastPer];
    unsigned alive = 0u;
    {
        if (xcheck && tid < xc) {
            // A cross box x (class k) meets a box y of a class c < k only if y.x2 > x.x1 + (k - c) * span - 1 (rounding
            // < 1) and the same in y: with k - c >= 1 and y.x2 <= the image's largest x2, only cross boxes with
            // x1 < max x2 - span + 1 and y1 < max y2 - span + 1 can matter at all — usually none of them.
            float4 x = __ldcg(p.ws.xbox + (size_t)b * kMaxCross + tid);
            const float reach_x = ordered_float((unsigned)__ldcg(&ctr[kCtrMaxX2])) - span + 1.0f;
            const float reach_y = ordered_float((unsigned)__ldcg(&ctr[kCtrMaxY2])) - span + 1.0f;
            if (x.x < reach_x && x.y < reach_y) {
                const unsigned long long kx = __ldcg(p.ws.xkey + (size_t)b * kMaxCross + tid);
                const int xi = atomicAdd(&g_nx, 1);
                if (xi < kFastCross) {
                    atomicMin(&x_minx[key_class(kx)], float_ordered(x.x));
                    atomicMin(&x_miny[key_class(kx)], float_ordered(x.y));
                    const float offx = (float)key_class(kx) * span;
                    x.x = x.x + offx; x.y = x.y + offx; x.z = x.z + offx; x.w = x.w + offx;
                    x_box[xi] = x;
                    x_key[xi] = kx;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kFastPer; ++k) {
            const int i = tid + k * kFastThreads;
            nk[k] = ~0ull; cl[k] = 0; bz[k] = 0.f; bw[k] = 0.f;
            if (i < n) {
                const int c = key_class(key[k]);
                cl[k] = c;
                nk[k] = ((key[k] >> kSlotBits) << 32) | ((key[k] & kSlotMask) << kIdxBits) | (unsigned)i;
                bz[k] = bx[k].z; bw[k] = bx[k].w;
                const float off = use_off ? (float)c * span : 0.f;  // tv:ops/boxes.py:100-101 (separate roundings)
                float4 y = bx[k];
                y.x = y.x + off; y.y = y.y + off; y.z = y.z + off; y.w = y.w + off;
                sbox[i] = y;
                atomicMin(&c_best_hi[c / kGroups], (unsigned)(nk[k] >> 32));
            }
        }
    }
    __syncthreads();
    // the best key of every class (64-bit shared-memory atomics are CAS loops: two native 32-bit minima instead)
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
        if (tid + k * kFastThreads < n && (unsigned)(nk[k] >> 32) == c_best_hi[cl[k] / kGroups])
            atomicMin(&c_best_lo[cl[k] / kGroups], (unsigned)nk[k]);
    const int nx = min(g_nx, kFastCross);
    if (tid == 0 && g_nx > kFastCross) g_fallback = 1;  // (never seen) more reaching cross boxes than the list holds
    if (nx > 0) {
        // A box of class c meets a cross box of a class k > c only if its x2 / y2 exceed (min x1 / y1 of that class's
        // cross boxes) + (k - c) * span (rounding: < 1): per-class limits, 16 threads per class of the group
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
    __syncthreads();
    FPROF(2);

    // ---- 2. pre-kill against the class's best box + exact cross-class check + histogram of the survivors
    {
#pragma unroll
        for (int k = 0; k < kFastPer; ++k) {
            const int i = tid + k * kFastThreads;
            if (i < n) {
                const int c = cl[k];
                const float4 y = sbox[i];
                const int ti = (int)(c_best_lo[c / kGroups] & (unsigned)kIdxMask);
                bool dead = false;
                if (ti != i) dead = suppresses(sbox[ti], y, p.flavor, p.thr_f, p.thr_d);
                if (nx > 0 && bz[k] > c_limx[c / kGroups] && bw[k] > c_limy[c / kGroups]) {
                    const unsigned long long oy = nk[k] >> kIdxBits;  // (~score, slot) with the slot in 21 bits
                    for (int q = 0; q < nx; ++q) {
                        const unsigned long long kx = x_key[q];
                        if (key_class(kx) <= c) continue;  // the pair is found from the lower class's side
                        const float4 x = x_box[q];
                        if (!(x.x < y.z && x.y < y.w && y.x < x.z && y.y < x.w)) continue;  // no overlap: quotient 0
                        const unsigned long long ox = (((kx & kOrderMask) >> kSlotBits) << kFastSlotBits) | (kx & kSlotMask);
                        const bool x_first = ox < oy;
                        if (x_first ? suppresses(x, y, p.flavor, p.thr_f, p.thr_d) : suppresses(y, x, p.flavor, p.thr_f, p.thr_d))
                            g_fallback = 1;
                    }
                }
                if (!dead) {
                    alive |= 1u << k;
                    atomicAdd(&g_cnt[c / kGroups], 1);
                }
            }
        }
    }
    __syncthreads();
    FPROF(3);

    // ---- class segments of the survivors (exclusive prefix), largest-first class order, sweep items
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
        const unsigned big = __ballot_sync(0xffffffffu, c > kTeamMin);
        if (lane == 0) g_nbig = __popc(big);
        __syncwarp();
        const int chunks = (g_cnt[g_order[lane]] + 31) >> 5;  // sweep items of the lane-th largest class
        int inc = chunks;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        g_cum[lane] = inc - chunks;
        if (lane == 31) g_cum[kGC] = inc;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kFastPer; ++k)
        if ((alive >> k) & 1u) skey[atomicAdd(&g_cursor[cl[k] / kGroups], 1)] = nk[k];
    __syncthreads();
    FPROF(4);

    // ---- 3. sort every class segment (the keys carry the box index: nothing to gather afterwards)
    {
        const int nbig = g_nbig;
        const int team = warp >> 3, tw = warp & 7;
        for (int oi = team; oi < nbig; oi += kFastWarps / 8) {  // large classes: one team of 8 warps each
            const int cs = g_order[oi];
            team_sort(skey + g_begin[cs], g_cnt[cs], tw, 1 + team);
        }
        for (;;) {  // the rest: one warp each, largest first
            int oi = 0;
            if (lane == 0) oi = nbig + atomicAdd(&g_next, 1);
            oi = __shfl_sync(0xffffffffu, oi, 0);
            if (oi >= kGC) break;
            const int cs = g_order[oi], nc = g_cnt[cs];
            if (nc == 0) break;
            warp_sort(skey + g_begin[cs], nc);
        }
    }
    __syncthreads();
    FPROF(5);

    // ---- 4. sweep: one item per chunk of 32 boxes, handed out in (class, chunk) order
    const int n_items = g_cum[kGC];
    for (;;) {
        int item = 0;
        if (lane == 0) item = atomicAdd(&g_next2, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= n_items) break;
        int lo = 0, hi = kGC;  // g_cum[lo] <= item < g_cum[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (g_cum[mid] <= item) lo = mid; else hi = mid;
        }
        const int cs = g_order[lo], s = g_begin[cs], nc = g_cnt[cs];
        const int j = item - g_cum[lo], item0 = g_cum[lo];
        const bool valid = 32 * j + lane < nc;
        const unsigned long long ky = valid ? skey[s + 32 * j + lane] : ~0ull;
        const int bi = fkey_idx(ky);
        const float4 bx = valid ? sbox[bi] : make_float4(0.f, 0.f, 0.f, 0.f);
        bool dead = !valid;
        int K = 0;
        for (int jj = 0; jj < j; ++jj) {
            while (lds_acquire_cta(&g_fin[cs]) <= jj) __nanosleep(20);  // chunk jj's keeps were written before its flag
            const int Knew = g_kcum[item0 + jj];
            for (int k = K; k < Knew; k += 4) {
                if (__all_sync(0xffffffffu, dead)) break;  // dense clusters die against the first keeps
                // 4 independent tests per trip (entries past Knew: masked)
                const float4 k0 = sbox[klist[s + k]];
                const float4 k1 = sbox[klist[s + min(k + 1, Knew - 1)]];
                const float4 k2 = sbox[klist[s + min(k + 2, Knew - 1)]];
                const float4 k3 = sbox[klist[s + min(k + 3, Knew - 1)]];
                const bool s0 = suppresses(k0, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s1 = k + 1 < Knew && suppresses(k1, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s2 = k + 2 < Knew && suppresses(k2, bx, p.flavor, p.thr_f, p.thr_d);
                const bool s3 = k + 3 < Knew && suppresses(k3, bx, p.flavor, p.thr_f, p.thr_d);
                dead = dead || s0 || s1 || s2 || s3;
            }
            K = Knew;
        }
        // settle the chunk: lowest surviving lane == best remaining score: kept (stops at max_det keeps: later
        // boxes of the class cannot reach the output)
        unsigned live = __ballot_sync(0xffffffffu, !dead);
        unsigned keepm = 0u;
        int room = p.max_det - K;
        while (live && room > 0) {
            const int jl = __ffs(live) - 1;
            keepm |= 1u << jl;
            live &= ~(1u << jl);
            --room;
            if (!live) break;
            const float4 jb = make_float4(__shfl_sync(0xffffffffu, bx.x, jl), __shfl_sync(0xffffffffu, bx.y, jl),
                                          __shfl_sync(0xffffffffu, bx.z, jl), __shfl_sync(0xffffffffu, bx.w, jl));
            const bool sup = ((live >> lane) & 1u) && suppresses(jb, bx, p.flavor, p.thr_f, p.thr_d);
            live &= ~__ballot_sync(0xffffffffu, sup);
        }
        const int nkeep = __popc(keepm), r = __popc(keepm & ((1u << lane) - 1u));
        int base2 = 0;
        if (lane == 0 && nkeep) base2 = atomicAdd(&g_nk2, nkeep);
        base2 = __shfl_sync(0xffffffffu, base2, 0);
        if ((keepm >> lane) & 1u) {
            klist[s + K + r] = (unsigned short)bi;
            if (base2 + r < kKeys2Cap) keys2[base2 + r] = ky >> kIdxBits;  // the group's kept keys (any order)
        }
        if (lane == 0) g_kcum[item] = K + nkeep;
        __syncwarp();
        if (lane == 0) sts_release_cta(&g_fin[cs], j + 1);
    }
    __syncthreads();
    FPROF(6);

    // ---- 5. merge across the image's kGroups CTAs through distributed shared memory
    const int Kg = g_nk2;
    const int Kpub = min(Kg, p.max_det);
    if (Kg > p.max_det && Kg <= kKeys2Cap) {
        // rare: only the group's first max_det keeps in global order can reach the output — rank by counting
        unsigned long long *tmp = allk;  // the kept index lists are dead
        for (int i = tid; i < Kg; i += kFastThreads) {
            const unsigned long long key = keys2[i];
            int rank = 0;
            for (int j = 0; j < Kg; ++j) rank += keys2[j] < key ? 1 : 0;
            if (rank < p.max_det) tmp[rank] = key;
        }
        __syncthreads();
        for (int i = tid; i < p.max_det; i += kFastThreads) keys2[i] = tmp[i];
        __syncthreads();
    }
    if (tid == 0) { c_kpub = Kpub; c_fallback = (g_fallback || Kg > kKeys2Cap) ? 1 : 0; }
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    FPROF(7);
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
    FPROF(8);
    if (prof && tid == 0) { prof[10] = n; prof[11] = Kg; prof[12] = xc; prof[13] = g_cum[kGC]; prof[14] = g_begin[kGC - 1] + g_cnt[kGC - 1]; }
    if (fb) {  // a pair of different classes suppresses (or a list overflowed): the exact global sweep redoes the image
        if (g == 0 && tid == 0) ctr[kCtrGeneral] = 1;
        FAST_EXIT();
    }
    const int nkept = min(total_k, p.max_det);
    // rank of every own key = number of smaller keys in all lists (padded with +inf to a multiple of 64):
    // 4 threads per key (128 keys per pass: normally one pass), 8 independent compares per trip; the key's record
    // is fetched while the count runs
    constexpr int kSubT = 4;
    const int total_pad = (total_k + 63) & ~63;
    for (int i0 = 0; i0 < Kpub; i0 += kFastThreads / kSubT) {
        const int i = i0 + tid / kSubT, sub = tid % kSubT;
        const bool owner = i < Kpub && sub == 0;
        const unsigned long long key = i < Kpub ? keys2[i] : 0ull;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        int meta = 0;
        if (owner) {
            const int slot = (int)(key & ((1ull << kFastSlotBits) - 1ull));
            bx = __ldcg(p.ws.box + slot0 + slot);
            meta = __ldcg(p.ws.meta + slot0 + slot);
        }
        int rank = 0;
        if (i < Kpub) {
            for (int j = sub; j < total_pad; j += 16 * kSubT) {
#pragma unroll
                for (int u = 0; u < 16; ++u) rank += allk[j + u * kSubT] < key ? 1 : 0;
            }
        }
#pragma unroll
        for (int o = 1; o < kSubT; o <<= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
        if (owner && rank < nkept) {
            const float sc = ordered_float(~(unsigned)(key >> kFastSlotBits));  // the key holds the score bits
            float2 *d = reinterpret_cast<float2 *>(p.dets + ((size_t)b * p.max_det + rank) * 6);
            d[0] = make_float2(bx.x, bx.y);
            d[1] = make_float2(bx.z, bx.w);
            d[2] = make_float2(sc, (float)(meta >> 24));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + rank] = meta & 0xffffff;
        }
    }
    if (g == 0) {
        for (int i = nkept + tid; i < p.max_det; i += kFastThreads) {
            float2 *d = reinterpret_cast<float2 *>(p.dets + ((size_t)b * p.max_det + i) * 6);
            d[0] = make_float2(0.f, 0.f); d[1] = make_float2(0.f, 0.f); d[2] = make_float2(0.f, 0.f);
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
        }
        if (tid == 0) p.counts[b] = nkept;
    }
    FPROF(9);
    FAST_EXIT();
#undef FPROF
#undef FAST_EXIT
}

}  // namespace plyolo
