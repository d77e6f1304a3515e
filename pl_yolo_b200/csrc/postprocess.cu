// postprocess.cu — confidence filter + class-aware NMS (models/evaluators/postprocess.py:7-48,
// tv:ops/boxes.py:51-120, torchvision::nms), from materialised predictions or fused with the decode.
//
// Stage 1  score_kernel<FUSED>   one CTA per tile of 128 anchors, HBM-bound.
//   The tile lands in shared memory through the TMA engine's 1-D bulk copies
//   (cp.async.bulk + mbarrier, no register staging):
//     FUSED : 5+C row copies of 512 B from the channel-planar head maps  -> tile[channel][anchor]
//     preds : one contiguous copy of 128*(5+C) floats                     -> tile[anchor][channel]
//   Pass A (thread per anchor): sigmoid(obj), max raw class logit, a conservative pre-filter that can
//   only over-admit.  Pass B (4 lanes per admitted anchor): exact max / first argmax over the
//   sigmoid VALUES (T1), conf = sigmoid(obj) * class_conf, conf >= thr in fp32.  The survivors are
//   compacted IN ANCHOR ORDER (warp ballot + prefix) into the tile's slot range of the candidate arrays.
// Stage 2  nms_kernel            one CTA per image, latency-bound.
//   tile counts -> prefix -> the first max_nms candidates in anchor order (T2) -> 64-bit keys
//   (~ordered(score) << 32 | slot) -> bitonic sort (register/shuffle steps inside warps, shared
//   memory only for distances >= 64) == stable descending sort -> greedy NMS in rounds of 128
//   candidates against the kept list (<= max_det, early exit: output order == score order == sweep
//   order) with torchvision's arithmetic (coordinate-trick offsets, asymmetric FMA, IEEE division;
//   see `suppresses`).
#include <cfloat>

#include "common.cuh"

namespace plyolo {

constexpr int kPpTile = 128;
constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
constexpr int kMaxSortCap = 16384;
constexpr int kRound = 256;  // candidates per NMS round
constexpr int kSub = kNmsThreads / kRound;  // threads per candidate
constexpr int kCrossBit = 0x100;  // class word flag: this box must be tested against every class

struct CandWs {
    int *tile_count;    // [B, NT]
    float4 *box;        // [B, NT*128]  original (un-offset) corners
    float *score;       // [B, NT*128]
    int *meta;          // [B, NT*128]  anchor | class << 24
};

struct ScoreParams {
    Levels lv;           // FUSED only
    const float *preds;  // !FUSED only
    int B, A, C, ch, NT;
    float conf_thr;
    int bulk_ok;
    CandWs ws;
};

template <bool FUSED>
__global__ void __launch_bounds__(kPpTile) score_kernel(const ScoreParams p) {
    extern __shared__ __align__(128) float tile[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int warp_cnt[kPpTile / 32];

    const int b = blockIdx.y;
    const int tile_id = blockIdx.x;
    const int tid = threadIdx.x;
    const int ch = p.ch;
    int l = 0, a0, cnt, anchor_base;
    const float *src;
    if (FUSED) {
#pragma unroll
        for (int i = 1; i < PLYOLO_MAX_LEVELS; ++i)
            if (i < p.lv.n && tile_id >= p.lv.tile0[i]) l = i;
        a0 = (tile_id - p.lv.tile0[l]) * kPpTile;
        cnt = min(kPpTile, p.lv.hw[l] - a0);
        anchor_base = p.lv.off[l] + a0;
        src = p.lv.ptr[l] + (size_t)b * ch * p.lv.hw[l] + a0;
    } else {
        a0 = tile_id * kPpTile;
        cnt = min(kPpTile, p.A - a0);
        anchor_base = a0;
        src = p.preds + ((size_t)b * p.A + a0) * ch;
    }

    // ---- stage the tile in shared memory
    if (p.bulk_ok) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (FUSED) {
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx(&bar, (uint32_t)(ch * cnt * 4));
                __syncwarp();
                for (int c = tid; c < ch; c += 32)
                    bulk_g2s(tile + c * kPpTile, src + (size_t)c * p.lv.hw[l], (uint32_t)(cnt * 4), &bar);
            }
        } else if (tid == 0) {
            mbar_expect_tx(&bar, (uint32_t)(cnt * ch * 4));
            bulk_g2s(tile, src, (uint32_t)(cnt * ch * 4), &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        if (FUSED) {
            for (int c = 0; c < ch; ++c)
                if (tid < cnt) tile[c * kPpTile + tid] = __ldg(src + (size_t)c * p.lv.hw[l] + tid);
        } else {
            for (int i = tid; i < cnt * ch; i += kPpTile) tile[i] = __ldg(src + i);
        }
        __syncthreads();
    }

    const int lane = tid & 31, warp = tid >> 5;
    bool pass = false;
    float conf = 0.f;
    int cls = 0;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    int src_t = tid;  // tile-local anchor this thread finally reports

    if (FUSED) {
        // ---- pass A (one thread per anchor): sigmoid(obj), max raw class logit, conservative pre-filter.
        // conf = fl(so * max_c sigmoid(x_c)).  sigmoid is monotone up to a few ulp, so with m = max_c x_c:
        // conf <= so * sigmoid(m) * (1 + 1e-6); anchors with fl(fl(so * sigmoid(m)) * 1.00001f) < thr cannot pass.
        // Survivors (exactness restored in pass B) are compacted in anchor order.
        __shared__ unsigned char surv[kPpTile];
        __shared__ float s_so[kPpTile];
        __shared__ float s_best[kPpTile];
        __shared__ int s_cls[kPpTile];
        bool pre = false;
        float so = 0.f;
        if (tid < cnt) {
            so = sigmoid_ref(tile[4 * kPpTile + tid]);  // yolox_loss.py:26
            if (so >= p.conf_thr) {                     // class_conf <= 1 and fp32 multiply is monotone
                float m = tile[5 * kPpTile + tid];
                for (int c = 1; c < p.C; ++c) m = fmaxf(m, tile[(5 + c) * kPpTile + tid]);
                pre = (so * sigmoid_ref(m)) * 1.00001f >= p.conf_thr;
            }
        }
        const unsigned pm = __ballot_sync(0xffffffffu, pre);
        if (lane == 0) warp_cnt[warp] = __popc(pm);
        __syncthreads();
        int sbase = 0, nsurv = 0;
#pragma unroll
        for (int w = 0; w < kPpTile / 32; ++w) {
            if (w < warp) sbase += warp_cnt[w];
            nsurv += warp_cnt[w];
        }
        if (pre) {
            const int si = sbase + __popc(pm & ((1u << lane) - 1u));
            surv[si] = (unsigned char)tid;
            s_so[si] = so;
        }
        __syncthreads();
        // ---- pass B: 4 lanes per survivor, each takes a quarter of the classes: exact max / first argmax
        // over the sigmoid VALUES (postprocess.py:18 works on the already-squashed tensor; T1)
        const int cq = (p.C + 3) >> 2;
        for (int it = 0; it * (kPpTile >> 2) < nsurv; ++it) {
            const int si = it * (kPpTile >> 2) + (tid >> 2), q = tid & 3;
            float best = -1.f;
            int bi = 0x7fffffff;
            if (si < nsurv) {
                const int t = surv[si];
                const int c0 = q * cq, c1 = min(p.C, c0 + cq);
                for (int c = c0; c < c1; ++c) {
                    const float v = sigmoid_ref(tile[(5 + c) * kPpTile + t]);  // yolox_loss.py:27
                    if (v > best) { best = v; bi = c; }
                }
            }
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (si < nsurv && q == 0) { s_best[si] = best; s_cls[si] = bi; }
        }
        __syncthreads();
        // ---- final filter + box decode, one thread per survivor (still anchor order)
        if (tid < nsurv) {
            src_t = surv[tid];
            cls = s_cls[tid];
            conf = s_so[tid] * s_best[tid];    // postprocess.py:19
            pass = conf >= p.conf_thr;         // :20 (fp32 compare)
            if (pass) {
                const int a = a0 + src_t;
                const int W = p.lv.w[l];
                const float s = p.lv.stride[l];
                const float cx = (tile[0 * kPpTile + src_t] + (float)(a % W)) * s;  // yolox_loss.py:217
                const float cy = (tile[1 * kPpTile + src_t] + (float)(a / W)) * s;
                const float w = expf(tile[2 * kPpTile + src_t]) * s;                // :219
                const float h = expf(tile[3 * kPpTile + src_t]) * s;
                box = make_float4(cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2);  // :31-34
            }
        }
        __syncthreads();  // warp_cnt is reused below
    } else if (tid < cnt) {
        const float *r = tile + tid * ch;
        float best = r[5];
        for (int c = 1; c < p.C; ++c) {
            const float s = r[5 + c];
            if (s > best) { best = s; cls = c; }  // postprocess.py:18, first max index
        }
        conf = r[4] * best;
        pass = conf >= p.conf_thr;
        box = make_float4(r[0], r[1], r[2], r[3]);
    }

    // ---- order-preserving compaction into the tile's slots (postprocess.py:23 keeps anchor order)
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kPpTile / 32; ++w) {
        if (w < warp) base += warp_cnt[w];
        total += warp_cnt[w];
    }
    if (pass) {
        const size_t slot = ((size_t)b * p.NT + tile_id) * kPpTile + base + __popc(m & ((1u << lane) - 1u));
        p.ws.box[slot] = box;
        p.ws.score[slot] = conf;
        p.ws.meta[slot] = (anchor_base + src_t) | (cls << 24);
    }
    if (tid == 0) p.ws.tile_count[b * p.NT + tile_id] = total;
}

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

struct NmsParams {
    int B, NT, max_nms, max_det, flavor, agnostic, sort_cap;
    float thr_f;
    double thr_d;
    CandWs ws;
    float *dets;
    int32_t *counts;
    int32_t *keep_idx;
};

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const NmsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *keys = reinterpret_cast<unsigned long long *>(smem_raw);            // [sort_cap]
    float4 *kept_box = reinterpret_cast<float4 *>(keys + p.sort_cap);                         // [max_det]
    int *kept_cls = reinterpret_cast<int *>(kept_box + p.max_det);                            // [max_det]
    int *kept_slot = kept_cls + p.max_det;                                                    // [max_det]
    int *pref = kept_slot + p.max_det;                                                        // [NT+1]
    __shared__ float4 cbox[kRound];
    __shared__ int ccls[kRound];
    __shared__ int cidx[kRound];
    __shared__ unsigned cmask[kRound * (kRound / 32)];
    __shared__ int s_wcnt[kNmsWarps];
    __shared__ float red[kNmsWarps];
    __shared__ int s_nkept, s_total;

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NT = p.NT;
    const int *tcount = p.ws.tile_count + (size_t)b * NT;
    const size_t slot0 = (size_t)b * NT * kPpTile;

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
        if (lane == 0) s_nkept = 0;
    }
    __syncthreads();
    const int Nk = min(s_total, p.max_nms);  // postprocess.py:24-25 — first max_nms in anchor order
    int n_pad = 64;
    while (n_pad < Nk) n_pad <<= 1;

    // ---- keys + max coordinate (tv:ops/boxes.py:99 boxes.max())
    float mx = -FLT_MAX;
    for (int idx = tid; idx < NT * kPpTile; idx += kNmsThreads) {
        const int t = idx >> 7, j = idx & (kPpTile - 1);
        if (j < pref[t + 1] - pref[t]) {
            const int rank = pref[t] + j;
            if (rank < Nk) {
                const float sc = p.ws.score[slot0 + idx];
                keys[rank] = ((unsigned long long)(~float_ordered(sc)) << 32) | (unsigned)idx;
                const float4 bx = p.ws.box[slot0 + idx];
                mx = fmaxf(mx, fmaxf(fmaxf(bx.x, bx.y), fmaxf(bx.z, bx.w)));
            }
        }
    }
    for (int i = Nk + tid; i < n_pad; i += kNmsThreads) keys[i] = ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int w = 1; w < kNmsWarps; ++w) mx = fmaxf(mx, red[w]);
    const float span = mx + 1.0f;  // max_coordinate + 1 (tv:ops/boxes.py:100)

    // ---- bitonic sort ascending on (~score, slot): score descending, ties -> lower slot == stable.
    // Steps with partner distance <= 32 run in registers (each warp owns 64 consecutive keys, two per
    // lane, exchanged by shuffles); only distances >= 64 go through shared memory with a block barrier.
    auto local_steps = [&](const int k_lo, const int k_hi, const int j_hi) {
        // for every 64-key chunk: stages k = k_lo..k_hi (powers of two), steps j = min(k/2, j_hi)..1
        for (int chunk = warp; chunk < (n_pad >> 6); chunk += kNmsWarps) {
            const int i0 = (chunk << 6) + lane;
            unsigned long long a = keys[i0], c = keys[i0 + 32];
            for (int k = k_lo; k <= k_hi; k <<= 1) {
                const bool up = (i0 & k) == 0;  // same for i0 + 32 whenever k != 32 ... handled below
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

    // batched_nms branch (tv:ops/boxes.py:80): per-class loop vs coordinate trick
    const bool per_class = !p.agnostic && 4 * (long long)Nk > ((p.flavor & PLYOLO_NMS_RULE_CPU) ? 4000 : 100000);
    const bool use_off = !p.agnostic && !per_class;
    // the same-class shortcut needs offsets that dwarf their own rounding error (ulp(C*span) << 0.5)
    const bool filter_ok = span > 0.f && span * 256.f < 4.0e6f;

    // ---- greedy NMS in rounds of kRound candidates (score order), kSub threads per candidate:
    //   (A) test every candidate of the round against the kept list (<= max_det boxes in shared memory);
    //   (B) compact the survivors (usually a small fraction: dense clusters die against earlier keeps) and
    //       build the suppression bit-matrix among survivors only;
    //   (C) warp 0 sweeps the survivors sequentially over the remaining bits (ffs), appends the keeps,
    // and the loop exits as soon as max_det boxes are kept (output order == score order == sweep order).
    for (int base = 0; base < Nk; base += kRound) {
        const int nkept = s_nkept;
        if (nkept >= p.max_det) break;
        const int nch = min(kRound, Nk - base);
        const int ci = tid / kSub, sub = tid % kSub;
        // (A) every thread of a candidate fetches the same record (one broadcast request per candidate)
        bool sup = false;
        float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
        int cl = 0;
        if (ci < nch) {
            const unsigned slot = (unsigned)(keys[base + ci] & 0xffffffffu);
            bx = p.ws.box[slot0 + slot];
            cl = p.ws.meta[slot0 + slot] >> 24;
            if (use_off) {
                // A box of a HIGHER class can reach back into a lower class's offset range only if both its
                // x1 and y1 lie below -1 (+- rounding); everything else never overlaps another class.
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
                    kept_slot[dst] = (int)(keys[base + cidx[r]] & 0xffffffffu);
                }
                pos += __popc(kw);
            }
            if (lane == 0) s_nkept = nk;
        }
        __syncthreads();
    }

    // ---- output (postprocess.py:43-46): rows in score order, zero padded to max_det
    const int nkept = s_nkept;
    for (int i = tid; i < p.max_det; i += kNmsThreads) {
        float2 *d = reinterpret_cast<float2 *>(p.dets + ((size_t)b * p.max_det + i) * 6);
        if (i < nkept) {
            const int slot = kept_slot[i];
            const float4 bx = p.ws.box[slot0 + slot];
            const int meta = p.ws.meta[slot0 + slot];
            d[0] = make_float2(bx.x, bx.y);
            d[1] = make_float2(bx.z, bx.w);
            d[2] = make_float2(p.ws.score[slot0 + slot], (float)(meta >> 24));
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = meta & 0xffffff;
        } else {
            d[0] = make_float2(0.f, 0.f); d[1] = make_float2(0.f, 0.f); d[2] = make_float2(0.f, 0.f);
            if (p.keep_idx) p.keep_idx[(size_t)b * p.max_det + i] = -1;
        }
    }
    if (tid == 0) p.counts[b] = nkept;
}

static size_t cand_ws_layout(int B, int NT, CandWs *ws, unsigned char *base) {
    size_t off = 0;
    const size_t slots = (size_t)B * NT * kPpTile;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_cnt = take((size_t)B * NT * sizeof(int));
    size_t o_box = take(slots * sizeof(float4));
    size_t o_sc = take(slots * sizeof(float));
    size_t o_meta = take(slots * sizeof(int));
    if (ws) {
        ws->tile_count = reinterpret_cast<int *>(base + o_cnt);
        ws->box = reinterpret_cast<float4 *>(base + o_box);
        ws->score = reinterpret_cast<float *>(base + o_sc);
        ws->meta = reinterpret_cast<int *>(base + o_meta);
    }
    return off;
}

// worst-case tile count for A anchors split into at most PLYOLO_MAX_LEVELS levels
static int max_tiles(int A) { return (A + kPpTile - 1) / kPpTile + PLYOLO_MAX_LEVELS; }

static int run_nms(int B, int A, int NT, double nms_thre, int class_agnostic, int max_nms, int max_det, int flavor,
                   const CandWs &ws, float *dets, int32_t *counts, int32_t *keep_idx, cudaStream_t stream) {
    NmsParams np;
    np.B = B; np.NT = NT; np.max_nms = max_nms; np.max_det = max_det; np.flavor = flavor;
    np.agnostic = class_agnostic ? 1 : 0;
    np.thr_f = (float)nms_thre; np.thr_d = nms_thre;
    int cap = 64;
    const int need = max_nms < A ? max_nms : A;
    while (cap < need) cap <<= 1;
    np.sort_cap = cap;
    np.ws = ws; np.dets = dets; np.counts = counts; np.keep_idx = keep_idx;
    const size_t smem = (size_t)cap * 8 + (size_t)max_det * (sizeof(float4) + 2 * sizeof(int)) + (size_t)(NT + 1) * sizeof(int);
    PLYOLO_REQUIRE(smem <= 200 * 1024, "nms working set (%zu B) exceeds shared memory", smem);
    cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nms_kernel<<<B, kNmsThreads, smem, stream>>>(np);
    PLYOLO_CHECK_LAUNCH("nms_kernel");
    return PLYOLO_OK;
}

static int check_post_args(int B, int A, int C, int max_nms, int max_det, int flavor, float *dets, int32_t *counts,
                           void *workspace, size_t workspace_bytes) {
    PLYOLO_REQUIRE(B >= 1 && B <= 65535, "B=%d not in [1,65535]", B);
    PLYOLO_REQUIRE(A >= 1 && A < (1 << 24), "A=%d not in [1,2^24)", A);
    PLYOLO_REQUIRE(C >= 1 && C <= PLYOLO_MAX_CLASSES, "C=%d not in [1,%d]", C, PLYOLO_MAX_CLASSES);
    PLYOLO_REQUIRE(max_det >= 1 && max_det <= 1024, "max_det=%d not in [1,1024]", max_det);
    PLYOLO_REQUIRE(max_nms >= 1, "max_nms=%d must be positive", max_nms);
    PLYOLO_REQUIRE((max_nms < A ? max_nms : A) <= kMaxSortCap, "min(max_nms, A)=%d exceeds %d", max_nms < A ? max_nms : A,
                   kMaxSortCap);
    PLYOLO_REQUIRE(flavor >= 0 && flavor <= 7, "flavor=%d not in [0,7]", flavor);
    PLYOLO_REQUIRE(dets && counts, "dets / counts is null");
    if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < plyolo_postprocess_workspace_bytes(B, A)) {
        set_error("workspace null, not 256-byte aligned, or smaller than plyolo_postprocess_workspace_bytes()");
        return PLYOLO_ERR_WORKSPACE;
    }
    return PLYOLO_OK;
}

}  // namespace plyolo

extern "C" size_t plyolo_postprocess_workspace_bytes(int B, int A) {
    if (B < 1 || A < 1) return 0;
    return plyolo::cand_ws_layout(B, plyolo::max_tiles(A), nullptr, nullptr);
}

extern "C" int plyolo_postprocess_f32(const float *preds, int B, int A, int C, double conf_thre, double nms_thre,
                                      int class_agnostic, int max_nms, int max_det, int flavor, float *dets,
                                      int32_t *counts, int32_t *keep_idx, void *workspace, size_t workspace_bytes,
                                      plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(preds != nullptr, "preds is null");
    int rc = check_post_args(B, A, C, max_nms, max_det, flavor, dets, counts, workspace, workspace_bytes);
    if (rc != PLYOLO_OK) return rc;
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    ScoreParams sp;
    sp.preds = preds; sp.B = B; sp.A = A; sp.C = C; sp.ch = 5 + C;
    sp.NT = (A + kPpTile - 1) / kPpTile;
    sp.conf_thr = (float)conf_thre;  // `tensor >= python float` compares in fp32
    sp.bulk_ok = (((uintptr_t)preds & 15) == 0 && (A & 3) == 0) ? 1 : 0;
    sp.lv.n = 0; sp.lv.A = A;
    cand_ws_layout(B, sp.NT, &sp.ws, static_cast<unsigned char *>(workspace));
    const size_t smem = (size_t)kPpTile * sp.ch * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(score_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    score_kernel<false><<<dim3(sp.NT, B), kPpTile, smem, (cudaStream_t)stream>>>(sp);
    PLYOLO_CHECK_LAUNCH("score_kernel<preds>");
    return run_nms(B, A, sp.NT, nms_thre, class_agnostic, max_nms, max_det, flavor, sp.ws, dets, counts, keep_idx,
                   (cudaStream_t)stream);
}

extern "C" int plyolo_decode_postprocess_f32(const float *const *host_lvl, const int *hs, const int *ws,
                                             const int *strides, int n_levels, int B, int C, double conf_thre,
                                             double nms_thre, int class_agnostic, int max_nms, int max_det,
                                             int flavor, float *dets, int32_t *counts, int32_t *keep_idx,
                                             void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    using namespace plyolo;
    ScoreParams sp;
    int rc = make_levels(sp.lv, host_lvl, hs, ws, strides, n_levels, kPpTile);
    if (rc != PLYOLO_OK) return rc;
    const int A = sp.lv.A;
    rc = check_post_args(B, A, C, max_nms, max_det, flavor, dets, counts, workspace, workspace_bytes);
    if (rc != PLYOLO_OK) return rc;
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    sp.preds = nullptr; sp.B = B; sp.A = A; sp.C = C; sp.ch = 5 + C;
    sp.NT = sp.lv.tile0[sp.lv.n];
    sp.conf_thr = (float)conf_thre;
    bool bulk = true;
    for (int l = 0; l < sp.lv.n; ++l) bulk = bulk && ((uintptr_t)sp.lv.ptr[l] & 15) == 0 && (sp.lv.hw[l] & 3) == 0;
    sp.bulk_ok = bulk ? 1 : 0;
    cand_ws_layout(B, sp.NT, &sp.ws, static_cast<unsigned char *>(workspace));
    const size_t smem = (size_t)kPpTile * sp.ch * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(score_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    score_kernel<true><<<dim3(sp.NT, B), kPpTile, smem, (cudaStream_t)stream>>>(sp);
    PLYOLO_CHECK_LAUNCH("score_kernel<fused>");
    return run_nms(B, A, sp.NT, nms_thre, class_agnostic, max_nms, max_det, flavor, sp.ws, dets, counts, keep_idx,
                   (cudaStream_t)stream);
}
