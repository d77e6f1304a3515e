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
#include "common.cuh"

namespace plyolo {

constexpr int kPpTile = 128;
constexpr int kGroups = 4;      // class groups per image (class & 3): one NMS CTA each
constexpr int kMaxCross = 512;  // boxes that may reach into another class's offset range (x1, y1 < -0.5)
constexpr int kFastCap = 4096;  // candidates per NMS CTA whose boxes are staged in shared memory
constexpr int kImgCtr = 8;      // ints per image in the counter block (zeroed before every call)

constexpr int kRecCap = 1024;   // >= max_det: a group contributes at most max_det rows to the image's output

struct __align__(16) KeptRec {  // one kept box as the merge needs it
    unsigned long long key;     // ~ordered(score) << 25 | slot: the global order
    float score;
    int meta;                   // anchor | class << 24
    float4 box;
};

struct CandWs {
    int *tile_count;    // [B, NT]
    float4 *box;        // [B, NT*128]  original (un-offset) corners
    float *score;       // [B, NT*128]
    int *meta;          // [B, NT*128]  anchor | class << 24
    // per image: candidates bucketed by class group, in arrival order (the keys carry the anchor order)
    int *ctr;                     // [B, kImgCtr]  gcount[kGroups] | max coordinate (ordered uint) | #cross | #done CTAs | fallback
    unsigned long long *gkey;     // [B, kGroups, NT*128]  class << 57 | ~ordered(score) << 25 | slot
    unsigned long long *xkey;     // [B, kMaxCross] keys of the cross boxes
    float4 *xbox;                 // [B, kMaxCross]
    KeptRec *krec;                // [B, kGroups, kRecCap] first kept boxes of every group, global order
    int *kcount;                  // [B, kGroups]
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
    const int islot = tile_id * kPpTile + base + __popc(m & ((1u << lane) - 1u));  // slot inside the image
    if (pass) {
        const size_t slot = (size_t)b * p.NT * kPpTile + islot;
        p.ws.box[slot] = box;
        p.ws.score[slot] = conf;
        p.ws.meta[slot] = (anchor_base + src_t) | (cls << 24);
    }
    if (tid == 0) p.ws.tile_count[b * p.NT + tile_id] = total;

    // ---- class-group buckets for the NMS stage (one CTA per image and group): keys in arrival order, the
    // image's max coordinate (tv:ops/boxes.py:99) and the boxes that can reach another class's offset range
    if (total == 0) return;
    __shared__ int g_wcnt[kPpTile / 32][kGroups];
    __shared__ int g_base[kGroups];
    __shared__ float w_max[kPpTile / 32];
    const int grp = cls & (kGroups - 1);
    unsigned gm = 0u;
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const unsigned mg = __ballot_sync(0xffffffffu, pass && grp == g);
        if (lane == 0) g_wcnt[warp][g] = __popc(mg);
        if (grp == g) gm = mg;
    }
    float cm = pass ? fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)) : -3.0e38f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cm = fmaxf(cm, __shfl_xor_sync(0xffffffffu, cm, o));
    if (lane == 0) w_max[warp] = cm;
    __syncthreads();
    int *ctr = p.ws.ctr + b * kImgCtr;
    if (tid < kGroups) {
        int n = 0;
#pragma unroll
        for (int w = 0; w < kPpTile / 32; ++w) n += g_wcnt[w][tid];
        g_base[tid] = n ? atomicAdd(&ctr[tid], n) : 0;
    } else if (tid == 32) {
        float mx = w_max[0];
#pragma unroll
        for (int w = 1; w < kPpTile / 32; ++w) mx = fmaxf(mx, w_max[w]);
        atomicMax(reinterpret_cast<unsigned *>(&ctr[kGroups]), float_ordered(mx));
    }
    __syncthreads();
    if (pass) {
        int pos = g_base[grp] + __popc(gm & ((1u << lane) - 1u));
        for (int w = 0; w < warp; ++w) pos += g_wcnt[w][grp];
        const unsigned long long key = ((unsigned long long)cls << 57) |
                                       ((unsigned long long)(~float_ordered(conf)) << 25) | (unsigned)islot;
        p.ws.gkey[((size_t)b * kGroups + grp) * ((size_t)p.NT * kPpTile) + pos] = key;
        if (box.x < -0.5f && box.y < -0.5f) {
            const int xi = atomicAdd(&ctr[kGroups + 1], 1);
            if (xi < kMaxCross) {
                p.ws.xkey[(size_t)b * kMaxCross + xi] = key;
                p.ws.xbox[(size_t)b * kMaxCross + xi] = box;
            }
        }
    }
}

}  // namespace plyolo

#include "nms.cuh"

namespace plyolo {

static thread_local long long *g_nms_prof = nullptr;

static size_t cand_ws_layout(int B, int NT, CandWs *ws, unsigned char *base) {
    size_t off = 0;
    const size_t slots = (size_t)B * NT * kPpTile;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_ctr = take((size_t)B * kImgCtr * sizeof(int));  // first: zeroed before every call
    size_t o_cnt = take((size_t)B * NT * sizeof(int));
    size_t o_box = take(slots * sizeof(float4));
    size_t o_sc = take(slots * sizeof(float));
    size_t o_meta = take(slots * sizeof(int));
    size_t o_gkey = take(slots * kGroups * sizeof(unsigned long long));
    size_t o_xkey = take((size_t)B * kMaxCross * sizeof(unsigned long long));
    size_t o_xbox = take((size_t)B * kMaxCross * sizeof(float4));
    size_t o_kept = take((size_t)B * kGroups * kRecCap * sizeof(KeptRec));
    size_t o_kcnt = take((size_t)B * kGroups * sizeof(int));
    if (ws) {
        ws->ctr = reinterpret_cast<int *>(base + o_ctr);
        ws->tile_count = reinterpret_cast<int *>(base + o_cnt);
        ws->box = reinterpret_cast<float4 *>(base + o_box);
        ws->score = reinterpret_cast<float *>(base + o_sc);
        ws->meta = reinterpret_cast<int *>(base + o_meta);
        ws->gkey = reinterpret_cast<unsigned long long *>(base + o_gkey);
        ws->xkey = reinterpret_cast<unsigned long long *>(base + o_xkey);
        ws->xbox = reinterpret_cast<float4 *>(base + o_xbox);
        ws->krec = reinterpret_cast<KeptRec *>(base + o_kept);
        ws->kcount = reinterpret_cast<int *>(base + o_kcnt);
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
    np.fast_cap = cap < kFastCap ? cap : kFastCap;
    np.ws = ws; np.dets = dets; np.counts = counts; np.keep_idx = keep_idx;
    np.prof = g_nms_prof;
    const size_t smem = nms_group_smem_bytes(cap, np.fast_cap, max_det, NT);
    PLYOLO_REQUIRE(smem <= 190 * 1024, "nms working set (%zu B) exceeds shared memory", smem);
    cudaFuncSetAttribute(nms_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    nms_group_kernel<<<dim3(kGroups, B), kNmsThreads, smem, stream>>>(np);
    PLYOLO_CHECK_LAUNCH("nms_group_kernel");
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

// debug hook (not part of include/plyolo.h): device buffer [B][16] of int64 receiving the NMS kernel's phase
// timestamps for the calling thread's next launches; null switches it off
extern "C" void plyolo_debug_nms_profile(void *device_buf) { plyolo::g_nms_prof = static_cast<long long *>(device_buf); }

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
    if (cudaMemsetAsync(sp.ws.ctr, 0, (size_t)B * kImgCtr * sizeof(int), (cudaStream_t)stream) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
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
    if (cudaMemsetAsync(sp.ws.ctr, 0, (size_t)B * kImgCtr * sizeof(int), (cudaStream_t)stream) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    score_kernel<true><<<dim3(sp.NT, B), kPpTile, smem, (cudaStream_t)stream>>>(sp);
    PLYOLO_CHECK_LAUNCH("score_kernel<fused>");
    return run_nms(B, A, sp.NT, nms_thre, class_agnostic, max_nms, max_det, flavor, sp.ws, dets, counts, keep_idx,
                   (cudaStream_t)stream);
}
