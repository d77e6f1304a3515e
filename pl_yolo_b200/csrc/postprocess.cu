// postprocess.cu — confidence filter + class-aware NMS (models/evaluators/postprocess.py:7-48,
// tv:ops/boxes.py:51-120, torchvision::nms), from materialised predictions or fused with the decode.
//
// Stage 1  score_kernel<FUSED>   persistent, one CTA per SM, HBM-bound.
//   A producer warp streams tiles of 128 anchors into a 4-stage shared-memory ring through the TMA engine
//   (FUSED: one 2-D tensor-map request per tile = 5+C channel-plane rows of 512 B -> tile[channel][anchor], L2
//   evict-first; preds: one contiguous bulk copy -> tile[anchor][channel]); four consumer groups of 128 threads
//   score them (thread per anchor): sigmoid(obj); the largest raw class logit from three-input maxima over chunks
//   of 16 classes; a conservative pre-filter that can only over-admit; for the admitted anchors the exact max /
//   first argmax over the sigmoid VALUES (T1) inside the rounding window of the fp32 sigmoid;
//   conf = sigmoid(obj) * class_conf, conf >= thr in fp32.  The survivors are compacted IN ANCHOR ORDER (warp
//   ballot + prefix, the tile's ONE group barrier) into the tile's slot range of the candidate arrays and bucketed by
//   class group (per-warp atomics, issued right after the sweep).  Every warp releases the stage on its own; a
//   publisher warp releases the per-image scored-tile counters at GPU scope for stage 2.
// Stage 2  nms_fast_kernel       one 4-CTA cluster per image (nms_fast.cuh), launched with programmatic dependent
//   launch so that it runs UNDER stage 1: small CTAs (512 threads, 52 KB) co-resident with the score CTAs, every
//   cluster starts when its image's scored-tile counter is complete.  Per-class pre-kill, sorts and greedy sweeps
//   with torchvision's arithmetic (coordinate-trick offsets, asymmetric FMA, IEEE division; see `suppresses`),
//   kept lists merged by rank through distributed shared memory.
// Stage 3  nms_general_kernel    one CTA per image (nms.cuh: global sort + bit-matrix rounds), only for the images
//   the class split cannot handle exactly: launched from the device by stage 2 for those images (plain stream launches)
//   or by the host behind stage 2 under stream capture (flag per image; exits at once otherwise).
// Several batches can be in flight on different streams (pl_yolo_b200/pipeline.py): the workspace is per call, and
// stage 1 / stage 2 CTAs of different batches share the SMs the same way.
#include <cstdlib>
#include <cstring>
#include <mutex>

#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

#include "common.cuh"

namespace plyolo {

constexpr int kPpTile = 128;
constexpr int kGroups = 4;      // class groups per image (class & 3): one NMS CTA each
constexpr int kMaxCross = 512;  // boxes that may reach into another class's offset range (x1, y1 < -0.5)
constexpr int kFastCap = 4096;  // candidates per NMS CTA whose boxes are staged in shared memory
constexpr int kImgCtr = 16;     // ints per image in the counter block (zeroed before every call)
// counter block of an image: [0, kGroups) candidates per class group, then
constexpr int kCtrMaxCoord = 8;  // largest coordinate (ordered uint): tv:ops/boxes.py:99
constexpr int kCtrCross = 9;     // boxes listed in the cross list
constexpr int kCtrGeneral = 10;  // the image must be redone by the general path
constexpr int kCtrDone = 11;     // tiles of the image that have been scored (released by the score kernel)
constexpr int kCtrMaxX2 = 12;    // largest x2 / y2 of the image's candidates (ordered uint): bounds which cross boxes
constexpr int kCtrMaxY2 = 13;    //   can reach another class's offset range at all
constexpr int kBucketCap = 2048; // bucket entries per (image, class group); the NMS kernel stages the first 1536 of them
constexpr int kBucketImg = kGroups * kBucketCap;

constexpr size_t kNmsSmemLimit = 190 * 1024;  // dynamic shared memory of the NMS kernels (33 KB are static)
struct CandWs {
    int *tile_count;    // [B, NT]
    float4 *box;        // [B, NT*128]  original (un-offset) corners
    float *score;       // [B, NT*128]
    int *meta;          // [B, NT*128]  anchor | class << 24
    float *aux;         // [B, NT*128]  second score of the YOLOv3 / YOLOv5 call sites (objectness / best class score)
    // per image: candidates bucketed by class group, in arrival order (the keys carry the anchor order)
    int *ctr;                     // [B, kImgCtr]  counter block (kCtr*)
    unsigned long long *gkey;     // [B, kGroups, kBucketCap]  class << 57 | ~ordered(score) << 25 | anchor
    float4 *gbox;                 // [B, kGroups, kBucketCap]  the same candidates' corners
    unsigned long long *xkey;     // [B, kMaxCross] keys of the cross boxes
    float4 *xbox;                 // [B, kMaxCross]
};

struct ScoreParams {
    Levels lv;           // FUSED only
    const float *preds;  // !FUSED only
    int B, A, C, ch, NT;
    float conf_thr;
    int bulk_ok;
    int variant;  // PLYOLO_NMS_YOLOX / _YOLOV3 / _YOLOV5: which reference call site's filter + score arithmetic (preds input only)
    CandWs ws;
    long long *prof;  // debug: [gridDim.x][kConsumers][8] accumulated cycles per consumer phase, or null
};

// per consumer group (128 threads = one tile at a time) scratch
struct TileShared {
    int warp_cnt[2][kPpTile / 32];  // candidates per warp, by tile parity (one group barrier per tile: a warp may already
                                    // be writing the next tile's counts while a slower one still reads these)
    unsigned done;                  // warps of the group that have issued every record of a tile: 4 per tile, monotonic
    unsigned pad[7];
};

struct TileCoord {
    int b, tile_id, l, a0, cnt, anchor_base;
    const float *src;  // FUSED: first anchor of channel 0 of the tile; else first row of the tile in preds
};

template <bool FUSED>
__device__ __forceinline__ TileCoord tile_coord(const ScoreParams &p, const int b, const int tile_id) {
    TileCoord t;
    t.b = b; t.tile_id = tile_id; t.l = 0;
    if (FUSED) {
#pragma unroll
        for (int i = 1; i < PLYOLO_MAX_LEVELS; ++i)
            if (i < p.lv.n && tile_id >= p.lv.tile0[i]) t.l = i;
        t.a0 = (tile_id - p.lv.tile0[t.l]) * kPpTile;
        t.cnt = min(kPpTile, p.lv.hw[t.l] - t.a0);
        t.anchor_base = p.lv.off[t.l] + t.a0;
        t.src = p.lv.ptr[t.l] + (size_t)b * p.ch * p.lv.hw[t.l] + t.a0;
    } else {
        t.a0 = tile_id * kPpTile;
        t.cnt = min(kPpTile, p.A - t.a0);
        t.anchor_base = t.a0;
        t.src = p.preds + ((size_t)b * p.A + t.a0) * p.ch;
    }
    return t;
}

__device__ __forceinline__ void group_barrier(const int bar_id) {
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
}

// three-input maximum (FMNMX3, sm_100+); NaN operands are ignored like fmaxf
__device__ __forceinline__ float max3f(const float a, const float b, const float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// Scores one staged tile (128 threads, `tid` in [0,128), barrier `bar_id`).  `tile` is read-only here; after
// the call returns the group no longer needs it IF `release` was invoked (it is called once, by every thread,
// right after the last read of the tile).
#define TPROF(slot)                                             \
    do {                                                        \
        if (acc && tid == 0) {                                  \
            const long long now__ = clock64();                  \
            acc[slot] += now__ - tlast;                         \
            tlast = now__;                                      \
        }                                                       \
    } while (0)

template <bool FUSED, typename Release>
__device__ __forceinline__ void score_tile(const ScoreParams &p, const float *tile, const TileCoord tc, const int tid,
                                           const int bar_id, TileShared &sh, const int par, Release release,
                                           long long *acc = nullptr, long long tlast = 0) {
    const int ch = p.ch, b = tc.b, l = tc.l, a0 = tc.a0, cnt = tc.cnt;
    const int lane = tid & 31, warp = tid >> 5;
    bool pass = false;
    float conf = 0.f, aux = 0.f;
    int cls = 0;
    float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
    const int src_t = tid;  // tile-local anchor of this thread

    if (FUSED) {
        // One thread per anchor.  conf = fl(sigmoid(obj) * max_c sigmoid(x_c)) and the FIRST class attaining the
        // max over the sigmoid VALUES (postprocess.py:18 works on the already-squashed tensor; T1).
        //  1. so = sigmoid(obj); so < thr cannot pass (class_conf <= 1, fp32 multiply is monotone);
        //  2. m = max raw logit (4 independent chains); conservative pre-filter on fl(so * sigmoid(m)) * 1.00001;
        //  3. exact argmax: the fp32 sigmoid is monotone only up to rounding (relative error < 4e-7), so every
        //     class whose logit is >= t can still tie with or beat sigmoid(m), where
        //     t = m - 2.5e-6 (1 + e^m) - margin  (sigmoid(t) <= sigmoid(m) (1 - 2.5e-6), see below);
        //     only those (normally just the arg max; many when the logits saturate) are evaluated exactly.
        if (tid < cnt) {
            const float so = sigmoid_ref(tile[4 * kPpTile + tid]);  // yolox_loss.py:26
            if (so >= p.conf_thr) {
                // Stage 1 (every thread that can still pass): the largest raw logit of every chunk of 16 classes
                // with three-input max (FMNMX3): 1.5 instructions per class.  Stage 2 (only the anchors that
                // survive the pre-filter, i.e. real candidates): the chunks whose maximum reaches the window are
                // rescanned in ascending class order for the first arg max over the sigmoid VALUES.
                constexpr int CS = 16, NCK = (PLYOLO_MAX_CLASSES + CS - 1) / CS;
                const float *col = tile + 5 * kPpTile + tid;
                float M[NCK];
#pragma unroll
                for (int k = 0; k < NCK; ++k) {
                    M[k] = -3.0e38f;
                    if (k * CS < p.C) {
                        const float *ck = col + k * CS * kPpTile;
                        if ((k + 1) * CS <= p.C) {
                            float x[CS];
#pragma unroll
                            for (int u = 0; u < CS; ++u) x[u] = ck[u * kPpTile];
                            const float q0 = max3f(x[0], x[1], x[2]), q1 = max3f(x[3], x[4], x[5]);
                            const float q2 = max3f(x[6], x[7], x[8]), q3 = max3f(x[9], x[10], x[11]);
                            const float q4 = max3f(x[12], x[13], x[14]);
                            M[k] = max3f(max3f(q0, q1, q2), max3f(q3, q4, x[15]), -3.0e38f);
                        } else {
                            for (int cc = k * CS; cc < p.C; ++cc) M[k] = fmaxf(M[k], ck[(cc - k * CS) * kPpTile]);
                        }
                    }
                }
                float m = M[0];
#pragma unroll
                for (int k = 1; k < NCK; ++k) m = fmaxf(m, M[k]);
                const float em = expf(-m);
                const float sm = 1.0f / (1.0f + em);  // sigmoid_ref(m): yolox_loss.py:27
                if ((so * sm) * 1.00001f >= p.conf_thr) {
                    // window: d/dx ln sigmoid(x) = 1 / (1 + e^x) >= 1 / (1 + e^m) below m, so a logit under
                    // m - 2.5e-6 (1 + e^m) has a sigmoid under sigmoid(m) (1 - 2.5e-6): it cannot tie with or beat
                    // the maximum within the rounding of the fp32 sigmoid.  Saturated (e^-m = 0) or vanishing
                    // (denormal) maxima take every class.
                    float t = m - 2.5e-6f * (1.0f + __frcp_rn(em));
                    t = t - 1.0e-5f * (1.0f + fabsf(t));
                    if (!(t <= m) || sm < 1.0e-30f) t = -__int_as_float(0x7f800000);
                    // Normally only the arg max lies inside the window: ONE chunk reaches t and one logit of it.
                    // That chunk is rescanned branch-free (every lane its own chunk); anything else (saturation,
                    // near ties) takes the exact sweep over the sigmoid values.
                    int kstar = 0, nhit = 0;
#pragma unroll
                    for (int k = NCK - 1; k >= 0; --k) {
                        const bool h = k * CS < p.C && M[k] >= t;
                        kstar = h ? k : kstar;
                        nhit += h ? 1 : 0;
                    }
                    const float *ck = col + kstar * CS * kPpTile;
                    unsigned hits = 0u;
                    if ((kstar + 1) * CS <= p.C) {
#pragma unroll
                        for (int u = 0; u < CS; ++u) hits |= (ck[u * kPpTile] >= t) ? (1u << u) : 0u;
                    } else {
#pragma unroll
                        for (int u = 0; u < CS; ++u) {
                            const float x = (kstar * CS + u < p.C) ? ck[u * kPpTile] : -3.0e38f;
                            hits |= (x >= t) ? (1u << u) : 0u;
                        }
                    }
                    float best = sm;
                    cls = kstar * CS + __ffs(hits) - 1;
                    if (nhit != 1 || __popc(hits) != 1) {  // rare: several logits inside the window
                        best = -1.f;
                        cls = 0;
                        for (int cc = 0; cc < p.C; ++cc) {
                            const float x = col[cc * kPpTile];
                            if (x >= t) {
                                const float v = sigmoid_ref(x);
                                if (v > best) { best = v; cls = cc; }
                            }
                        }
                    }
                    conf = so * best;           // postprocess.py:19
                    pass = conf >= p.conf_thr;  // :20 (fp32 compare)
                    if (pass) {
                        const int a = a0 + tid;
                        const int W = p.lv.w[l];
                        const float s = p.lv.stride[l];
                        // row = a / W: exact from the fp32 product for a < 2^21 ((a + 0.5) / W is at least 0.5 / W away
                        // from an integer, the product is off by < 2e-7 * a / W)
                        const int gy = p.lv.hw[l] <= (1 << 21) ? __float2int_rz(((float)a + 0.5f) * p.lv.inv_w[l]) : a / W;
                        const int gx = a - gy * W;
                        const float cx = (tile[0 * kPpTile + tid] + (float)gx) * s;  // yolox_loss.py:217
                        const float cy = (tile[1 * kPpTile + tid] + (float)gy) * s;
                        const float w = expf(tile[2 * kPpTile + tid]) * s;                // :219
                        const float h = expf(tile[3 * kPpTile + tid]) * s;
                        box = make_float4(cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2);  // :31-34
                    }
                }
            }
        }
    } else if (tid < cnt) {
        const float *r = tile + tid * ch;
        if (p.variant == PLYOLO_NMS_YOLOX) {
            float best = r[5];
            for (int c = 1; c < p.C; ++c) {
                const float s = r[5 + c];
                if (s > best) { best = s; cls = c; }  // postprocess.py:18, first max index
            }
            conf = r[4] * best;
            pass = conf >= p.conf_thr;
            box = make_float4(r[0], r[1], r[2], r[3]);
        } else {
            // YOLOv3 / YOLOv5 decoders: rows are (cx, cy, w, h, obj, cls..) already squashed
            const float obj = r[4];
            if (obj > p.conf_thr) {  // yolov3_decoder.py:74, yolov5_decoder.py:23 (strict)
                if (p.variant == PLYOLO_NMS_YOLOV3) {
                    float best = r[5] * obj;  // :79 cls *= obj, :86 max over the products (first max index)
                    for (int c = 1; c < p.C; ++c) {
                        const float s = r[5 + c] * obj;
                        if (s > best) { best = s; cls = c; }
                    }
                    conf = best;               // the NMS score (:104)
                    aux = obj;
                    pass = conf > p.conf_thr;  // :87 (strict)
                } else {
                    float best = r[5];         // yolov5_decoder.py:57 max over the raw class scores
                    for (int c = 1; c < p.C; ++c) {
                        const float s = r[5 + c];
                        if (s > best) { best = s; cls = c; }
                    }
                    pass = (obj * best) >= p.conf_thr;  // :58
                    conf = obj;                // the NMS score is the objectness (:71)
                    aux = best;
                }
                // box_corner (yolov3_decoder.py:64-68) / xywh2xyxy (models/utils/bbox.py:5-12)
                box = make_float4(r[0] - r[2] / 2, r[1] - r[3] / 2, r[0] + r[2] / 2, r[1] + r[3] / 2);
            }
        }
    }
    TPROF(2);
    release();  // last read of the tile is done (every thread calls it)

    // ---- per warp, no barrier: the warp's candidates per class group take their place in the image's buckets (arrival
    // order: the keys carry the anchor order) and the warp's largest x2 / y2 go into the image's maxima.  The atomics are
    // issued here, their results are used after the slot records below: the L2 round trip is hidden.
    const unsigned m = __ballot_sync(0xffffffffu, pass);
    const int grp = cls & (kGroups - 1);
    int *ctr = p.ws.ctr + b * kImgCtr;
    unsigned gm = 0u;
    int gbase = 0;  // lane g: where the warp's candidates of class group g start in the bucket
    if (m) {        // uniform for the warp
        int gcnt = 0;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const unsigned mg = __ballot_sync(0xffffffffu, pass && grp == g);
            if (lane == g) gcnt = __popc(mg);
            if (grp == g) gm = mg;
        }
        if (gcnt) gbase = atomicAdd(&ctr[lane], gcnt);  // lanes < kGroups only
        // maxima as ordered uints (NaN coordinates are ignored like fmaxf does); max over all four coordinates
        // (tv:ops/boxes.py:99): the fused decode gives x2 >= x1, y2 >= y1 (w, h = exp(..) * stride)
        unsigned oz = float_ordered(pass ? fmaxf(box.z, -3.0e38f) : -3.0e38f);
        unsigned ow = float_ordered(pass ? fmaxf(box.w, -3.0e38f) : -3.0e38f);
        oz = __reduce_max_sync(0xffffffffu, oz);
        ow = __reduce_max_sync(0xffffffffu, ow);
        unsigned oc = oz > ow ? oz : ow;
        if (!FUSED) {
            const unsigned oxy = float_ordered(pass ? fmaxf(fmaxf(box.x, box.y), -3.0e38f) : -3.0e38f);
            const unsigned r = __reduce_max_sync(0xffffffffu, oxy);
            oc = oc > r ? oc : r;
        }
        if (lane >= 8 && lane < 11)
            atomicMax(reinterpret_cast<unsigned *>(&ctr[lane == 8 ? kCtrMaxCoord : (lane == 9 ? kCtrMaxX2 : kCtrMaxY2)]),
                      lane == 8 ? oc : (lane == 9 ? oz : ow));
    }

    // ---- order-preserving compaction into the tile's slots (postprocess.py:23 keeps anchor order): the one group
    // barrier of the tile
    if (lane == 0) sh.warp_cnt[par][warp] = __popc(m);
    group_barrier(bar_id);
    TPROF(4);
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kPpTile / 32; ++w) {
        const int c = sh.warp_cnt[par][w];
        if (w < warp) base += c;
        total += c;
    }
    const int islot = tc.tile_id * kPpTile + base + __popc(m & ((1u << lane) - 1u));  // slot inside the image
    if (pass) {
        const size_t slot = (size_t)b * p.NT * kPpTile + islot;
        p.ws.box[slot] = box;
        p.ws.score[slot] = conf;
        p.ws.meta[slot] = (tc.anchor_base + src_t) | (cls << 24);
        if (!FUSED && p.variant != PLYOLO_NMS_YOLOX) p.ws.aux[slot] = aux;
    }
    if (tid == 0) p.ws.tile_count[b * p.NT + tc.tile_id] = total;
    TPROF(5);

    // ---- class-group buckets for the NMS stage (one CTA per image and group) and the boxes that can reach another
    // class's offset range
    if (m) {
        const int pos = __shfl_sync(0xffffffffu, gbase, grp) + __popc(gm & ((1u << lane) - 1u));
        if (pass) {
            // the low bits order equal scores by anchor (ascending anchor == the reference's candidate order)
            const unsigned long long key = ((unsigned long long)cls << 57) |
                                           ((unsigned long long)(~float_ordered(conf)) << 25) | (unsigned)(tc.anchor_base + src_t);
            if (pos < kBucketCap) {  // a fuller group is redone by the general path from the slot arrays
                p.ws.gkey[(size_t)b * kBucketImg + grp * kBucketCap + pos] = key;
                p.ws.gbox[(size_t)b * kBucketImg + grp * kBucketCap + pos] = box;
            }
            if (box.x < -0.5f && box.y < -0.5f) {
                const int xi = atomicAdd(&ctr[kCtrCross], 1);
                if (xi < kMaxCross) {
                    p.ws.xkey[(size_t)b * kMaxCross + xi] = key;
                    p.ws.xbox[(size_t)b * kMaxCross + xi] = box;
                }
            }
        }
    }
    TPROF(6);
}

// ---- one CTA per tile, plain loads: the fallback for unaligned inputs (no 16-byte alignment for bulk copies)
template <bool FUSED>
__global__ void __launch_bounds__(kPpTile) score_kernel_simple(const ScoreParams p) {
    extern __shared__ __align__(128) float tile[];
    __shared__ TileShared sh;
    const int tid = threadIdx.x;
    const TileCoord tc = tile_coord<FUSED>(p, blockIdx.y, blockIdx.x);
    if (FUSED) {
        for (int c = 0; c < p.ch; ++c)
            if (tid < tc.cnt) tile[c * kPpTile + tid] = __ldg(tc.src + (size_t)c * p.lv.hw[tc.l] + tid);
    } else {
        for (int i = tid; i < tc.cnt * p.ch; i += kPpTile) tile[i] = __ldg(tc.src + i);
    }
    __syncthreads();
    score_tile<FUSED>(p, tile, tc, tid, 0, sh, 0, [] {});
}

// ---- persistent, TMA-pipelined score kernel ---------------------------------------------------------
// One CTA per SM walks the (image, tile) list with stride gridDim.x.  Warp 0 is the producer: for every
// tile it waits for a free stage and copies the tile asynchronously onto the stage's mbarrier — from the
// channel-planar head maps with 16-byte cp.async (one 512 B row per warp instruction: the TMA engine's
// per-request service time caps 512 B bulk copies near 3 TB/s chip-wide, measured), from preds with one
// contiguous TMA bulk copy of the whole tile.  kConsumers consumer
// groups of 128 threads take the tiles round-robin, so the scoring latency of one tile (dependent max
// chains, barriers, the bucket atomics) overlaps the next tiles' loads and compute.
constexpr int kStages = 4;
#ifndef PLYOLO_SCORE_CONSUMERS
#define PLYOLO_SCORE_CONSUMERS 4
#endif
#ifndef PLYOLO_SCORE_PRODUCERS
#define PLYOLO_SCORE_PRODUCERS 2
#endif
constexpr int kConsumers = PLYOLO_SCORE_CONSUMERS;
static_assert(kConsumers == kStages, "only the 4-group / 4-stage layout is validated (3 groups fault, 1 producer warp gains nothing)");
constexpr int kProducers = PLYOLO_SCORE_PRODUCERS;  // producer warps (cp.async fallback: each copies part of the channel rows)
constexpr int kScoreThreads = 32 * (kProducers + 1) + kPpTile * kConsumers;  // producers | publisher warp | consumer groups

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and its completion hook: the mbarrier
// receives one (pre-counted) arrival from this thread once all of its earlier cp.async have landed
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one 2-D tensor map per FPN level: dim0 = the H*W anchors of a channel plane (contiguous), dim1 = the
// B * (5+C) planes; a tile is the box (128 anchors, 5+C planes) at (a0, b * (5+C)) — ONE request to the TMA
// engine per tile; anchors past the end of the level are zero-filled
struct TmapPack {
    CUtensorMap m[PLYOLO_MAX_LEVELS];
};

// The head maps are read exactly once: an evict-first L2 policy keeps the 91 MB stream from pushing the
// candidate records (written by this kernel, read back by the NMS kernel a few microseconds later) out to DRAM.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *map, const int c0, const int c1, uint64_t *bar,
                                            const uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
            smem_u32(dst_smem)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

// the scored-tile counter of an image: release at GPU scope (the tile's candidate records were written by the whole
// group before the group barrier that precedes this call); nms_fast_kernel acquires it
__device__ __forceinline__ void publish_tile(int *ctr) {
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr + kCtrDone) : "memory");
}

// kScoreRegs registers per thread leave room on the SM for the co-resident NMS CTA (budget: see nms_fast.cuh)
#ifndef PLYOLO_SCORE_REGS
#define PLYOLO_SCORE_REGS 64
#endif
constexpr int kScoreRegs = PLYOLO_SCORE_REGS;

template <bool FUSED>
__global__ void __maxnreg__(kScoreRegs)
score_kernel(const ScoreParams p, const __grid_constant__ TmapPack tmaps, const int use_tmap) {
    extern __shared__ __align__(128) float stages[];  // [kStages][ch * 128]
    // programmatic dependent launch: the NMS grid may be scheduled from now on (its clusters wait for their image)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // Static shared memory comes first in the CTA's window and the dynamic part follows it with 16-byte alignment only
    // (the __align__ on an extern array is not honoured across the two): the tensor copies need their stage 128-byte
    // aligned, so the static part is one struct whose size is a multiple of 128 (checked below, trap if it ever moves).
    struct alignas(128) ScoreStatic {
        uint64_t full_bar[kStages], empty_bar[kStages];
        TileShared sh[kConsumers];
        TileCoord tc[kStages];  // the staged tile's coordinates, written by the producer before the copy is issued
    };
    static_assert(sizeof(ScoreStatic) % 128 == 0, "dynamic shared memory must start 128-byte aligned");
    __shared__ ScoreStatic st;
    uint64_t *const full_bar = st.full_bar, *const empty_bar = st.empty_bar;
    TileShared *const sh = st.sh;
    if (threadIdx.x == 0 && (smem_u32(stages) & 127u)) __trap();
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int stage_floats = p.ch * kPpTile;
    const int total = p.NT * p.B;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], (FUSED && !use_tmap) ? 32 * kProducers : 1);
            mbar_init(&empty_bar[s], kPpTile / 32);  // every warp of the consuming group arrives on its own
        }
        for (int g = 0; g < kConsumers; ++g) sh[g].done = 0u;
        mbar_fence_init();
    }
    __syncthreads();
    // the CTA's seq-th tile: adjacent tiles come in pairs, so that one CTA asks for 1 KB of every channel plane
    // back to back
    // (DRAM page locality); the tiles left over after the last complete round of pairs are dealt singly, so that
    // no CTA gets more than one tile above the average (2144 tiles on 148 SMs: 15 instead of 16)
    const int grid = (int)gridDim.x, full = ((total + 1) / 2) / grid;
    auto tile_of = [&](const int seq) {
        if (seq < 2 * full) return 2 * ((int)blockIdx.x + (seq >> 1) * grid) + (seq & 1);
        return 2 * full * grid + (seq - 2 * full) * grid + (int)blockIdx.x;
    };
    if (warp < kProducers) {
        // ---- producers
        if ((!FUSED || use_tmap) && warp > 0) return;  // one request per tile: one warp is plenty
        const uint64_t policy = l2_evict_first_policy();
        for (int seq = 0;; ++seq) {
            const int t = tile_of(seq);
            if (t >= total) break;
            const int s = seq % kStages, k = seq / kStages;
            if (k > 0) mbar_wait(&empty_bar[s], (k - 1) & 1);
            const TileCoord tc = tile_coord<FUSED>(p, t / p.NT, t % p.NT);
            float *dst = stages + (size_t)s * stage_floats;
            if (FUSED && use_tmap) {
                if (lane == 0) {
                    st.tc[s] = tc;  // ordered before the consumers' reads by the arrive (release) / their wait (acquire)
                    mbar_expect_tx(&full_bar[s], (uint32_t)(p.ch * kPpTile * 4));  // the whole box, zero fill included
                    tma_load_2d(dst, &tmaps.m[tc.l], tc.a0, tc.b * p.ch, &full_bar[s], policy);
                }
                continue;
            }
            if (FUSED) {
                if (lane * 4 < tc.cnt) {  // cnt is a multiple of 4 (bulk_ok)
                    const int c_per = (p.ch + kProducers - 1) / kProducers;
                    const int c_lo = warp * c_per, c_hi = min(p.ch, c_lo + c_per);
                    const size_t hw = (size_t)p.lv.hw[tc.l];
                    const float *g = tc.src + lane * 4 + c_lo * hw;
                    float *d = dst + lane * 4 + c_lo * kPpTile;
#pragma unroll 4
                    for (int c = c_lo; c < c_hi; ++c, g += hw, d += kPpTile) cp_async16(d, g);
                }
                cp_async_arrive(&full_bar[s]);
            } else if (lane == 0) {
                st.tc[s] = tc;
                mbar_expect_tx(&full_bar[s], (uint32_t)(p.ch * tc.cnt * 4));
                bulk_g2s(dst, tc.src, (uint32_t)(tc.cnt * p.ch * 4), &full_bar[s]);
            }
        }
    } else if (warp == kProducers) {
        // ---- publisher: the scored-tile counter of an image is released at GPU scope (nms_fast_kernel acquires it)
        // once the four warps of the tile's group have issued every record of the tile.  The release waits for those
        // stores to be performed; done here it costs the consumers nothing (it was 6 % of their time).
        // One fence per polling round covers every tile that completed since the last one.
        if (lane != 0) return;
        int pub[kConsumers];  // tiles of group g already published
        int open_groups = 0;
#pragma unroll
        for (int g = 0; g < kConsumers; ++g) { pub[g] = 0; open_groups += tile_of(g) < total ? 1 : 0; }
        while (open_groups > 0) {
            int fresh[kConsumers];
            int any = 0;
#pragma unroll
            for (int g = 0; g < kConsumers; ++g) {
                unsigned v;
                asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&sh[g].done)) : "memory");
                fresh[g] = (int)(v / (kPpTile / 32)) - pub[g];
                any += fresh[g];
            }
            if (!any) { __nanosleep(64); continue; }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
#pragma unroll
            for (int g = 0; g < kConsumers; ++g) {
                for (int i = 0; i < fresh[g]; ++i) {
                    const int t = tile_of(g + kConsumers * (pub[g] + i));
                    asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p.ws.ctr + (t / p.NT) * kImgCtr + kCtrDone) : "memory");
                }
                if (fresh[g]) {
                    pub[g] += fresh[g];
                    if (tile_of(g + kConsumers * pub[g]) >= total) --open_groups;
                }
            }
        }
    } else {
        // ---- consumers
        const int cw = warp - kProducers - 1, grp = cw >> 2, gtid = tid - 32 * (kProducers + 1) - grp * kPpTile;
        for (int seq = grp, it = 0;; seq += kConsumers, ++it) {
            const int t = tile_of(seq);
            if (t >= total) break;
            const int s = seq % kStages, k = seq / kStages;
            // Stage s may be consumed by a different group every time (kConsumers vs kStages), and an mbarrier wait only
            // knows the phase PARITY: waiting for fill k while fill k-1 has not even landed would return at
            // once.  The stage's release k-1 (which implies fill k-1 completed and was consumed; release k-2 is
            // already implied by this group's own progress) is therefore awaited first.
            long long *acc = p.prof ? p.prof + ((size_t)blockIdx.x * kConsumers + grp) * 8 : nullptr;
            const long long tw0 = acc ? clock64() : 0;
            if (kConsumers != kStages && k > 0) mbar_wait(&empty_bar[s], (k - 1) & 1);
            mbar_wait(&full_bar[s], k & 1);
            const long long tw1 = acc ? clock64() : 0;
            if (acc && gtid == 0) { acc[1] += tw1 - tw0; acc[7] += 1; }
            // the tile's coordinates: from the producer (it issued the copy from them), except in the cp.async mode,
            // whose completion mechanism does not order the producer's plain stores
            const TileCoord tc = (FUSED && !use_tmap) ? tile_coord<FUSED>(p, t / p.NT, t % p.NT) : st.tc[s];
            score_tile<FUSED>(p, stages + (size_t)s * stage_floats, tc, gtid, 1 + grp, sh[grp], it & 1, [&] {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }, acc, tw1);
            // every record of the tile has been issued by this warp: hand it to the publisher
            __syncwarp();
            if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], 1;" ::"r"(smem_u32(&sh[grp].done)) : "memory");
        }
    }
}

}  // namespace plyolo

#include "nms_fast.cuh"

namespace plyolo {

static thread_local long long *g_nms_prof = nullptr;
static thread_local long long *g_score_prof = nullptr;
static thread_local bool g_skip_nms = false;

static size_t cand_ws_layout(int B, int NT, CandWs *ws, unsigned char *base) {
    size_t off = 0;
    const size_t slots = (size_t)B * NT * kPpTile;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_ctr = take((size_t)B * kImgCtr * sizeof(int));  // first: zeroed before every call
    size_t o_cnt = take((size_t)B * NT * sizeof(int));
    size_t o_box = take(slots * sizeof(float4));
    size_t o_sc = take(slots * sizeof(float));
    size_t o_meta = take(slots * sizeof(int));
    size_t o_aux = take(slots * sizeof(float));
    size_t o_gkey = take((size_t)B * kBucketImg * sizeof(unsigned long long));
    size_t o_gbox = take((size_t)B * kBucketImg * sizeof(float4));
    size_t o_xkey = take((size_t)B * kMaxCross * sizeof(unsigned long long));
    size_t o_xbox = take((size_t)B * kMaxCross * sizeof(float4));
    if (ws) {
        ws->ctr = reinterpret_cast<int *>(base + o_ctr);
        ws->tile_count = reinterpret_cast<int *>(base + o_cnt);
        ws->box = reinterpret_cast<float4 *>(base + o_box);
        ws->score = reinterpret_cast<float *>(base + o_sc);
        ws->meta = reinterpret_cast<int *>(base + o_meta);
        ws->aux = reinterpret_cast<float *>(base + o_aux);
        ws->gkey = reinterpret_cast<unsigned long long *>(base + o_gkey);
        ws->gbox = reinterpret_cast<float4 *>(base + o_gbox);
        ws->xkey = reinterpret_cast<unsigned long long *>(base + o_xkey);
        ws->xbox = reinterpret_cast<float4 *>(base + o_xbox);
    }
    return off;
}

// worst-case tile count for A anchors split into at most PLYOLO_MAX_LEVELS levels
static int max_tiles(int A) { return (A + kPpTile - 1) / kPpTile + PLYOLO_MAX_LEVELS; }

// SMs the persistent score kernel uses: all of them, minus PLYOLO_SCORE_SMS_RESERVED (default 0).  In multi-GPU
// evaluation the detection all-gather of step i runs under the score kernel of step i+1 (side stream); the score CTA
// and the co-resident NMS CTA fill an SM completely, so a communication kernel only finds room on SMs left free.
static int sm_count() {
    static thread_local int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        const char *v = getenv("PLYOLO_SCORE_SMS_RESERVED");
        const int r = v ? atoi(v) : 0;
        if (r > 0 && r < n) n -= r;
    }
    return n;
}

// Per-device, once: the kernels' shared-memory limits (fixed maxima, so that no launch ever lowers another thread's
// limit) and the maximal shared-memory carveout — the score CTA (171.5 KB) and the NMS CTA (54.3 KB) only fit on
// one SM together when the SM is configured for 228 KB of shared memory.
constexpr size_t kScoreSmemLimit = (size_t)kStages * kPpTile * (5 + PLYOLO_MAX_CLASSES) * sizeof(float);  // 192 KB
static int ensure_kernel_attributes() {
    static std::mutex mu;
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (done[dev]) return PLYOLO_OK;
    cudaError_t e = cudaSuccess;
    auto set = [&](const void *fn, size_t smem) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    set((const void *)score_kernel<true>, kScoreSmemLimit);
    set((const void *)score_kernel<false>, kScoreSmemLimit);
    set((const void *)score_kernel_simple<true>, kScoreSmemLimit / kStages);
    set((const void *)score_kernel_simple<false>, kScoreSmemLimit / kStages);
    set((const void *)nms_fast_kernel, kFastSmemBytes);
    set((const void *)nms_general_kernel, kNmsSmemLimit);
    if (e != cudaSuccess) {
        set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PLYOLO_ERR_CUDA;
    }
    done[dev] = true;
    return PLYOLO_OK;
}

static bool no_device_launch() {
    static const bool v = [] { const char *e = getenv("PLYOLO_NO_CDP"); return e && e[0] == '1'; }();
    return v;
}

// `overlap`: the score stage was the persistent kernel (it triggers its dependents at once and publishes the
// per-image scored-tile counters), so the class-split NMS may start under it.
static int run_nms(int B, int A, int NT, double nms_thre, int class_agnostic, int max_nms, int max_det, int flavor,
                   const CandWs &ws, float *dets, int32_t *counts, int32_t *keep_idx, bool overlap, cudaStream_t stream,
                   int variant = PLYOLO_NMS_YOLOX, int n_peers = 0, float *const *peer_dets = nullptr,
                   int32_t *const *peer_counts = nullptr) {
    NmsParams np;
    np.n_peers = n_peers;
    for (int r = 0; r < PLYOLO_MAX_PEERS; ++r) {
        np.peer_dets[r] = r < n_peers ? peer_dets[r] : nullptr;
        np.peer_counts[r] = r < n_peers ? peer_counts[r] : nullptr;
    }
    np.B = B; np.NT = NT; np.max_nms = max_nms; np.max_det = max_det; np.flavor = flavor;
    np.variant = variant;
    np.fixed_span = variant == PLYOLO_NMS_YOLOV5 ? 4096.f : 0.f;             // yolov5_decoder.py:27, :70
    if (variant == PLYOLO_NMS_YOLOV3) class_agnostic = 1;                     // yolov3_decoder.py:103-106: offset never applied
    np.agnostic = class_agnostic ? 1 : 0;
    np.thr_f = (float)nms_thre; np.thr_d = nms_thre;
    int cap = 64;
    const int need = variant != PLYOLO_NMS_YOLOX ? (A < kMaxSortCap ? A : kMaxSortCap) : (max_nms < A ? max_nms : A);
    while (cap < need) cap <<= 1;
    np.sort_cap = cap;
    np.fast_cap = cap < kFastCap ? cap : kFastCap;
    np.ws = ws; np.dets = dets; np.counts = counts; np.keep_idx = keep_idx;
    np.prof = g_nms_prof;
    const size_t smem = nms_smem_bytes(cap, np.fast_cap, max_det, NT);
    PLYOLO_REQUIRE(smem <= kNmsSmemLimit, "nms working set (%zu B) exceeds shared memory", smem);
    np.only_image = -1;
    np.general_smem = (unsigned)smem;
    // the class-split kernel: class-aware NMS, max_det within its kept-key lists, slots within its key layout
    const bool split = variant == PLYOLO_NMS_YOLOX && !class_agnostic && max_det <= kFastMaxDet && A <= (1 << kFastAnchorBits) &&
                       !g_skip_nms;
    const bool pdl = pdl_enabled();
    record_stage_event(1, stream);
    if (g_skip_nms) return PLYOLO_OK;  // debug: time the score stage alone
    // device-side launches of the general path only outside stream capture (see nms_general_kernel); PLYOLO_NO_CDP=1
    // forces the host-launched variant everywhere
    bool host_general = no_device_launch();
    if (split && !host_general) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) { cudaGetLastError(); cs = cudaStreamCaptureStatusActive; }
        host_general = cs != cudaStreamCaptureStatusNone;
    }
    np.device_launch = host_general ? 0 : 1;
    if (split) {
        np.wait_tiles = overlap ? 1 : 0;
        np.all_general = 0;
        const cudaError_t e = launch_ex(nms_fast_kernel, dim3(kGroups, B), dim3(kFastThreads), kFastSmemBytes, stream,
                                        pdl && overlap, np);
        if (e != cudaSuccess) {
            set_error("nms_fast_kernel: %s", cudaGetErrorString(e));
            cudaGetLastError();
            return PLYOLO_ERR_CUDA;
        }
        count_launch();
    }
    if (split && !host_general) {  // the class-split kernel launches the general path itself, per image, when it has to
        record_stage_event(2, stream);
        return PLYOLO_OK;
    }
    np.wait_tiles = 0;
    np.all_general = split ? 0 : 1;
    np.prof = nullptr;
    // behind the class-split kernel: a plain (fully ordered) launch — its 180 KB CTAs must not sit on SMs while the
    // class-split clusters still need them
    const cudaError_t e = launch_ex(nms_general_kernel, dim3(B), dim3(kNmsThreads), smem, stream, false, np);
    if (e != cudaSuccess) {
        set_error("nms_general_kernel: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PLYOLO_ERR_CUDA;
    }
    count_launch();
    record_stage_event(2, stream);
    return PLYOLO_OK;
}

// zeroes the per-image counters, then scores every tile (persistent TMA pipeline; plain-load kernel if unaligned).
// *overlap = the persistent kernel ran (dependent launch trigger + scored-tile counters).
template <bool FUSED>
static int launch_score(const ScoreParams &sp, cudaStream_t stream, bool *overlap) {
    int rc = ensure_kernel_attributes();
    if (rc != PLYOLO_OK) return rc;
    if (cudaMemsetAsync(sp.ws.ctr, 0, (size_t)sp.B * kImgCtr * sizeof(int), stream) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    const size_t tile_b = (size_t)kPpTile * sp.ch * sizeof(float);
    record_stage_event(0, stream);
    *overlap = sp.bulk_ok != 0;
    if (sp.bulk_ok) {
        const size_t smem = tile_b * kStages;
        const int total = sp.NT * sp.B, sms = sm_count();
        TmapPack pack;
        memset(&pack, 0, sizeof(pack));
        int use_tmap = 0;
        if (FUSED) {
            // 2-D tensor maps over the head maps: [B * (5+C) planes][H*W anchors], box = (128 anchors, 5+C planes)
            typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static thread_local EncodeFn encode = nullptr;
            static thread_local bool looked_up = false;
            if (!looked_up) {
                looked_up = true;
                void *fn = nullptr;
                cudaDriverEntryPointQueryResult qres;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
                    qres == cudaDriverEntryPointSuccess)
                    encode = reinterpret_cast<EncodeFn>(fn);
                else
                    cudaGetLastError();
            }
            use_tmap = encode != nullptr && sp.ch <= 256;
            for (int l = 0; use_tmap && l < sp.lv.n; ++l) {
                const cuuint64_t gdim[2] = {(cuuint64_t)sp.lv.hw[l], (cuuint64_t)sp.B * sp.ch};
                const cuuint64_t gstride[1] = {(cuuint64_t)sp.lv.hw[l] * sizeof(float)};
                const cuuint32_t box[2] = {(cuuint32_t)kPpTile, (cuuint32_t)sp.ch};
                const cuuint32_t estr[2] = {1, 1};
                const CUresult r = encode(&pack.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(sp.lv.ptr[l]), gdim,
                                          gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) use_tmap = 0;  // fall back to the cp.async producer
            }
        }
        // Grid: the step takes as long as the busiest consumer group, k = ceil(total / (groups x SMs)) tiles; the fewest CTAs
        // that keep that k do the same work in the same time and leave the other SMs to the NMS clusters and to the next
        // batch's score CTAs (cfg2: 2144 tiles -> k = 4 -> 134 CTAs of 16 tiles instead of 148 of 14-15; 23.6 -> 23.2 us
        // per step with four batches in flight).  PLYOLO_SCORE_GRID=full restores one CTA per SM.
        static const bool full_grid = [] { const char *v = getenv("PLYOLO_SCORE_GRID"); return v && v[0] == 'f'; }();
        const int k_tiles = (total + kConsumers * sms - 1) / (kConsumers * sms);
        int grid = (total + kConsumers * k_tiles - 1) / (kConsumers * k_tiles);
        if (full_grid || grid > sms) grid = sms;
        if (grid > (total + 1) / 2) grid = (total + 1) / 2;  // tiles are dealt in pairs
        if (grid < 1) grid = 1;
        score_kernel<FUSED><<<grid, kScoreThreads, smem, stream>>>(sp, pack, use_tmap);
        PLYOLO_CHECK_LAUNCH("score_kernel");
    } else {
        score_kernel_simple<FUSED><<<dim3(sp.NT, sp.B), kPpTile, tile_b, stream>>>(sp);
        PLYOLO_CHECK_LAUNCH("score_kernel_simple");
    }
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
extern "C" void plyolo_debug_score_profile(void *device_buf) { plyolo::g_score_prof = static_cast<long long *>(device_buf); }
extern "C" void plyolo_debug_nms_profile(void *device_buf) { plyolo::g_nms_prof = static_cast<long long *>(device_buf); }
// debug hook: the calling thread's next postprocess calls stop after the score stage (bench.py times it alone)
extern "C" void plyolo_debug_skip_nms(int on) { plyolo::g_skip_nms = on != 0; }

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
    sp.prof = nullptr;
    sp.variant = PLYOLO_NMS_YOLOX;
    sp.lv.n = 0; sp.lv.A = A;
    cand_ws_layout(B, sp.NT, &sp.ws, static_cast<unsigned char *>(workspace));
    bool overlap = false;
    rc = launch_score<false>(sp, (cudaStream_t)stream, &overlap);
    if (rc != PLYOLO_OK) return rc;
    return run_nms(B, A, sp.NT, nms_thre, class_agnostic, max_nms, max_det, flavor, sp.ws, dets, counts, keep_idx,
                   overlap, (cudaStream_t)stream);
}

extern "C" int plyolo_decode_postprocess_bcast_f32(const float *const *host_lvl, const int *hs, const int *ws,
                                                   const int *strides, int n_levels, int B, int C, double conf_thre,
                                                   double nms_thre, int class_agnostic, int max_nms, int max_det,
                                                   int flavor, float *dets, int32_t *counts, int32_t *keep_idx, int n_peers,
                                                   float *const *host_peer_dets, int32_t *const *host_peer_counts,
                                                   void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(n_peers >= 0 && n_peers <= PLYOLO_MAX_PEERS, "n_peers=%d not in [0,%d]", n_peers, PLYOLO_MAX_PEERS);
    for (int r = 0; r < n_peers; ++r)
        PLYOLO_REQUIRE(host_peer_dets && host_peer_counts && host_peer_dets[r] && host_peer_counts[r], "peer %d: null pointer", r);
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
    sp.prof = g_score_prof;
    sp.variant = PLYOLO_NMS_YOLOX;
    cand_ws_layout(B, sp.NT, &sp.ws, static_cast<unsigned char *>(workspace));
    bool overlap = false;
    rc = launch_score<true>(sp, (cudaStream_t)stream, &overlap);
    if (rc != PLYOLO_OK) return rc;
    return run_nms(B, A, sp.NT, nms_thre, class_agnostic, max_nms, max_det, flavor, sp.ws, dets, counts, keep_idx,
                   overlap, (cudaStream_t)stream, PLYOLO_NMS_YOLOX, n_peers, host_peer_dets, host_peer_counts);
}

extern "C" int plyolo_decode_postprocess_f32(const float *const *host_lvl, const int *hs, const int *ws,
                                             const int *strides, int n_levels, int B, int C, double conf_thre,
                                             double nms_thre, int class_agnostic, int max_nms, int max_det,
                                             int flavor, float *dets, int32_t *counts, int32_t *keep_idx,
                                             void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    return plyolo_decode_postprocess_bcast_f32(host_lvl, hs, ws, strides, n_levels, B, C, conf_thre, nms_thre, class_agnostic,
                                               max_nms, max_det, flavor, dets, counts, keep_idx, 0, nullptr, nullptr, workspace,
                                               workspace_bytes, stream);
}

extern "C" int plyolo_postprocess_yolo_f32(const float *preds, int B, int N, int C, double conf_thre, double nms_thre,
                                           int variant, int class_agnostic, int max_nms, int max_det, float *dets,
                                           int32_t *counts, int32_t *keep_idx, void *workspace, size_t workspace_bytes,
                                           plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(preds != nullptr, "preds is null");
    PLYOLO_REQUIRE(variant == PLYOLO_NMS_YOLOV3 || variant == PLYOLO_NMS_YOLOV5, "variant=%d is not a YOLOv3 / YOLOv5 call site", variant);
    int rc = check_post_args(B, N, C, max_nms < kMaxSortCap ? max_nms : kMaxSortCap, max_det, 0, dets, counts, workspace, workspace_bytes);
    if (rc != PLYOLO_OK) return rc;
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    ScoreParams sp;
    sp.preds = preds; sp.B = B; sp.A = N; sp.C = C; sp.ch = 5 + C;
    sp.NT = (N + kPpTile - 1) / kPpTile;
    sp.conf_thr = (float)conf_thre;  // `tensor > python float` compares in fp32
    sp.bulk_ok = (((uintptr_t)preds & 15) == 0 && (N & 3) == 0) ? 1 : 0;
    sp.prof = nullptr;
    sp.variant = variant;
    sp.lv.n = 0; sp.lv.A = N;
    cand_ws_layout(B, sp.NT, &sp.ws, static_cast<unsigned char *>(workspace));
    bool overlap = false;
    rc = launch_score<false>(sp, (cudaStream_t)stream, &overlap);
    if (rc != PLYOLO_OK) return rc;
    return run_nms(B, N, sp.NT, nms_thre, class_agnostic, max_nms, max_det, 0, sp.ws, dets, counts, keep_idx, overlap,
                   (cudaStream_t)stream, variant);
}
