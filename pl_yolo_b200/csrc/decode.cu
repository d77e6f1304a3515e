// decode.cu — YOLOX head decode (yolox_loss.py:175-228, eval branch :25-36; yolox_decoder.py:16-58).
//
// One CTA per tile of kTile anchors of one (image, level).  The head maps are channel-planar
// ([B, 5+C, H, W]), the output is anchor-major ([B, A, 5+C]); the tile is transposed through shared
// memory so that both sides are fully coalesced:
//   load   : a warp reads 32 consecutive anchors of one channel plane (128 B per request), many
//            independent loads in flight per thread, sigmoid applied on the fly (inference)
//   smem   : tile[anchor][channel], row pitch 5+C (85: odd -> conflict-free column writes)
//   box    : one thread per anchor: (xy + grid) * s, exp(wh) * s, optional xyxy corners
//   store  : the tile is ONE contiguous chunk of preds (kTile*(5+C) floats) -> 128-bit stores
// HBM traffic = the algorithmic bytes: every input float is read once, every output written once.
#include "common.cuh"

namespace plyolo {

constexpr int kDecTile = 128;
constexpr int kDecThreads = 256;

struct DecodeParams {
    Levels lv;
    int B, C, ch;
    float *preds;
    float *ori;
    int inference;
    int vec_ok;  // preds base 16-byte aligned and A % 4 == 0
};

__global__ void __launch_bounds__(kDecThreads) decode_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) float tile[];  // [kDecTile][ch]
    const int b = blockIdx.y;
    const int tile_id = blockIdx.x;
    int l = 0;
#pragma unroll
    for (int i = 1; i < PLYOLO_MAX_LEVELS; ++i)
        if (i < p.lv.n && tile_id >= p.lv.tile0[i]) l = i;
    const int hw = p.lv.hw[l];
    const int a0 = (tile_id - p.lv.tile0[l]) * kDecTile;
    const int cnt = min(kDecTile, hw - a0);
    const int ch = p.ch;
    const float *__restrict__ src = p.lv.ptr[l] + (size_t)b * ch * hw + a0;
    const int t = threadIdx.x & (kDecTile - 1);
    const int half = threadIdx.x >> 7;

    // ---- load + transpose (+ sigmoid): thread (t, half) walks channels half, half+2, ...
    if (t < cnt) {
        constexpr int U = 8;
        int c = half;
        for (; c + 2 * (U - 1) < ch; c += 2 * U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(src + (size_t)(c + 2 * u) * hw + t);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int cc = c + 2 * u;
                float x = v[u];
                if (p.inference && cc >= 4) x = sigmoid_ref(x);  // yolox_loss.py:26-27
                tile[t * ch + cc] = x;
            }
        }
        for (; c < ch; c += 2) {
            float x = __ldg(src + (size_t)c * hw + t);
            if (p.inference && c >= 4) x = sigmoid_ref(x);
            tile[t * ch + c] = x;
        }
    }
    __syncthreads();

    // ---- box decode, one thread per anchor
    const size_t row0 = (size_t)b * p.lv.A + p.lv.off[l] + a0;
    if (threadIdx.x < cnt) {
        const int a = a0 + threadIdx.x;
        const int W = p.lv.w[l];
        const float s = p.lv.stride[l];
        float *r = tile + threadIdx.x * ch;
        const float px = r[0], py = r[1], pw = r[2], ph = r[3];
        if (p.ori) {
            // yolox_loss.py:214 — raw regression outputs
            reinterpret_cast<float4 *>(p.ori)[row0 + threadIdx.x] = make_float4(px, py, pw, ph);
        }
        const float gx = (float)(a % W), gy = (float)(a / W);  // :198-200
        const float cx = (px + gx) * s;                        // :217
        const float cy = (py + gy) * s;
        const float w = expf(pw) * s;                          // :219
        const float h = expf(ph) * s;
        if (p.inference) {
            r[0] = cx - w / 2;  // :31-34
            r[1] = cy - h / 2;
            r[2] = cx + w / 2;
            r[3] = cy + h / 2;
        } else {
            r[0] = cx; r[1] = cy; r[2] = w; r[3] = h;
        }
    }
    __syncthreads();

    // ---- store: contiguous chunk of cnt*ch floats
    float *dst = p.preds + row0 * ch;
    const int n = cnt * ch;
    if (p.vec_ok && (n & 3) == 0) {
        const float4 *s4 = reinterpret_cast<const float4 *>(tile);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int i = threadIdx.x; i < (n >> 2); i += kDecThreads) d4[i] = s4[i];
    } else {
        for (int i = threadIdx.x; i < n; i += kDecThreads) dst[i] = tile[i];
    }
}

}  // namespace plyolo

extern "C" int plyolo_decode_f32(const float *const *host_lvl, const int *hs, const int *ws, const int *strides,
                                 int n_levels, int B, int C, float *preds, float *ori_boxes, int inference,
                                 plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(B >= 1 && B <= 65535, "B=%d not in [1,65535]", B);
    PLYOLO_REQUIRE(C >= 1 && C <= PLYOLO_MAX_CLASSES, "C=%d not in [1,%d]", C, PLYOLO_MAX_CLASSES);
    PLYOLO_REQUIRE(preds != nullptr, "preds is null");
    DecodeParams p;
    int rc = make_levels(p.lv, host_lvl, hs, ws, strides, n_levels, kDecTile);
    if (rc != PLYOLO_OK) return rc;
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    p.B = B; p.C = C; p.ch = 5 + C;
    p.preds = preds; p.ori = ori_boxes; p.inference = inference ? 1 : 0;
    PLYOLO_REQUIRE(ori_boxes == nullptr || ((uintptr_t)ori_boxes & 15) == 0, "ori_boxes must be 16-byte aligned");
    // every tile start is a multiple of 4 anchors when all level sizes are, so rows*ch*4 stays 16-byte aligned
    bool vec = ((uintptr_t)preds & 15) == 0 && (p.lv.A & 3) == 0;
    for (int l = 0; l < p.lv.n; ++l) vec = vec && (p.lv.hw[l] & 3) == 0;
    p.vec_ok = vec ? 1 : 0;
    const size_t smem = (size_t)kDecTile * p.ch * sizeof(float);
    if (first_use_on_device(0)) cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    dim3 grid(p.lv.tile0[p.lv.n], B);
    decode_kernel<<<grid, kDecThreads, smem, (cudaStream_t)stream>>>(p);
    PLYOLO_CHECK_LAUNCH("decode_kernel");
    return PLYOLO_OK;
}

// ---- format_outputs (models/evaluators/postprocess.py:95-138), device part ---------------------------
// per detection: bboxes /= scale (in place: the VOC rows keep the rescaled corners), xyxy2xywh
// (models/utils/bbox.py:58-63).  `tensor /= python_float` on CUDA multiplies by the reciprocal of the scalar
// (ATen BinaryDivTrueKernel.cu, CPU-scalar fast path); measured on B200 with torch 2.11: the reciprocal is
// taken in double and then rounded, x * (float)(1.0 / scale) — the caller passes that factor.
namespace plyolo {
__global__ void format_dets_kernel(const float *dets, const int32_t *counts, const float *inv_scales, int B, int max_det,
                                   float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * max_det) return;
    const int b = i / max_det, r = i - b * max_det;
    float *o = out + (size_t)i * 8;
    if (r >= counts[b]) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = 0.f;
        return;
    }
    const float *d = dets + (size_t)i * 6;
    const float inv = inv_scales[b];
    const float x1 = d[0] * inv, y1 = d[1] * inv, x2 = d[2] * inv, y2 = d[3] * inv;
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
    o[4] = x2 - x1; o[5] = y2 - y1;  // xyxy2xywh
    o[6] = d[4]; o[7] = d[5];
}
}  // namespace plyolo

extern "C" int plyolo_format_dets_f32(const float *dets, const int32_t *counts, const float *inv_scales, int B, int max_det,
                                      float *out, plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(dets && counts && inv_scales && out, "null pointer");
    PLYOLO_REQUIRE(B >= 1 && max_det >= 1, "B=%d max_det=%d must be positive", B, max_det);
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    const int n = B * max_det;
    format_dets_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dets, counts, inv_scales, B, max_det, out);
    PLYOLO_CHECK_LAUNCH("format_dets_kernel");
    return PLYOLO_OK;
}
