// loss.cu — loss tail of YOLOXLoss (yolox_loss.py:121-163) for the whole batch: target building
// (one_hot * IoU, matched GT boxes, fg mask; :123-127, :142-147), GIoU loss (IOUloss, loss_type "giou",
// models/layers/losses/iou_loss.py:7-50), BCE-with-logits on objectness (all B*A anchors) and on the classes
// of the foreground anchors (:150-154) — forward as three sums, backward straight into the head maps.
//
//  yolox_loss_fwd_kernel   one thread per anchor: objectness term of every anchor; the (few) foreground anchors
//     additionally get the GIoU term (owning thread) and the class terms (warp-cooperative: lanes over classes,
//     coalesced row read).  Per-CTA partial sums, then yolox_loss_reduce_kernel adds them in a fixed order
//     (deterministic results, no float atomics).
//  yolox_loss_bwd_kernel   one CTA per tile of 128 anchors of one (image, level): d(sum)/d(preds) chained through
//     the decode (cx = (px + gx) s, w = exp(pw) s: yolox_loss.py:217-219) and written channel-planar into the
//     head-map gradients with coalesced 512-byte rows — neither d(loss)/d(preds) [B,A,5+C] nor the reference's
//     gathered / concatenated target tensors ever exist in HBM.
//
// Floating point: plain fp32, the summation order differs from ATen's (parity within 1e-5 relative, stated in
// the tests); this is the only part of the library that is not bit-exact by contract.
#include "common.cuh"

namespace plyolo {

constexpr int kLossThreads = 256;
constexpr int kLossTile = 128;

struct LossParams {
    const float *preds;    // [B,A,ch] training-mode decode output
    const float *labels;   // [B,Lmax,5]
    const uint8_t *fg;     // [B,A]
    const int32_t *mg;     // [B,A]
    const float *miou;     // [B,A]
    int B, A, C, ch, Lmax;
    float *partial;        // [gridDim.x * gridDim.y][3]
    float *sums;           // [3]
};

// BCEWithLogitsLoss(reduction="none"): (1 - t) x - log_sigmoid(x), log_sigmoid(x) = min(x, 0) - log1p(exp(-|x|))
__device__ __forceinline__ float bce_logits(const float x, const float t) {
    return (1.0f - t) * x - (fminf(x, 0.f) - log1pf(expf(-fabsf(x))));
}

struct GiouOut {
    float loss;
    float d[4];  // d loss / d (cx, cy, w, h) of the prediction
};

// IOUloss(loss_type="giou") for one (prediction, target) pair in (cx, cy, w, h) and, if WITH_GRAD, its gradient
// with respect to the prediction (torch autograd semantics: max/min send the gradient to the selected operand,
// half to each on ties; clamp passes it inside [min, max]; the `en` mask carries none).
template <bool WITH_GRAD>
__device__ __forceinline__ GiouOut giou_pair(const float cx, const float cy, const float w, const float h, const float gx,
                                             const float gy, const float gw, const float gh) {
    GiouOut o;
    const float px1 = cx - w / 2, px2 = cx + w / 2, py1 = cy - h / 2, py2 = cy + h / 2;
    const float tx1 = gx - gw / 2, tx2 = gx + gw / 2, ty1 = gy - gh / 2, ty2 = gy + gh / 2;
    const float tlx = fmaxf(px1, tx1), tly = fmaxf(py1, ty1), brx = fminf(px2, tx2), bry = fminf(py2, ty2);
    const float area_p = w * h, area_g = gw * gh;
    const float en = (tlx < brx ? 1.f : 0.f) * (tly < bry ? 1.f : 0.f);
    const float wi = brx - tlx, hi = bry - tly;
    const float area_i = (wi * hi) * en;
    const float U = area_p + area_g - area_i + 1e-16f;
    const float iou = area_i / U;
    const float cx1 = fminf(px1, tx1), cy1 = fminf(py1, ty1), cx2 = fmaxf(px2, tx2), cy2 = fmaxf(py2, ty2);
    const float wc = cx2 - cx1, hc = cy2 - cy1;
    const float area_c = wc * hc;
    const float Cc = fmaxf(area_c, 1e-16f);
    const float giou = iou - (area_c - area_i) / Cc;
    o.loss = 1.0f - fminf(fmaxf(giou, -1.0f), 1.0f);
    if (WITH_GRAD) {
        const float g_giou = (giou >= -1.0f && giou <= 1.0f) ? -1.0f : 0.f;  // d loss / d giou
        // giou = iou - (area_c - area_i) / Cc
        const float g_iou = g_giou;
        float g_area_i = g_giou / Cc;
        float g_area_c = -g_giou / Cc;
        if (area_c >= 1e-16f) g_area_c += g_giou * (area_c - area_i) / (Cc * Cc);  // through Cc = clamp(area_c)
        // iou = area_i / U, U = area_p + area_g - area_i + eps
        g_area_i += g_iou * (1.0f / U + area_i / (U * U));
        const float g_area_p = -g_iou * area_i / (U * U);
        // area_i = wi * hi * en
        const float g_wi = g_area_i * hi * en, g_hi = g_area_i * wi * en;
        // area_c = wc * hc
        const float g_wc = g_area_c * hc, g_hc = g_area_c * wc;
        auto sel_gt = [](const float a, const float b) { return a > b ? 1.0f : (a == b ? 0.5f : 0.f); };  // d max(a,b)/da
        auto sel_lt = [](const float a, const float b) { return a < b ? 1.0f : (a == b ? 0.5f : 0.f); };  // d min(a,b)/da
        // wi = brx - tlx, tlx = max(px1, tx1), brx = min(px2, tx2); wc = cx2 - cx1, cx1 = min(px1, tx1), cx2 = max(px2, tx2)
        const float g_px1 = -g_wi * sel_gt(px1, tx1) - g_wc * sel_lt(px1, tx1);
        const float g_px2 = g_wi * sel_lt(px2, tx2) + g_wc * sel_gt(px2, tx2);
        const float g_py1 = -g_hi * sel_gt(py1, ty1) - g_hc * sel_lt(py1, ty1);
        const float g_py2 = g_hi * sel_lt(py2, ty2) + g_hc * sel_gt(py2, ty2);
        o.d[0] = g_px1 + g_px2;
        o.d[1] = g_py1 + g_py2;
        o.d[2] = 0.5f * (g_px2 - g_px1) + g_area_p * h;
        o.d[3] = 0.5f * (g_py2 - g_py1) + g_area_p * w;
    }
    return o;
}

__global__ void __launch_bounds__(kLossThreads) yolox_loss_fwd_kernel(const LossParams p) {
    __shared__ float red[kLossThreads / 32][3];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a = blockIdx.x * kLossThreads + tid;
    const bool in = a < p.A;
    const size_t row = (size_t)b * p.A + (in ? a : 0);
    const float *r = p.preds + row * p.ch;
    const bool fg = in && p.fg[row] != 0;
    float s_iou = 0.f, s_obj = 0.f, s_cls = 0.f;
    int gc = -1;
    float tiou = 0.f;
    if (in) s_obj = bce_logits(__ldg(r + 4), fg ? 1.0f : 0.f);  // :152, target = fg mask (:126)
    if (fg) {
        const float *L = p.labels + ((size_t)b * p.Lmax + p.mg[row]) * 5;  // reg_target / class of the matched GT (:127)
        gc = (int)L[0];
        tiou = p.miou[row];
        s_iou = giou_pair<false>(__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3), L[1], L[2], L[3], L[4]).loss;  // :150
    }
    // class terms of the warp's foreground anchors, lanes over classes (:154; target one_hot(class) * IoU, :123-125)
    unsigned m = __ballot_sync(0xffffffffu, fg);
    while (m) {  // four foreground anchors per trip: their rows are in flight together
        constexpr int U = 4, NK = (PLYOLO_MAX_CLASSES + 31) / 32;
        int jj[U];
        float x[U][NK];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            jj[u] = m ? __ffs(m) - 1 : -1;
            m &= m - 1;  // 0 stays 0
            const float *rj = p.preds + ((size_t)b * p.A + (a - lane + max(jj[u], 0))) * p.ch + 5;
#pragma unroll
            for (int k = 0; k < NK; ++k) x[u][k] = (jj[u] >= 0 && lane + 32 * k < p.C) ? __ldg(rj + lane + 32 * k) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (jj[u] < 0) break;  // warp-uniform
            const int gcj = __shfl_sync(0xffffffffu, gc, jj[u]);
            const float tj = __shfl_sync(0xffffffffu, tiou, jj[u]);
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < NK; ++k) {
                const int c = lane + 32 * k;
                if (c < p.C) v += bce_logits(x[u][k], c == gcj ? tj : 0.f);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_cls += v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_iou += __shfl_xor_sync(0xffffffffu, s_iou, o);
        s_obj += __shfl_xor_sync(0xffffffffu, s_obj, o);
    }
    if (lane == 0) { red[warp][0] = s_iou; red[warp][1] = s_obj; red[warp][2] = s_cls; }
    __syncthreads();
    if (tid < 3) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][tid];
        p.partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 3 + tid] = v;
    }
}

// sums[k] = sum of partial[i][k], fixed order: 256 strided accumulators, then a tree
__global__ void __launch_bounds__(256) yolox_loss_reduce_kernel(const float *partial, const int n, float *sums) {
    __shared__ float sh[256];
    for (int k = 0; k < 3; ++k) {
        float v = 0.f;
        for (int i = threadIdx.x; i < n; i += 256) v += partial[(size_t)i * 3 + k];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) sums[k] = sh[0];
        __syncthreads();
    }
}

struct LossBwdParams {
    LossParams f;
    Levels lv;            // ptr[l] = gradient of head map l (written)
    const float *gscale;  // [3] device: upstream gradient of (sum giou, sum obj, sum cls)
    int vec_ok;           // 128-bit stores are aligned
};

__global__ void __launch_bounds__(kLossThreads) yolox_loss_bwd_kernel(const LossBwdParams p) {
    extern __shared__ __align__(16) float gt[];  // [ch][kLossTile] gradient tile, channel-major like the head maps
    const int b = blockIdx.y, tile_id = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int l = 0;
#pragma unroll
    for (int i = 1; i < PLYOLO_MAX_LEVELS; ++i)
        if (i < p.lv.n && tile_id >= p.lv.tile0[i]) l = i;
    const int hw = p.lv.hw[l];
    const int a0 = (tile_id - p.lv.tile0[l]) * kLossTile;
    const int cnt = min(kLossTile, hw - a0);
    const int ch = p.f.ch, C = p.f.C;
    const float g_iou = p.gscale[0], g_obj = p.gscale[1], g_cls = p.gscale[2];
    for (int i = tid; i < ch * kLossTile / 4; i += kLossThreads) reinterpret_cast<float4 *>(gt)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    // threads [0, 128): one anchor each — objectness of every anchor, box terms of the foreground ones;
    // then every warp handles the class rows of its own foreground anchors, lanes over classes
    const int t = tid;
    const bool in = t < cnt;
    const size_t row = (size_t)b * p.f.A + p.lv.off[l] + a0 + (in ? t : 0);
    const float *r = p.f.preds + row * ch;
    const bool fg = in && t < kLossTile && p.f.fg[row] != 0;
    int gc = -1;
    float tiou = 0.f;
    if (in && t < kLossTile) {
        const float x = __ldg(r + 4);
        gt[4 * kLossTile + t] = g_obj * (sigmoid_ref(x) - (fg ? 1.0f : 0.f));  // d bce / dx = sigmoid(x) - t
    }
    if (fg) {
        const float *L = p.f.labels + ((size_t)b * p.f.Lmax + p.f.mg[row]) * 5;
        gc = (int)L[0];
        tiou = p.f.miou[row];
        const float w = __ldg(r + 2), h = __ldg(r + 3);
        const GiouOut o = giou_pair<true>(__ldg(r), __ldg(r + 1), w, h, L[1], L[2], L[3], L[4]);
        const float s = p.lv.stride[l];
        gt[0 * kLossTile + t] = g_iou * o.d[0] * s;  // cx = (px + grid) * s
        gt[1 * kLossTile + t] = g_iou * o.d[1] * s;
        gt[2 * kLossTile + t] = g_iou * o.d[2] * w;  // w = exp(pw) * s
        gt[3 * kLossTile + t] = g_iou * o.d[3] * h;
    }
    if (tid < kLossTile) {  // warps 0..3 (uniform per warp)
        unsigned m = __ballot_sync(0xffffffffu, fg);
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const int tj = t - lane + j;
            const float *rj = p.f.preds + ((size_t)b * p.f.A + p.lv.off[l] + a0 + tj) * ch + 5;
            const int gcj = __shfl_sync(0xffffffffu, gc, j);
            const float ij = __shfl_sync(0xffffffffu, tiou, j);
            for (int c = lane; c < C; c += 32)
                gt[(5 + c) * kLossTile + tj] = g_cls * (sigmoid_ref(__ldg(rj + c)) - (c == gcj ? ij : 0.f));
        }
    }
    __syncthreads();
    // channel-planar store: row c of the tile = cnt consecutive cells of plane c of the head-map gradient
    float *dst = const_cast<float *>(p.lv.ptr[l]) + (size_t)b * ch * hw + a0;
    if (p.vec_ok) {  // every level size a multiple of 4, 16-byte aligned maps: 128-bit stores
        for (int i = tid; i < ch * (kLossTile / 4); i += kLossThreads) {
            const int c = i / (kLossTile / 4), x4 = i - c * (kLossTile / 4);
            if (x4 * 4 < cnt) reinterpret_cast<float4 *>(dst + (size_t)c * hw)[x4] = reinterpret_cast<const float4 *>(gt + c * kLossTile)[x4];
        }
    } else {
        for (int i = tid; i < ch * kLossTile; i += kLossThreads) {
            const int c = i / kLossTile, x = i - c * kLossTile;
            if (x < cnt) dst[(size_t)c * hw + x] = gt[i];
        }
    }
}


// ---- use_l1 term (yolox_loss.py:128-133, :158; get_l1_type :373-378) ------------------------------------------------
// L1 distance between the RAW regression outputs of the foreground anchors and the matched GT expressed in the head's
// own units: (gx / s - grid_x, gy / s - grid_y, log(gw / s + 1e-8), log(gh / s + 1e-8)).
struct L1Params {
    const float *ori;      // [B,A,4] raw regression outputs (second output of the training-mode decode)
    const float *labels;   // [B,Lmax,5]
    const uint8_t *fg;     // [B,A]
    const int32_t *mg;     // [B,A]
    Levels lv;             // geometry; bwd: ptr[l] = gradient of head map l (accumulated into)
    int B, A, Lmax, ch;
    float *partial;        // fwd: [B * ceil(A / kLossThreads)]
    const float *gscale;   // bwd: [1] device, upstream gradient of the sum
};

// target and level / cell of anchor a (a foreground anchor of image b)
__device__ __forceinline__ void l1_target(const L1Params &p, const int b, const int a, const size_t row, float (&t)[4], int &l,
                                          int &cell) {
    l = 0;
#pragma unroll
    for (int i = 1; i < PLYOLO_MAX_LEVELS; ++i)
        if (i < p.lv.n && a >= p.lv.off[i]) l = i;
    cell = a - p.lv.off[l];
    const int W = p.lv.w[l], gy = cell / W, gx = cell - gy * W;
    const float s = p.lv.stride[l];
    const float *L = p.labels + ((size_t)b * p.Lmax + p.mg[row]) * 5;
    t[0] = L[1] / s - (float)gx;          // :374
    t[1] = L[2] / s - (float)gy;          // :375
    t[2] = logf(L[3] / s + 1e-8f);        // :376
    t[3] = logf(L[4] / s + 1e-8f);        // :377
}

__global__ void __launch_bounds__(kLossThreads) yolox_l1_fwd_kernel(const L1Params p) {
    __shared__ float red[kLossThreads / 32];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a = blockIdx.x * kLossThreads + tid;
    const size_t row = (size_t)b * p.A + (a < p.A ? a : 0);
    float v = 0.f;
    if (a < p.A && p.fg[row] != 0) {
        float t[4];
        int l, cell;
        l1_target(p, b, a, row, t, l, cell);
        const float4 o = *reinterpret_cast<const float4 *>(p.ori + row * 4);
        v = ((fabsf(o.x - t[0]) + fabsf(o.y - t[1])) + fabsf(o.z - t[2])) + fabsf(o.w - t[3]);  // nn.L1Loss(reduction="none")
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) sum += red[w];
        p.partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = sum;
    }
}

// sum[0] = sum of partial[i], fixed order: 256 strided accumulators, then a tree
__global__ void __launch_bounds__(256) yolox_l1_reduce_kernel(const float *partial, const int n, float *sum) {
    __shared__ float sh[256];
    float v = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) v += partial[i];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sum[0] = sh[0];
}

// d(sum * g) / d(raw outputs) = g * sign(ori - target), added to the regression planes of the head-map gradients
__global__ void __launch_bounds__(kLossThreads) yolox_l1_bwd_kernel(const L1Params p) {
    const int b = blockIdx.y, a = blockIdx.x * kLossThreads + threadIdx.x;
    if (a >= p.A) return;
    const size_t row = (size_t)b * p.A + a;
    if (p.fg[row] == 0) return;
    float t[4];
    int l, cell;
    l1_target(p, b, a, row, t, l, cell);
    const float4 o = *reinterpret_cast<const float4 *>(p.ori + row * 4);
    const float g = p.gscale[0];
    const float d[4] = {o.x - t[0], o.y - t[1], o.z - t[2], o.w - t[3]};
    const int hw = p.lv.hw[l];
    float *dst = const_cast<float *>(p.lv.ptr[l]) + (size_t)b * p.ch * hw + cell;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float sg = d[c] > 0.f ? 1.0f : (d[c] < 0.f ? -1.0f : 0.f);  // sign(); NaN -> 0 like torch.sign's comparison form
        dst[(size_t)c * hw] += g * sg;
    }
}

static size_t loss_partials(int B, int A) { return (size_t)B * ((A + kLossThreads - 1) / kLossThreads); }

}  // namespace plyolo

extern "C" size_t plyolo_yolox_loss_workspace_bytes(int B, int A) {
    if (B < 1 || A < 1) return 0;
    return plyolo::align_up(plyolo::loss_partials(B, A) * 3 * sizeof(float), 256);
}

static int check_loss_args(const float *preds, const float *labels, const uint8_t *fg_mask, const int32_t *matched_gt,
                           const float *matched_iou, int B, int A, int C, int Lmax) {
    using namespace plyolo;
    PLYOLO_REQUIRE(preds && labels && fg_mask && matched_gt && matched_iou, "an input pointer is null");
    PLYOLO_REQUIRE(B >= 1 && B <= 65535, "B=%d not in [1,65535]", B);
    PLYOLO_REQUIRE(A >= 1 && A < (1 << 24), "A=%d not in [1,2^24)", A);
    PLYOLO_REQUIRE(C >= 1 && C <= PLYOLO_MAX_CLASSES, "C=%d not in [1,%d]", C, PLYOLO_MAX_CLASSES);
    PLYOLO_REQUIRE(Lmax >= 1 && Lmax <= 32767, "Lmax=%d not in [1,32767]", Lmax);
    return PLYOLO_OK;
}

extern "C" int plyolo_yolox_loss_f32(const float *preds, const float *labels, const uint8_t *fg_mask,
                                     const int32_t *matched_gt, const float *matched_iou, int B, int A, int C, int Lmax,
                                     float *sums, void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    using namespace plyolo;
    int rc = check_loss_args(preds, labels, fg_mask, matched_gt, matched_iou, B, A, C, Lmax);
    if (rc != PLYOLO_OK) return rc;
    PLYOLO_REQUIRE(sums != nullptr, "sums is null");
    if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < plyolo_yolox_loss_workspace_bytes(B, A)) {
        set_error("workspace null, not 256-byte aligned, or smaller than plyolo_yolox_loss_workspace_bytes()");
        return PLYOLO_ERR_WORKSPACE;
    }
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    LossParams p;
    p.preds = preds; p.labels = labels; p.fg = fg_mask; p.mg = matched_gt; p.miou = matched_iou;
    p.B = B; p.A = A; p.C = C; p.ch = 5 + C; p.Lmax = Lmax;
    p.partial = static_cast<float *>(workspace); p.sums = sums;
    const dim3 grid((A + kLossThreads - 1) / kLossThreads, B);
    yolox_loss_fwd_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(p);
    PLYOLO_CHECK_LAUNCH("yolox_loss_fwd_kernel");
    yolox_loss_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p.partial, (int)loss_partials(B, A), sums);
    PLYOLO_CHECK_LAUNCH("yolox_loss_reduce_kernel");
    return PLYOLO_OK;
}

extern "C" int plyolo_yolox_loss_backward_f32(const float *preds, const float *labels, const uint8_t *fg_mask,
                                              const int32_t *matched_gt, const float *matched_iou, int B, int C, int Lmax,
                                              const float *grad_sums, float *const *host_grad_lvl, const int *hs,
                                              const int *ws, const int *strides, int n_levels, plyolo_stream_t stream) {
    using namespace plyolo;
    LossBwdParams p;
    int rc = make_levels(p.lv, const_cast<const float *const *>(host_grad_lvl), hs, ws, strides, n_levels, kLossTile);
    if (rc != PLYOLO_OK) return rc;
    rc = check_loss_args(preds, labels, fg_mask, matched_gt, matched_iou, B, p.lv.A, C, Lmax);
    if (rc != PLYOLO_OK) return rc;
    PLYOLO_REQUIRE(grad_sums != nullptr, "grad_sums is null");
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    p.f.preds = preds; p.f.labels = labels; p.f.fg = fg_mask; p.f.mg = matched_gt; p.f.miou = matched_iou;
    p.f.B = B; p.f.A = p.lv.A; p.f.C = C; p.f.ch = 5 + C; p.f.Lmax = Lmax;
    p.f.partial = nullptr; p.f.sums = nullptr;
    p.gscale = grad_sums;
    bool vec = true;
    for (int l = 0; l < p.lv.n; ++l) vec = vec && ((uintptr_t)p.lv.ptr[l] & 15) == 0 && (p.lv.hw[l] & 3) == 0;
    p.vec_ok = vec ? 1 : 0;
    const size_t smem = (size_t)kLossTile * p.f.ch * sizeof(float);
    if (first_use_on_device(1)) cudaFuncSetAttribute(yolox_loss_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    yolox_loss_bwd_kernel<<<dim3(p.lv.tile0[p.lv.n], B), kLossThreads, smem, (cudaStream_t)stream>>>(p);
    PLYOLO_CHECK_LAUNCH("yolox_loss_bwd_kernel");
    return PLYOLO_OK;
}

static int check_l1_args(const float *ori, const float *labels, const uint8_t *fg_mask, const int32_t *matched_gt, int B, int Lmax) {
    using namespace plyolo;
    PLYOLO_REQUIRE(ori && labels && fg_mask && matched_gt, "an input pointer is null");
    PLYOLO_REQUIRE(((uintptr_t)ori & 15) == 0, "ori must be 16-byte aligned");
    PLYOLO_REQUIRE(B >= 1 && B <= 65535, "B=%d not in [1,65535]", B);
    PLYOLO_REQUIRE(Lmax >= 1 && Lmax <= 32767, "Lmax=%d not in [1,32767]", Lmax);
    return PLYOLO_OK;
}

extern "C" int plyolo_yolox_l1_f32(const float *ori, const float *labels, const uint8_t *fg_mask, const int32_t *matched_gt,
                                   int B, int Lmax, const int *hs, const int *ws, const int *strides, int n_levels,
                                   float *sum, void *workspace, size_t workspace_bytes, plyolo_stream_t stream) {
    using namespace plyolo;
    L1Params p;
    const float *none[PLYOLO_MAX_LEVELS] = {nullptr};
    PLYOLO_REQUIRE(n_levels >= 1 && n_levels <= PLYOLO_MAX_LEVELS, "n_levels=%d not in [1,%d]", n_levels, PLYOLO_MAX_LEVELS);
    for (int l = 0; l < n_levels; ++l) none[l] = ori;  // make_levels wants non-null level pointers: unused by the forward
    int rc = make_levels(p.lv, none, hs, ws, strides, n_levels, kLossTile);
    if (rc != PLYOLO_OK) return rc;
    rc = check_l1_args(ori, labels, fg_mask, matched_gt, B, Lmax);
    if (rc != PLYOLO_OK) return rc;
    PLYOLO_REQUIRE(sum != nullptr, "sum is null");
    const int A = p.lv.A;
    if (!workspace || ((uintptr_t)workspace & 255) || workspace_bytes < plyolo_yolox_loss_workspace_bytes(B, A)) {
        set_error("workspace null, not 256-byte aligned, or smaller than plyolo_yolox_loss_workspace_bytes()");
        return PLYOLO_ERR_WORKSPACE;
    }
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    p.ori = ori; p.labels = labels; p.fg = fg_mask; p.mg = matched_gt;
    p.B = B; p.A = A; p.Lmax = Lmax; p.ch = 0;
    p.partial = static_cast<float *>(workspace); p.gscale = nullptr;
    const dim3 grid((A + kLossThreads - 1) / kLossThreads, B);
    yolox_l1_fwd_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(p);
    PLYOLO_CHECK_LAUNCH("yolox_l1_fwd_kernel");
    yolox_l1_reduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p.partial, (int)loss_partials(B, A), sum);
    PLYOLO_CHECK_LAUNCH("yolox_l1_reduce_kernel");
    return PLYOLO_OK;
}

extern "C" int plyolo_yolox_l1_backward_f32(const float *ori, const float *labels, const uint8_t *fg_mask,
                                            const int32_t *matched_gt, int B, int C, int Lmax, const float *grad_sum,
                                            float *const *host_grad_lvl, const int *hs, const int *ws, const int *strides,
                                            int n_levels, plyolo_stream_t stream) {
    using namespace plyolo;
    L1Params p;
    int rc = make_levels(p.lv, const_cast<const float *const *>(host_grad_lvl), hs, ws, strides, n_levels, kLossTile);
    if (rc != PLYOLO_OK) return rc;
    rc = check_l1_args(ori, labels, fg_mask, matched_gt, B, Lmax);
    if (rc != PLYOLO_OK) return rc;
    PLYOLO_REQUIRE(C >= 1 && C <= PLYOLO_MAX_CLASSES, "C=%d not in [1,%d]", C, PLYOLO_MAX_CLASSES);
    PLYOLO_REQUIRE(grad_sum != nullptr, "grad_sum is null");
    rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    p.ori = ori; p.labels = labels; p.fg = fg_mask; p.mg = matched_gt;
    p.B = B; p.A = p.lv.A; p.Lmax = Lmax; p.ch = 5 + C;
    p.partial = nullptr; p.gscale = grad_sum;
    const dim3 grid((p.A + kLossThreads - 1) / kLossThreads, B);
    yolox_l1_bwd_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(p);
    PLYOLO_CHECK_LAUNCH("yolox_l1_bwd_kernel");
    return PLYOLO_OK;
}
