// simota_ops.cu — the two stand-alone pieces of the SimOTA assignment that the reference exposes as module-level
// functions: get_in_boxes_info (models/losses/yolox/yolox_loss.py:231-315) and dynamic_k_matching
// (:318-370).  plyolo_simota_f32 (simota.cu) fuses both with the cost computation for the training path; these entry
// points serve callers that use the reference functions directly, on tensors of their own (any anchor layout, any
// cost matrix).  Same arithmetic contract: one fp32 rounding per reference op, ATen's CUDA reduce order for the
// top-10 IoU sum, ties -> lowest index (stable-sort semantics).
#include "common.cuh"

namespace plyolo {

// ---- get_in_boxes_info -------------------------------------------------------------------------------------------
// one thread per anchor, the GT boxes staged in shared memory in chunks
constexpr int kIbThreads = 256;
constexpr int kIbChunk = 512;

__global__ void __launch_bounds__(kIbThreads) in_boxes_kernel(const float *gt /*[G,4] cx,cy,w,h*/, const float *es,
                                                             const float *xs, const float *ys, const int A, const int G,
                                                             uint8_t *fg /*[A]*/, uint8_t *in_box /*[G,A]*/, uint8_t *in_ctr /*[G,A]*/) {
    __shared__ float4 edge_b[kIbChunk];  // l, t, r, b of the box test (:249-268)
    __shared__ float2 ctr[kIbChunk];     // gt centre (:284-298)
    const int a = blockIdx.x * kIbThreads + threadIdx.x;
    float s = 0.f, xc = 0.f, yc = 0.f, r = 0.f;
    if (a < A) {
        s = es[a];
        const float xsh = xs[a] * s, ysh = ys[a] * s;  // :238-239
        xc = xsh + 0.5f * s;                            // :240-247
        yc = ysh + 0.5f * s;
        r = 2.5f * s;                                   // center_radius * stride (:284)
    }
    bool any_box = false, any_ctr = false;
    for (int g0 = 0; g0 < G; g0 += kIbChunk) {
        const int n = min(kIbChunk, G - g0);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += kIbThreads) {
            const float *b = gt + 4 * (size_t)(g0 + i);
            edge_b[i] = make_float4(b[0] - 0.5f * b[2], b[1] - 0.5f * b[3], b[0] + 0.5f * b[2], b[1] + 0.5f * b[3]);
            ctr[i] = make_float2(b[0], b[1]);
        }
        __syncthreads();
        if (a >= A) continue;
        for (int i = 0; i < n; ++i) {
            const float4 e = edge_b[i];
            // bbox_deltas.min(-1).values > 0 (:272-276): all four deltas positive (NaN -> false either way)
            const float b_l = xc - e.x, b_t = yc - e.y, b_r = e.z - xc, b_b = e.w - yc;
            const bool ib = fminf(fminf(b_l, b_t), fminf(b_r, b_b)) > 0.0f && !(b_l != b_l || b_t != b_t || b_r != b_r || b_b != b_b);
            const float2 c = ctr[i];
            const float c_l = xc - (c.x - r), c_r = (c.x + r) - xc, c_t = yc - (c.y - r), c_b = (c.y + r) - yc;  // :287-303
            const bool ic = fminf(fminf(c_l, c_t), fminf(c_r, c_b)) > 0.0f && !(c_l != c_l || c_t != c_t || c_r != c_r || c_b != c_b);
            in_box[(size_t)(g0 + i) * A + a] = ib ? 1 : 0;
            in_ctr[(size_t)(g0 + i) * A + a] = ic ? 1 : 0;
            any_box |= ib;
            any_ctr |= ic;
        }
    }
    if (a < A) fg[a] = (any_box || any_ctr) ? 1 : 0;  // :310
}

// ---- dynamic_k_matching ------------------------------------------------------------------------------------------
// K1: one warp per GT row — the 10 largest IoUs (values, descending), dynamic k with ATen's reduce tree (:336-340),
//     then the k smallest (cost, index) of the row marked in the matching matrix (:341-348; quirk Q3: k >= Nc - 1
//     takes the whole row).
__device__ __forceinline__ void top_insert_desc(float &top, const float x, const int lane) {
    const float up = __shfl_up_sync(0xffffffffu, top, 1);
    if (top < x) top = (lane == 0 || up >= x) ? x : up;
}

__global__ void __launch_bounds__(128) dyn_k_rows_kernel(const float *cost, const float *ious, const int G, const int Nc,
                                                        const int exact_k, uint8_t *M /*[G,Nc], zeroed*/, int32_t *dyn_k /*[G]*/) {
    const int lane = threadIdx.x & 31, g = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (g >= G) return;
    const float *ir = ious + (size_t)g * Nc, *cr = cost + (size_t)g * Nc;
    // lane i = i-th largest IoU (lists start at -inf; fewer than 10 columns leave -inf behind the real values)
    float top = -__int_as_float(0x7f800000);
    float thresh = top;
    for (int n0 = 0; n0 < Nc; n0 += 32) {
        const float v = n0 + lane < Nc ? ir[n0 + lane] : top;
        unsigned m = __ballot_sync(0xffffffffu, n0 + lane < Nc && (v > thresh || n0 < 32));
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            top_insert_desc(top, __shfl_sync(0xffffffffu, v, j), lane);
        }
        thresh = __shfl_sync(0xffffffffu, top, 9);
    }
    // topk_ious.sum(1): ATen's CUDA reduce over the strided [G, kc] slice — 8 lanes: lane t adds elements t, t + bw
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
    int k = max((int)__shfl_sync(0xffffffffu, v, 0), 1);  // .int() truncates; clamp(min=1)
    if (lane == 0) dyn_k[g] = k;
    uint8_t *mr = M + (size_t)g * Nc;
    if (exact_k) k = min(k, Nc);  // yolov7_loss.py:243-247: torch.topk(cost, k, largest=False) — exactly k columns
    else if (!(k < Nc - 1)) {  // :343 — pos_idx is not cut: every candidate of the row
        for (int n = lane; n < Nc; n += 32) mr[n] = 1;
        return;
    }
    // k smallest (cost, index): k rounds of "smallest key above the previous pick" (ties -> lowest index)
    unsigned long long prev = 0ull;
    for (int r = 0; r < k; ++r) {
        unsigned long long best = ~0ull;
        for (int n = lane; n < Nc; n += 32) {
            const unsigned long long key = ((unsigned long long)float_ordered(cr[n]) << 32) | (unsigned)n;
            if ((r == 0 || key > prev) && key < best) best = key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other < best ? other : best;
        }
        if (best == ~0ull) break;
        prev = best;
        if (lane == 0) mr[(int)(best & 0xffffffffu)] = 1;
    }
}

// K2: one thread per candidate column — anchors claimed by several GTs go to the argmin of the cost column over ALL
//     GT rows, first minimum (:352-356); selected flag, matched GT and its IoU (:357-369)
__global__ void __launch_bounds__(256) dyn_k_cols_kernel(const float *cost, const float *ious, const int G, const int Nc,
                                                        uint8_t *M, uint8_t *sel, int32_t *mgt, float *miou) {
    const int n = blockIdx.x * 256 + threadIdx.x;
    if (n >= Nc) return;
    int cnt = 0, first = -1;
    for (int g = 0; g < G; ++g)
        if (M[(size_t)g * Nc + n]) { ++cnt; if (first < 0) first = g; }
    int gsel = first;
    if (cnt > 1) {
        float best = cost[n];
        gsel = 0;
        for (int g = 1; g < G; ++g) {
            const float c = cost[(size_t)g * Nc + n];
            if (c < best) { best = c; gsel = g; }  // torch.min(dim=0): first minimum
        }
        for (int g = 0; g < G; ++g) M[(size_t)g * Nc + n] = g == gsel ? 1 : 0;  // :355-356
    }
    sel[n] = cnt > 0 ? 1 : 0;
    mgt[n] = cnt > 0 ? gsel : -1;
    miou[n] = cnt > 0 ? ious[(size_t)gsel * Nc + n] : 0.f;
}

}  // namespace plyolo

extern "C" int plyolo_in_boxes_info_f32(const float *gt, const float *expanded_strides, const float *x_shifts,
                                        const float *y_shifts, int A, int G, uint8_t *fg_mask, uint8_t *in_boxes,
                                        uint8_t *in_centers, plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(A >= 0 && G >= 0, "negative size");
    if (A == 0) return PLYOLO_OK;
    PLYOLO_REQUIRE(expanded_strides && x_shifts && y_shifts && fg_mask, "null pointer");
    PLYOLO_REQUIRE(G == 0 || (gt && in_boxes && in_centers), "null pointer");
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    in_boxes_kernel<<<(A + kIbThreads - 1) / kIbThreads, kIbThreads, 0, (cudaStream_t)stream>>>(
        gt, expanded_strides, x_shifts, y_shifts, A, G, fg_mask, in_boxes, in_centers);
    PLYOLO_CHECK_LAUNCH("in_boxes_kernel");
    return PLYOLO_OK;
}

extern "C" int plyolo_dynamic_k_matching_f32(const float *cost, const float *ious, int G, int Nc, int exact_k, uint8_t *matching,
                                             int32_t *dynamic_ks, uint8_t *selected, int32_t *matched_gt, float *matched_iou,
                                             plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(G >= 0 && Nc >= 0, "negative size");
    if (G == 0 || Nc == 0) return PLYOLO_OK;
    PLYOLO_REQUIRE(cost && ious && matching && dynamic_ks && selected && matched_gt && matched_iou, "null pointer");
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(matching, 0, (size_t)G * Nc, st) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    dyn_k_rows_kernel<<<(G + 3) / 4, 128, 0, st>>>(cost, ious, G, Nc, exact_k, matching, dynamic_ks);
    PLYOLO_CHECK_LAUNCH("dyn_k_rows_kernel");
    dyn_k_cols_kernel<<<(Nc + 255) / 256, 256, 0, st>>>(cost, ious, G, Nc, matching, selected, matched_gt, matched_iou);
    PLYOLO_CHECK_LAUNCH("dyn_k_cols_kernel");
    return PLYOLO_OK;
}
