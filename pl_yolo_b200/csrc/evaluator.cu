// evaluator.cu — the per-detection statistics of the VOC evaluator on the device:
// tpfp_default (models/evaluators/eval_voc.py:75-105) with bbox_overlaps (models/utils/bbox.py:97-139) for every
// (image, class) of a batch at once, on the dense detections postprocess / format_outputs already hold on the device.
// The reference runs it per class in a multiprocessing.Pool(8) over numpy arrays (:18-31); here one CTA per image
// does all classes, and what crosses the GPUs afterwards is fixed-size: the TP flags next to the detections
// (all-gathered with them) and the per-class GT counts (all-reduced).
#include "common.cuh"

namespace plyolo {

constexpr int kEvalThreads = 128;

// dets [B,max_det,6] rows (x1,y1,x2,y2,score,class) in score-descending order (what NMS produced: the order
// np.argsort(-score) of :92 visits them), counts [B]; gts [B,Gmax,5] rows (x1,y1,x2,y2,class), gt_counts [B]
__global__ void __launch_bounds__(kEvalThreads) voc_tpfp_kernel(const float *dets, const int32_t *counts, const int max_det,
                                                               const float *gts, const int32_t *gt_counts, const int Gmax,
                                                               const float iou_thr, const int C, uint8_t *tp,
                                                               int32_t *num_gts /*[C]*/) {
    extern __shared__ int first_det[];  // [Gmax] first detection (in score order) that claims the GT (:95-101)
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(max(counts[b], 0), max_det), G = min(max(gt_counts[b], 0), Gmax);
    const float *gb = gts + (size_t)b * Gmax * 5;
    for (int g = tid; g < G; g += kEvalThreads) {
        first_det[g] = 0x7fffffff;
        const int c = (int)gb[5 * g + 4];
        if (c >= 0 && c < C) atomicAdd(&num_gts[c], 1);  // :33-35
    }
    __syncthreads();
    // per detection: max IoU over the GTs of its class and the first GT attaining it (:85-89)
    for (int d0 = 0; d0 < max_det; d0 += kEvalThreads) {
        const int d = d0 + tid;
        float best = -1.f;
        int arg = -1;
        if (d < n) {
            const float *r = dets + ((size_t)b * max_det + d) * 6;
            const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
            const int c = (int)r[5];
            const float area1 = (x2 - x1) * (y2 - y1);
            for (int g = 0; g < G; ++g) {
                if ((int)gb[5 * g + 4] != c) continue;
                const float gx1 = gb[5 * g], gy1 = gb[5 * g + 1], gx2 = gb[5 * g + 2], gy2 = gb[5 * g + 3];
                const float area2 = (gx2 - gx1) * (gy2 - gy1);
                const float ow = fmaxf(fminf(x2, gx2) - fmaxf(x1, gx1), 0.f), oh = fmaxf(fminf(y2, gy2) - fmaxf(y1, gy1), 0.f);
                const float overlap = ow * oh;
                const float uni = fmaxf((area1 + area2) - overlap, 1e-6f);  // bbox.py:134-135
                const float iou = overlap / uni;
                if (iou > best) { best = iou; arg = g; }  // argmax: first maximum
            }
            if (arg >= 0 && best >= iou_thr) atomicMin(&first_det[arg], d);
        }
        __syncthreads();  // (claims of later detections cannot precede earlier ones: atomicMin keeps the first)
        (void)best;
    }
    __syncthreads();
    // a detection is a true positive iff it reaches the threshold and is the first to claim its GT (:94-103)
    for (int d = tid; d < max_det; d += kEvalThreads) {
        uint8_t t = 0;
        if (d < n) {
            const float *r = dets + ((size_t)b * max_det + d) * 6;
            const float x1 = r[0], y1 = r[1], x2 = r[2], y2 = r[3];
            const int c = (int)r[5];
            const float area1 = (x2 - x1) * (y2 - y1);
            float best = -1.f;
            int arg = -1;
            for (int g = 0; g < G; ++g) {
                if ((int)gb[5 * g + 4] != c) continue;
                const float gx1 = gb[5 * g], gy1 = gb[5 * g + 1], gx2 = gb[5 * g + 2], gy2 = gb[5 * g + 3];
                const float area2 = (gx2 - gx1) * (gy2 - gy1);
                const float ow = fmaxf(fminf(x2, gx2) - fmaxf(x1, gx1), 0.f), oh = fmaxf(fminf(y2, gy2) - fmaxf(y1, gy1), 0.f);
                const float overlap = ow * oh;
                const float iou = overlap / fmaxf((area1 + area2) - overlap, 1e-6f);
                if (iou > best) { best = iou; arg = g; }
            }
            t = (arg >= 0 && best >= iou_thr && first_det[arg] == d) ? 1 : 0;
        }
        tp[(size_t)b * max_det + d] = t;
    }
}

}  // namespace plyolo

extern "C" int plyolo_voc_tpfp_f32(const float *dets, const int32_t *counts, int B, int max_det, const float *gts,
                                   const int32_t *gt_counts, int Gmax, double iou_thr, int C, uint8_t *tp, int32_t *num_gts,
                                   plyolo_stream_t stream) {
    using namespace plyolo;
    PLYOLO_REQUIRE(B >= 0 && max_det >= 1 && Gmax >= 0 && C >= 1, "bad size");
    if (B == 0) return PLYOLO_OK;
    PLYOLO_REQUIRE(dets && counts && tp && num_gts && gt_counts && (Gmax == 0 || gts), "null pointer");
    PLYOLO_REQUIRE((size_t)Gmax * sizeof(int) <= 48 * 1024, "Gmax=%d: more ground-truth boxes per image than the kernel stages", Gmax);
    int rc = check_device();
    if (rc != PLYOLO_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (cudaMemsetAsync(num_gts, 0, (size_t)C * sizeof(int32_t), st) != cudaSuccess) {
        set_error("cudaMemsetAsync: %s", cudaGetErrorString(cudaGetLastError()));
        return PLYOLO_ERR_CUDA;
    }
    voc_tpfp_kernel<<<B, kEvalThreads, (size_t)(Gmax > 0 ? Gmax : 1) * sizeof(int), st>>>(dets, counts, max_det, gts, gt_counts, Gmax,
                                                                                         (float)iou_thr, C, tp, num_gts);
    PLYOLO_CHECK_LAUNCH("voc_tpfp_kernel");
    return PLYOLO_OK;
}
