// common.cuh — shared device/host helpers for libplyolo (sm_100a only).
//
// Numerics contract (include/plyolo.h): every reference op is its own fp32 rounding.  The whole
// library is compiled with -fmad=false; the only fused multiply-add is the deliberate
// __fmaf_rn in the torchvision-CUDA IoU (postprocess.cu).  Transcendentals are the libdevice
// functions ATen's CUDA kernels call (expf, logf, log1pf, sqrtf) and division is IEEE.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/plyolo.h"

namespace plyolo {

// ---- host side -----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_device();  // PLYOLO_OK or PLYOLO_ERR_NO_DEVICE
// debug hook: records the calling thread's stage event `i` (0 = before the first kernel of an entry point,
// 1 = between its two kernels, 2 = after the last) on `stream` if plyolo_debug_stage_events() armed them
void record_stage_event(int i, cudaStream_t stream);

#define PLYOLO_CHECK_LAUNCH(what)                                                 \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            plyolo::set_error("%s: %s", what, cudaGetErrorString(e__));           \
            return PLYOLO_ERR_CUDA;                                               \
        }                                                                         \
        plyolo::count_launch();                                                   \
    } while (0)

#define PLYOLO_REQUIRE(cond, ...)                                                 \
    do {                                                                          \
        if (!(cond)) {                                                            \
            plyolo::set_error(__VA_ARGS__);                                       \
            return PLYOLO_ERR_INVALID;                                            \
        }                                                                         \
    } while (0)

struct Levels {
    const float *ptr[PLYOLO_MAX_LEVELS];
    int hw[PLYOLO_MAX_LEVELS];         // H*W
    int w[PLYOLO_MAX_LEVELS];          // W
    int off[PLYOLO_MAX_LEVELS];        // first anchor index of the level
    int tile0[PLYOLO_MAX_LEVELS + 1];  // first tile index of the level (per-level tiling)
    float stride[PLYOLO_MAX_LEVELS];
    float inv_w[PLYOLO_MAX_LEVELS];    // 1 / W (fp32)
    int n;
    int A;
};

// Fills `lv` (tiles of `tile` anchors never straddle a level); returns PLYOLO_OK / error.
int make_levels(Levels &lv, const float *const *host_lvl, const int *hs, const int *ws, const int *strides,
                int n_levels, int tile);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// true exactly once per (call site tag, current device): kernel attributes are per device and are set to fixed maxima,
// so no launch ever changes what another host thread relies on (api.cu)
bool first_use_on_device(int tag);

// PLYOLO_NO_PDL=1 launches every kernel as a plain stream-ordered kernel (A/B measurements, debugging)
bool pdl_enabled();

// cudaLaunchKernelEx with (optionally) the programmatic-dependent-launch attribute: the kernel may be scheduled while
// its predecessor in the stream is still running; it orders itself with griddepcontrol.wait (or its own flags)
template <typename... Args>
inline cudaError_t launch_ex(void (*kernel)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                             Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ---- device side ---------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// sigmoid exactly as ATen's CUDA kernel: 1 / (1 + exp(-x)) (bit-equal on B200, tools/probe_aten_cuda.py)
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// order-preserving float -> uint32 (ascending); NaN unspecified
__device__ __forceinline__ uint32_t float_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + 1-D bulk async copy (TMA engine, no tensor map): global -> shared::cta
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace plyolo
