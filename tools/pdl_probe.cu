// pdl_probe.cu — does a dependent grid launched with programmatic stream serialization (PDL) become
// co-resident with a persistent one-CTA-per-SM primary grid on B200, eagerly and under CUDA-graph replay?
//
// primary : 148 CTAs x 576 threads, 174 KB dynamic shared memory, <= 64 registers: triggers its dependents at
//           once, then raises flag[i] at i * T / NF microseconds (release), like the score kernel's per-image
//           completion counters.
// secondary: NF clusters of 4 CTAs x 448 threads, 44 KB dynamic shared memory: thread 0 spins (acquire) on
//           flag[blockIdx.y], then every thread "works" for L microseconds; griddepcontrol.wait before exit.
// Prints the event time of the pair and the secondary CTAs' start / flag / end timestamps relative to the
// primary's start.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

constexpr int NF = 32;

__global__ void __maxnreg__(64)
primary(int *flags, unsigned long long *t0_out, int total_us, int epoch) {
    extern __shared__ float smem[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gtime();
    if (blockIdx.x == 0 && threadIdx.x == 0) *t0_out = t0;
    smem[threadIdx.x] = 1.f;
    __syncthreads();
    // CTA i raises flag i (i < NF) at its time; every CTA stays for total_us
    const unsigned long long mine = t0 + (unsigned long long)total_us * 1000ull * (blockIdx.x + 1) / NF;
    const unsigned long long end = t0 + (unsigned long long)total_us * 1000ull;
    bool raised = blockIdx.x >= NF;
    while (true) {
        const unsigned long long now = gtime();
        if (!raised && now >= mine && threadIdx.x == 0) {
            __threadfence();
            atomicExch(&flags[blockIdx.x], epoch);
            raised = true;
        }
        if (now >= end) break;
        __nanosleep(100);
    }
    if (!raised && threadIdx.x == 0) atomicExch(&flags[blockIdx.x], epoch);
}

template <int CL>
__device__ __forceinline__ void secondary_body(const int *flags, unsigned long long *stamps, int work_us, int epoch);

__global__ void __cluster_dims__(4, 1, 1) __maxnreg__(64)
secondary(const int *flags, unsigned long long *stamps, int work_us, int epoch) { secondary_body<1>(flags, stamps, work_us, epoch); }
__global__ void __maxnreg__(64)
secondary_nc(const int *flags, unsigned long long *stamps, int work_us, int epoch) { secondary_body<0>(flags, stamps, work_us, epoch); }

template <int CL>
__device__ __forceinline__ void secondary_body(const int *flags, unsigned long long *stamps, int work_us, int epoch) {
    extern __shared__ float smem[];
    const int b = blockIdx.y;
    const unsigned long long ts = gtime();
    if (threadIdx.x == 0) {
        int v;
        do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + b) : "memory");
            if (v != epoch) __nanosleep(64);
        } while (v != epoch);
    }
    __syncthreads();
    const unsigned long long tf = gtime();
    smem[threadIdx.x] = (float)tf;
    while (gtime() < tf + (unsigned long long)work_us * 1000ull) __nanosleep(100);
    const unsigned long long te = gtime();
    if (threadIdx.x == 0) {
        unsigned long long *s = stamps + ((size_t)b * 4 + blockIdx.x) * 3;
        s[0] = ts; s[1] = tf; s[2] = te;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

static int g_nocluster = 0;
static void launch_pair(cudaStream_t st, int *flags, unsigned long long *t0, unsigned long long *stamps, int total_us,
                        int work_us, int epoch, bool pdl) {
    primary<<<148, 576, 174 * 1024, st>>>(flags, t0, total_us, epoch);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4, NF);
    cfg.blockDim = dim3(448);
    cfg.dynamicSmemBytes = 44 * 1024;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    if (g_nocluster) CK(cudaLaunchKernelEx(&cfg, secondary_nc, (const int *)flags, stamps, work_us, epoch));
    else CK(cudaLaunchKernelEx(&cfg, secondary, (const int *)flags, stamps, work_us, epoch));
}

static void report(const char *name, float ms, unsigned long long t0, const std::vector<unsigned long long> &s) {
    double smin = 1e30, smax = 0, fmax = 0, emax = 0;
    for (size_t i = 0; i < s.size(); i += 3) {
        smin = std::min(smin, (double)(long long)(s[i] - t0));
        smax = std::max(smax, (double)(long long)(s[i] - t0));
        fmax = std::max(fmax, (double)(long long)(s[i + 1] - t0));
        emax = std::max(emax, (double)(long long)(s[i + 2] - t0));
    }
    printf("%-28s pair %.2f us | secondary CTA start %.2f .. %.2f us, last flag seen %.2f us, last end %.2f us (after primary start)\n",
           name, ms * 1e3, smin * 1e-3, smax * 1e-3, fmax * 1e-3, emax * 1e-3);
}

int main(int argc, char **argv) {
    const int total_us = argc > 1 ? atoi(argv[1]) : 22, work_us = argc > 2 ? atoi(argv[2]) : 10;
    int *flags; unsigned long long *t0, *stamps;
    CK(cudaMalloc(&flags, 256 * sizeof(int)));
    CK(cudaMemset(flags, 0, 256 * sizeof(int)));
    CK(cudaMalloc(&t0, 8));
    CK(cudaMalloc(&stamps, NF * 4 * 3 * 8));
    CK(cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 174 * 1024));
    CK(cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024));
    CK(cudaFuncSetAttribute(secondary_nc, cudaFuncAttributeMaxDynamicSharedMemorySize, 44 * 1024));
    const int carve = argc > 3 ? atoi(argv[3]) : 0;
    g_nocluster = argc > 4 ? atoi(argv[4]) : 0;
    if (carve) {
        CK(cudaFuncSetAttribute(primary, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(secondary, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(secondary_nc, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    printf("carveout max: %d, cluster: %d\n", carve, !g_nocluster);
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, primary));
    printf("primary regs %d, static smem %zu\n", fa.numRegs, fa.sharedSizeBytes);
    CK(cudaFuncGetAttributes(&fa, secondary));
    printf("secondary regs %d, static smem %zu\n", fa.numRegs, fa.sharedSizeBytes);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<unsigned long long> hs(NF * 4 * 3);
    unsigned long long ht0;
    int epoch = 0;
    for (int pdl = 0; pdl <= 1; ++pdl) {
        for (int rep = 0; rep < 3; ++rep) {
            ++epoch;
            CK(cudaEventRecord(e0, st));
            launch_pair(st, flags, t0, stamps, total_us, work_us, epoch, pdl != 0);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(hs.data(), stamps, hs.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&ht0, t0, 8, cudaMemcpyDeviceToHost));
            report(pdl ? "eager PDL" : "eager plain", ms, ht0, hs);
        }
    }
    // graph: capture 4 pairs back to back (the epoch is baked in: flags are reset by a memset node per pair)
    for (int pdl = 0; pdl <= 1; ++pdl) {
        cudaGraph_t g; cudaGraphExec_t ge;
        CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < 4; ++k) {
            CK(cudaMemsetAsync(flags, 0, 256 * sizeof(int), st));
            launch_pair(st, flags, t0, stamps, total_us, work_us, 7, pdl != 0);
        }
        CK(cudaStreamEndCapture(st, &g));
        CK(cudaGraphInstantiate(&ge, g, 0));
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0, st));
            CK(cudaGraphLaunch(ge, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            CK(cudaMemcpy(hs.data(), stamps, hs.size() * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&ht0, t0, 8, cudaMemcpyDeviceToHost));
            report(pdl ? "graph x4 PDL (per pair)" : "graph x4 plain (per pair)", ms / 4, ht0, hs);
        }
        CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
    }
    printf("done\n");
    return 0;
}
