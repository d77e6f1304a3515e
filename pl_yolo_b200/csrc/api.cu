// api.cu — error state, launch accounting and argument helpers shared by the C-ABI entry points.
#include <cstdarg>
#include <cstdio>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace plyolo {

static thread_local char g_err[512] = "";
static thread_local unsigned long long g_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += (unsigned long long)n; }

static thread_local cudaEvent_t g_stage_ev[3] = {nullptr, nullptr, nullptr};
void record_stage_event(int i, cudaStream_t stream) {
    if (g_stage_ev[i]) cudaEventRecord(g_stage_ev[i], stream);
}

int check_device() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("no CUDA device: %s (libplyolo has no CPU fallback)", cudaGetErrorString(e));
        return PLYOLO_ERR_NO_DEVICE;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        set_error("device %d has compute capability %d.x; libplyolo is built for sm_100a only", dev, major);
        return PLYOLO_ERR_NO_DEVICE;
    }
    return PLYOLO_OK;
}

int make_levels(Levels &lv, const float *const *host_lvl, const int *hs, const int *ws, const int *strides,
                int n_levels, int tile) {
    PLYOLO_REQUIRE(host_lvl && hs && ws && strides, "null level description");
    PLYOLO_REQUIRE(n_levels >= 1 && n_levels <= PLYOLO_MAX_LEVELS, "n_levels=%d not in [1,%d]", n_levels,
                   PLYOLO_MAX_LEVELS);
    int off = 0, t0 = 0;
    for (int l = 0; l < n_levels; ++l) {
        PLYOLO_REQUIRE(host_lvl[l] != nullptr, "level %d pointer is null", l);
        PLYOLO_REQUIRE(hs[l] > 0 && ws[l] > 0 && strides[l] > 0, "level %d has a non-positive dimension", l);
        // yolox_loss.py:198 builds the grid with indexing='xy' and views it as (h, w): only square maps
        // decode correctly; refuse the rest instead of reproducing a scrambled grid.
        PLYOLO_REQUIRE(hs[l] == ws[l], "level %d is %dx%d: the reference decode is only defined for square maps",
                       l, hs[l], ws[l]);
        lv.ptr[l] = host_lvl[l];
        lv.hw[l] = hs[l] * ws[l];
        lv.w[l] = ws[l];
        lv.off[l] = off;
        lv.tile0[l] = t0;
        lv.stride[l] = (float)strides[l];
        lv.inv_w[l] = 1.0f / (float)ws[l];
        off += lv.hw[l];
        t0 += (lv.hw[l] + tile - 1) / tile;
    }
    for (int l = n_levels; l < PLYOLO_MAX_LEVELS; ++l) {
        lv.ptr[l] = nullptr; lv.hw[l] = 0; lv.w[l] = 1; lv.off[l] = off; lv.tile0[l] = t0; lv.stride[l] = 1.f; lv.inv_w[l] = 1.f;
    }
    lv.tile0[n_levels] = t0;
    for (int l = n_levels + 1; l <= PLYOLO_MAX_LEVELS; ++l) lv.tile0[l] = t0;
    lv.n = n_levels;
    lv.A = off;
    return PLYOLO_OK;
}

}  // namespace plyolo

namespace plyolo {
bool pdl_enabled() {
    static const bool on = [] {
        const char *v = getenv("PLYOLO_NO_PDL");
        return !(v && v[0] == '1');
    }();
    return on;
}

bool first_use_on_device(int tag) {
    static std::mutex mu;
    static bool done[8][64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::lock_guard<std::mutex> lock(mu);
    if (tag < 0 || tag >= 8 || done[tag][dev]) return false;
    done[tag][dev] = true;
    return true;
}
}  // namespace plyolo

extern "C" {

int plyolo_version(void) { return PLYOLO_VERSION; }

int plyolo_peer_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
    if (!ptr || !handle64 || bytes == 0) { plyolo::set_error("plyolo_peer_alloc: bad argument"); return PLYOLO_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) {
        plyolo::set_error("plyolo_peer_alloc: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PLYOLO_ERR_CUDA;
    }
    memcpy(handle64, &h, 64);
    return PLYOLO_OK;
}

int plyolo_peer_open(const unsigned char *handle64, void **ptr) {
    if (!ptr || !handle64) { plyolo::set_error("plyolo_peer_open: bad argument"); return PLYOLO_ERR_INVALID; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    // opened from the CURRENT device: kernels of this device may then store to the other device's buffer over NVLink
    const cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        plyolo::set_error("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return PLYOLO_ERR_CUDA;
    }
    return PLYOLO_OK;
}

int plyolo_peer_close(void *ptr, int opened) {
    const cudaError_t e = opened ? cudaIpcCloseMemHandle(ptr) : cudaFree(ptr);
    cudaGetLastError();
    return e == cudaSuccess ? PLYOLO_OK : PLYOLO_ERR_CUDA;
}

int plyolo_enable_peer_access(int peer_device) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PLYOLO_ERR_NO_DEVICE;
    if (peer_device == dev) return PLYOLO_OK;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, dev, peer_device) != cudaSuccess || !can) {
        cudaGetLastError();
        plyolo::set_error("device %d cannot access device %d as a peer", dev, peer_device);
        return PLYOLO_ERR_INVALID;
    }
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        plyolo::set_error("cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
        cudaGetLastError();
        return PLYOLO_ERR_CUDA;
    }
    cudaGetLastError();
    return PLYOLO_OK;
}
const char *plyolo_last_error(void) { return plyolo::g_err; }
unsigned long long plyolo_launch_count(void) { return plyolo::g_launches; }
// debug hook (not part of include/plyolo.h): three cudaEvent_t handles (or nulls to disarm) that the two-kernel
// entry points of the calling thread record before / between / after their kernels — bench.py times the
// dominant kernel with them (CUDA events on the launching stream)
void plyolo_debug_stage_events(void *e0, void *e1, void *e2) {
    plyolo::g_stage_ev[0] = static_cast<cudaEvent_t>(e0);
    plyolo::g_stage_ev[1] = static_cast<cudaEvent_t>(e1);
    plyolo::g_stage_ev[2] = static_cast<cudaEvent_t>(e2);
}

}  // extern "C"
