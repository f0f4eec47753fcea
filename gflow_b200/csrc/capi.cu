// capi.cu -- library information entry points of libgflow_b200.so (see include/gflow_b200.h).
#include "common.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>

#ifndef GFB_SIMT_EMU
#include <nvtx3/nvToolsExt.h>
#endif

static std::atomic<long long> g_launches{0};
void gfb_internal_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static bool gfb_nvtx_enabled() {
    static const bool on = [] {
        const char* e = getenv("GFB_NVTX");
        return !(e && e[0] == '0');
    }();
    return on;
}
#ifndef GFB_SIMT_EMU
GfbRange::GfbRange(const char* name) : on(gfb_nvtx_enabled()) {
    if (on) nvtxRangePushA(name);
}
GfbRange::~GfbRange() {
    if (on) nvtxRangePop();
}
#else
GfbRange::GfbRange(const char*) : on(gfb_nvtx_enabled()) {}
GfbRange::~GfbRange() {}
#endif

namespace {
// K hand-off: kernels store K straight into mapped pinned host memory (no D2H copy in the stream,
// which would stall the following kernels behind the copy engine); an event recorded right after
// the producing kernel tells the host when the word is valid.
struct HostSync {
    int32_t* pinned = nullptr;   // host view
    int32_t* mapped = nullptr;   // device view of the same word
    cudaEvent_t ev = nullptr;
    bool pending = false;
};
std::mutex g_sync_mutex;
HostSync g_sync[64];
}  // namespace

int gfb_internal_host_sync(int32_t** pinned, int32_t** mapped, cudaEvent_t* ev) {
    int dev = 0;
    GFB_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GFB_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_sync_mutex);
    HostSync& h = g_sync[dev];
    if (!h.pinned) {
        GFB_TRY(cudaHostAlloc((void**)&h.pinned, 4 * sizeof(int32_t), cudaHostAllocMapped));
        GFB_TRY(cudaHostGetDevicePointer((void**)&h.mapped, h.pinned, 0));
        GFB_TRY(cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming));
    }
    *pinned = h.pinned;
    *mapped = h.mapped;
    *ev = h.ev;
    h.pending = true;
    return 0;
}

extern "C" {

int gfb_version(void) { return 100; }

int gfb_wait_k(int64_t* K_host) {
    if (!K_host) return GFB_E_BADARG;
    int dev = 0;
    GFB_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !g_sync[dev].pinned) return GFB_E_BADARG;
    GFB_TRY(cudaEventSynchronize(g_sync[dev].ev));
    *K_host = (int64_t)g_sync[dev].pinned[0];
    return 0;
}

int64_t gfb_kernel_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

const char* gfb_build_arch(void) { return "sm_100a"; }

const char* gfb_error_string(int code) {
    if (code == 0) return "success";
    if (code == GFB_E_BADARG) return "gflow_b200: bad argument (null pointer, negative size or unsupported channel group)";
    if (code == GFB_E_UNSUPPORTED) return "gflow_b200: unsupported configuration";
    if (code == GFB_E_CAPACITY) return "gflow_b200: intersection count exceeds the caller's capacity (retry with a larger buffer)";
    return cudaGetErrorString((cudaError_t)code);
}

}  // extern "C"
