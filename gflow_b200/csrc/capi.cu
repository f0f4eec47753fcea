// capi.cu -- library information entry points of libgflow_b200.so (see include/gflow_b200.h).
#include "common.cuh"

#include <atomic>
#include <mutex>

static std::atomic<long long> g_launches{0};
void gfb_internal_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct HostSync {
    int32_t* pinned = nullptr;
    cudaEvent_t ev = nullptr;
};
std::mutex g_sync_mutex;
HostSync g_sync[64];
}  // namespace

int gfb_internal_host_sync(int32_t** pinned, cudaEvent_t* ev) {
    int dev = 0;
    GFB_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GFB_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_sync_mutex);
    HostSync& h = g_sync[dev];
    if (!h.pinned) {
        GFB_TRY(cudaHostAlloc((void**)&h.pinned, 4 * sizeof(int32_t), cudaHostAllocDefault));
        GFB_TRY(cudaEventCreateWithFlags(&h.ev, cudaEventDisableTiming));
    }
    *pinned = h.pinned;
    *ev = h.ev;
    return 0;
}

extern "C" {

int gfb_version(void) { return 100; }

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
