// capi.cu -- library information entry points of libgflow_b200.so (see include/gflow_b200.h).
#include "common.cuh"

#include <atomic>

static std::atomic<long long> g_launches{0};
void gfb_internal_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" {

int gfb_version(void) { return 100; }

int64_t gfb_kernel_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

const char* gfb_build_arch(void) { return "sm_100a"; }

const char* gfb_error_string(int code) {
    if (code == 0) return "success";
    if (code == GFB_E_BADARG) return "gflow_b200: bad argument (null pointer, negative size or unsupported channel group)";
    if (code == GFB_E_UNSUPPORTED) return "gflow_b200: unsupported configuration";
    return cudaGetErrorString((cudaError_t)code);
}

}  // extern "C"
