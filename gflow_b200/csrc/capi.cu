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
//
// Re-entrant: every hand-off takes the next slot of a per-device ring (own pinned word, own event) and is named
// by a ticket = generation << 16 | device << 8 | slot.  Two streams or host threads that interleave
// gfb_render_forward / gfb_sort_gaussian on one GPU therefore never read each other's K, and a caller may
// pick its K up late (gfb_wait_k_ticket / gfb_query_k_ticket); a slot is reused after kRing further hand-offs
// on the device, and a ticket whose slot has been reused reports GFB_E_STALE instead of a foreign value.
constexpr int kRing = 256;
struct KSlot {
    cudaEvent_t ev = nullptr;
    uint32_t gen = 0;
};
struct DevRing {
    int32_t* pinned = nullptr;  // host view of kRing x 4 words
    int32_t* mapped = nullptr;  // device view of the same block
    KSlot slot[kRing];
    uint32_t next = 0;
};
std::mutex g_sync_mutex;
DevRing g_ring[64];
thread_local int64_t t_last_ticket = -1;

bool ticket_parts(int64_t ticket, int& dev, int& slot, uint32_t& gen) {
    if (ticket < 0) return false;
    slot = (int)(ticket & 0xff);
    dev = (int)((ticket >> 8) & 0xff);
    gen = (uint32_t)(ticket >> 16);
    return dev < 64 && slot < kRing;
}
}  // namespace

int gfb_internal_host_sync(int32_t** pinned, int32_t** mapped, cudaEvent_t* ev) {
    int dev = 0;
    GFB_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GFB_E_UNSUPPORTED;
    std::lock_guard<std::mutex> lock(g_sync_mutex);
    DevRing& r = g_ring[dev];
    if (!r.pinned) {
        GFB_TRY(cudaHostAlloc((void**)&r.pinned, (size_t)kRing * 4 * sizeof(int32_t), cudaHostAllocMapped));
        GFB_TRY(cudaHostGetDevicePointer((void**)&r.mapped, r.pinned, 0));
    }
    const int s = (int)(r.next++ % kRing);
    KSlot& k = r.slot[s];
    if (!k.ev) GFB_TRY(cudaEventCreateWithFlags(&k.ev, cudaEventDisableTiming));
    k.gen = (k.gen + 1) & 0x7fffffffu;
    *pinned = r.pinned + 4 * s;
    *mapped = r.mapped + 4 * s;
    *ev = k.ev;
    t_last_ticket = ((int64_t)k.gen << 16) | ((int64_t)dev << 8) | (int64_t)s;
    return 0;
}

static int gfb_k_from_ticket(int64_t ticket, int64_t* K_host, bool block) {
    if (!K_host) return GFB_E_BADARG;
    int dev, s;
    uint32_t gen;
    if (!ticket_parts(ticket, dev, s, gen)) return GFB_E_BADARG;
    cudaEvent_t ev;
    {
        std::lock_guard<std::mutex> lock(g_sync_mutex);
        if (!g_ring[dev].pinned || g_ring[dev].slot[s].gen != gen) return GFB_E_STALE;
        ev = g_ring[dev].slot[s].ev;
    }
    if (block) {
        GFB_TRY(cudaEventSynchronize(ev));
    } else {
        const cudaError_t q = cudaEventQuery(ev);
        if (q == cudaErrorNotReady) return GFB_E_NOTREADY;
        GFB_TRY(q);
    }
    const int32_t k = g_ring[dev].pinned[4 * s];
    std::lock_guard<std::mutex> lock(g_sync_mutex);
    if (g_ring[dev].slot[s].gen != gen) return GFB_E_STALE;  // reused while we were waiting
    *K_host = (int64_t)k;
    return 0;
}

extern "C" {

int gfb_version(void) { return 101; }

int64_t gfb_k_ticket(void) { return t_last_ticket; }
int gfb_wait_k_ticket(int64_t ticket, int64_t* K_host) { return gfb_k_from_ticket(ticket, K_host, true); }
int gfb_query_k_ticket(int64_t ticket, int64_t* K_host) { return gfb_k_from_ticket(ticket, K_host, false); }
int gfb_wait_k(int64_t* K_host) { return gfb_k_from_ticket(t_last_ticket, K_host, true); }

int64_t gfb_kernel_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

const char* gfb_build_arch(void) { return "sm_100a"; }

const char* gfb_error_string(int code) {
    if (code == 0) return "success";
    if (code == GFB_E_BADARG) return "gflow_b200: bad argument (null pointer, negative size or unsupported channel group)";
    if (code == GFB_E_UNSUPPORTED) return "gflow_b200: unsupported configuration";
    if (code == GFB_E_CAPACITY) return "gflow_b200: intersection count exceeds the caller's capacity (retry with a larger buffer)";
    if (code == GFB_E_STALE) return "gflow_b200: K ticket expired (its hand-off slot has been reused by later calls)";
    if (code == GFB_E_NOTREADY) return "gflow_b200: K has not been produced yet";
    return cudaGetErrorString((cudaError_t)code);
}

}  // extern "C"
