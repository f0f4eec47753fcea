// binning.cu -- msplat.sort_gaussian for sm_100a: tile binning + per-tile segmented sort.
//
// Reference interface: msplat.sort_gaussian(uv, depth, W, H, radius, tiles_touched) ->
// (gaussian_ids_sorted (K,), tile_range (T,2)), called from
// /root/reference/gflow/utils/render.py:52-54,138-140.
//
// The 3DGS lineage emits K 64-bit (tile << 32 | depth bits) keys and runs a global radix
// sort (~6 passes over 12-byte pairs).  The result of that sort is fully determined: for
// every tile, the Gaussians touching it ordered by (depth bits, Gaussian id).  We produce
// exactly that order without the global sort:
//   1. bin_count   : one thread per Gaussian, atomicAdd on its tiles' counters
//   2. tile_scan   : exclusive scan of the T x R counters, K = total, by the LAST CTA of step 1 (same launch);
//                    counters are replicated R ways to spread same-address atomics over R L2 lines
//   3. bin_scatter : one thread per Gaussian, claims a slot in each tile's segment (the
//                    counters of step 1 count back down) and writes the 64-bit key
//                    (depth bits << 32 | id)
//   4. tile_sort   : one 128-thread CTA per tile sorts its segment in registers (direction-free
//                    bitonic network: shuffles below stride 32, double-buffered shared-memory
//                    exchanges across warps, in-thread above 128; <= 512 keys); larger segments
//                    fall to a generic shared-memory / in-place pass
// Steps 1 and 3 walk the flattened (Gaussian, tile) list 32 pairs per warp round (WarpTileWalk), so
// their atomics are load balanced and each lane waits for one atomic per round, not one per tile of
// its own Gaussian in sequence.
// K is ~3 N and segments are ~100 entries, so everything after step 1 stays in L2/SMEM.
#include "common.cuh"
#include "sort_network.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool gaussian_rect(const float2* __restrict__ uv, const int32_t* __restrict__ radius,
                                              const int32_t* __restrict__ tiles_touched, int i, int gx, int gy,
                                              int& x0, int& y0, int& x1, int& y1) {
    const int r = radius[i];
    if (r <= 0 || tiles_touched[i] <= 0) return false;
    const float2 p = uv[i];
    gfb_tile_rect(p.x, p.y, (float)r, gx, gy, x0, y0, x1, y1);
    return (x1 > x0) && (y1 > y0);
}

// Count, and the last CTA to finish scans: one launch instead of two, and the host (which waits for K to size the
// result) is released one kernel earlier.  `done` is the ticket word behind the counters (zeroed with them).
static_assert(kThreads == kScanThreads, "the counting CTA doubles as the scanning CTA");
__global__ void __launch_bounds__(kThreads)
bin_count_scan_kernel(const float2* __restrict__ uv, const int32_t* __restrict__ radius,
                      const int32_t* __restrict__ tiles_touched, int N, int gx, int gy, int R,
                      int32_t* __restrict__ counts, int32_t* __restrict__ done, int32_t* __restrict__ offsets,
                      int32_t* __restrict__ k_mapped) {
    __shared__ int s_buf[kScanSmemInts];
    __shared__ int s_warp[34];
    __shared__ bool s_last;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (i >= N || !gaussian_rect(uv, radius, tiles_touched, i, gx, gy, x0, y0, x1, y1)) x1 = x0, y1 = y0;
    const WarpTileWalk walk(x0, y0, x1 - x0, y1 - y0, gx, threadIdx.x & 31);
    for (int base = 0; base < walk.total; base += 32) {
        int owner;
        const int t = walk.item(base, owner);
        if (t >= 0) red_add_s32(counts + t * R + (blockIdx.x % R), 1);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(done, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total = cta_exclusive_scan(counts, gx * gy * R, offsets, s_buf, s_warp);
    if (threadIdx.x == 0) {
        *done = 0;  // the ticket is back at zero and bin_scatter counts the counters back down: the block cleans itself
        if (k_mapped) {
            *k_mapped = total;  // mapped pinned host word
            __threadfence_system();
        }
    }
}

// Single CTA: exclusive scan of counts[0..n) into offsets[0..n].
__global__ void __launch_bounds__(kScanThreads)
tile_scan_kernel(const int32_t* __restrict__ counts, int n, int32_t* __restrict__ offsets,
                 int32_t* __restrict__ k_mapped) {
    __shared__ int s_buf[kScanSmemInts];
    __shared__ int s_warp[34];
    const int total = cta_exclusive_scan(counts, n, offsets, s_buf, s_warp);
    if (threadIdx.x == 0 && k_mapped) {
        *k_mapped = total;  // mapped pinned host word
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(kThreads)
bin_scatter_kernel(const float2* __restrict__ uv, const float* __restrict__ depth,
                   const int32_t* __restrict__ radius, const int32_t* __restrict__ tiles_touched, int N, int gx,
                   int gy, int R, const int32_t* __restrict__ offsets, int32_t* __restrict__ counts,
                   unsigned long long* __restrict__ keys, long long K) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    unsigned int dbits = 0;
    if (i >= N || !gaussian_rect(uv, radius, tiles_touched, i, gx, gy, x0, y0, x1, y1)) x1 = x0, y1 = y0;
    else dbits = __float_as_uint(depth[i]);
    const WarpTileWalk walk(x0, y0, x1 - x0, y1 - y0, gx, lane);
    for (int base = 0; base < walk.total; base += 32) {
        int owner;
        const int t = walk.item(base, owner);
        const unsigned int o_bits = __shfl_sync(0xffffffffu, dbits, owner);
        if (t >= 0) {
            // the phase-1 counters double as countdown cursors: slot = offset + (count-- - 1)
            const int slot = t * R + (blockIdx.x % R);
            const long long pos = (long long)offsets[slot] + (atomicSub(counts + slot, 1) - 1);
            const unsigned int id = (unsigned int)(i - lane + owner);
            if (pos >= 0 && pos < K) keys[pos] = ((unsigned long long)o_bits << 32) | id;  // never overrun
        }
    }
}

__global__ void __launch_bounds__(kSortThreads)
tile_sort_kernel(const int32_t* __restrict__ offsets, int R, unsigned long long* __restrict__ keys,
                 int32_t* __restrict__ ids_sorted, int2* __restrict__ tile_range, int T, long long capacity) {
    __shared__ unsigned long long s_keys[kSortSmemSmall];
    sort_tile_cta(offsets, R, keys, capacity, s_keys, tile_range,
                   [ids_sorted](long long pos, unsigned long long key) { ids_sorted[pos] = (int32_t)(unsigned int)key; });
}

}  // namespace

extern "C" {

size_t gfb_sort_workspace_bytes(int64_t K) { return K < 0 ? 0 : (size_t)K * sizeof(unsigned long long); }

// tile workspace: counts[T*R] | done ticket | offsets[T*R + 1]
size_t gfb_sort_tile_workspace_bytes(int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    const size_t T = (size_t)((W + GFB_TILE - 1) / GFB_TILE) * ((H + GFB_TILE - 1) / GFB_TILE);
    const size_t R = (size_t)gfb_tile_replicas((int)T);
    return (2 * T * R + 2) * sizeof(int32_t);
}

static int sort_gaussian_impl(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles_touched,
                              int N, int W, int H, void* tile_ws, int64_t capacity, void* keys_ws,
                              int32_t* gaussian_ids_sorted, int32_t* tile_range, int64_t* K_host, void* stream, bool keep) {
    if (N < 0 || W <= 0 || H <= 0 || capacity < 0 || !tile_ws || !tile_range || !K_host) return GFB_E_BADARG;
    if (N > 0 && (!uv || !depth || !radius || !tiles_touched)) return GFB_E_BADARG;
    if (capacity > 0 && (!keys_ws || !gaussian_ids_sorted)) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    int32_t* counts = (int32_t*)tile_ws;
    int32_t* done = counts + (size_t)T * R;
    int32_t* offsets = done + 1;
    unsigned long long* keys = (unsigned long long*)keys_ws;
    int32_t *pinned = nullptr, *mapped = nullptr;
    cudaEvent_t ev = nullptr;
    int rc = gfb_internal_host_sync(&pinned, &mapped, &ev);
    if (rc) return rc;
    // keep: counters + ticket are zero on entry (the caller zeroed the block once) and zero again when the call's
    // kernels have run -- the scatter counts every counter back down whatever the capacity -- so no memset per call
    if (!keep) GFB_TRY(cudaMemsetAsync(counts, 0, sizeof(int32_t) * ((size_t)T * R + 1), st));
    if (N > 0) {
        bin_count_scan_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(
            reinterpret_cast<const float2*>(uv), radius, tiles_touched, N, gx, gy, R, counts, done, offsets, mapped);
    } else {
        tile_scan_kernel<<<1, kScanThreads, 0, st>>>(counts, T * R, offsets, mapped);
    }
    GFB_CHECK_LAUNCH();
    GFB_TRY(cudaEventRecord(ev, st));
    // speculative: scatter + per-tile sort are enqueued with the caller's capacity before K is known
    if (N > 0 && (capacity > 0 || keep)) {
        bin_scatter_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(
            reinterpret_cast<const float2*>(uv), depth, radius, tiles_touched, N, gx, gy, R, offsets, counts, keys,
            (long long)capacity);
        GFB_CHECK_LAUNCH();
    }
    tile_sort_kernel<<<T, kSortThreads, 0, st>>>(
        offsets, R, keys, gaussian_ids_sorted, reinterpret_cast<int2*>(tile_range), T, (long long)capacity);
    GFB_CHECK_LAUNCH();
    GFB_TRY(cudaEventSynchronize(ev));  // waits for count + scan only
    *K_host = (int64_t)pinned[0];
    return (*K_host > capacity) ? GFB_E_CAPACITY : 0;
}

int gfb_sort_gaussian(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles_touched, int N,
                      int W, int H, void* tile_ws, int64_t capacity, void* keys_ws, int32_t* gaussian_ids_sorted,
                      int32_t* tile_range, int64_t* K_host, void* stream) {
    return sort_gaussian_impl(uv, depth, radius, tiles_touched, N, W, H, tile_ws, capacity, keys_ws, gaussian_ids_sorted,
                              tile_range, K_host, stream, false);
}

int gfb_sort_gaussian_keep(const float* uv, const float* depth, const int32_t* radius, const int32_t* tiles_touched,
                           int N, int W, int H, void* tile_ws_keep, int64_t capacity, void* keys_ws,
                           int32_t* gaussian_ids_sorted, int32_t* tile_range, int64_t* K_host, void* stream) {
    return sort_gaussian_impl(uv, depth, radius, tiles_touched, N, W, H, tile_ws_keep, capacity, keys_ws,
                              gaussian_ids_sorted, tile_range, K_host, stream, true);
}

}  // extern "C"
