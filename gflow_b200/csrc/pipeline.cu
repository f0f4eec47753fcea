// pipeline.cu -- fused render pipeline (msplat.rasterization) for sm_100a.
//
// The op-level surface GFlow calls (render.py:21-64) launches ten kernels and reads K back in the
// middle of sort_gaussian.  When the caller only wants the image -- msplat.rasterization, the
// benchmark's render step, the frame-sharded driver -- the same arithmetic runs as four forward
// kernels and two backward kernels:
//
//   forward   preprocess      project + cov3d + EWA per Gaussian (bit-identical to the separate ops,
//                             shared device code in splat_math.cuh), per-tile counting, and the
//                             exclusive scan of the tile counters by the last CTA to finish
//             scatter         claim slots, write (depth bits << 32 | id) keys
//             tile_sort_pack  per-tile segmented sort; writes gaussian_ids_sorted AND the packed
//                             A / B / F record streams the blend kernels read by TMA
//             blend_fwd       (blend.cu)
//   backward  blend_bwd       (blend.cu) -> packed 48-byte gradient records
//             geometry_bwd    unpack + EWA + cov3d + projection backward per Gaussian, camera
//                             gradients block-reduced
//
// K is not known on the host when the sort / pack / blend kernels are enqueued: the caller passes a
// capacity (its previous K plus slack), the kernels clamp to it, and `preprocess` stores K straight
// into mapped pinned host memory.  The host waits on that copy only -- the GPU keeps working on
// the speculatively enqueued kernels -- and a call whose K exceeded the capacity returns
// GFB_E_CAPACITY so the caller can retry with a larger buffer.
#include <cstdlib>

#include "records.cuh"
#include "sort_network.cuh"
#include "splat_math.cuh"

bool gfb_tight_tiles();

// blend.cu
int gfb_internal_blend_fwd(const void*, const void*, int64_t, const int32_t*, int, int, int, float, int, int, float*,
                           float*, int32_t*, void*, bool pdl);
int gfb_internal_sort_pack_blend_fwd(const int32_t*, int, void*, int32_t*, int64_t, const float*, const float*, const float*,
                                     const float*, int, int32_t*, void*, void*, float, int, int, float*, float*, int32_t*,
                                     void*, bool pdl);
int gfb_internal_blend_bwd(const void*, const void*, int64_t, const int32_t*, const int32_t*, int, int, int, float, int, int,
                           const float*, const int32_t*, const float*, float*, void*, bool no_rgb, float* zero16);

namespace {

using namespace gfbm;

// ctrl words that follow the T tile counters in the control buffer
enum { CTRL_DONE = GFB_CTRL_DONE, CTRL_K = GFB_CTRL_K, CTRL_WORDS = GFB_CTRL_WORDS };

// TIGHT (opt-in, GFB_TIGHT_TILES=1): bin by the alpha >= 1/255 box instead of the 3-sigma rectangle (tighten_rect)
template <bool TIGHT>
__global__ void __launch_bounds__(kThreads)
preprocess_kernel(const float* __restrict__ xyz, const float* __restrict__ scale, const float4* __restrict__ rotate,
                  const float* __restrict__ intr, const float* __restrict__ extr, int N, int W, int H, float nearest,
                  float extent, float2* __restrict__ uv, float* __restrict__ depth, float* __restrict__ conic,
                  int32_t* __restrict__ radius, ushort4* __restrict__ rect, int32_t* __restrict__ counts,
                  int32_t* __restrict__ offsets, int32_t* __restrict__ ctrl, int32_t* __restrict__ k_mapped, int T, int R,
                  const float* __restrict__ opacity) {
    gfb_pdl_launch_dependents();  // scatter may take SM slots while the counting tail drains
    __shared__ float s_cam[16];
    __shared__ int s_scan[34];
    __shared__ int s_buf[kScanSmemInts];
    static_assert(kThreads == kScanThreads, "the last CTA of preprocess runs the scan");
    __shared__ bool s_last;
    load_camera(s_cam, intr, extr);
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    if (i < N) {
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        float u, v, xc, yc, zc;
        const bool ok = project_one(s_cam + 12, s_cam, W, H, nearest, extent, p[0], p[1], p[2], u, v, xc, yc, zc);
        float ca = 0.0f, cb = 0.0f, cc = 0.0f;
        int rad = 0;
        if (ok) {
            const float s[3] = {scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]};
            float S[6];
            cov3d_fwd_one(s, rotate[i], S);
            EwaMid m;
            ewa_mid_eval(p, S, s_cam + 12, s_cam, W, H, m);
            float rf;
            int x0, y0, x1, y1;
            if (ewa_live(m, u, v, gx, gy, rf, x0, y0, x1, y1)) {
                const float dinv = 1.0f / m.det;
                ca = m.c * dinv;
                cb = -m.b * dinv;
                cc = m.a * dinv;
                rad = (int)rf;
                if constexpr (TIGHT) {
                    if (tighten_rect(u, v, ca, cb, cc, opacity[i], x0, y0, x1, y1))
                        rc = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
                } else {
                    rc = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
                }
            }
        }
        uv[i] = ok ? make_float2(u, v) : make_float2(0.0f, 0.0f);
        depth[i] = ok ? zc : 0.0f;
        conic[3 * i] = ca;
        conic[3 * i + 1] = cb;
        conic[3 * i + 2] = cc;
        radius[i] = rad;
        rect[i] = rc;
    }
    {   // per-tile counting, 32 (Gaussian, tile) pairs per warp round
        const WarpTileWalk walk(rc.x, rc.y, rc.z - rc.x, rc.w - rc.y, gx, threadIdx.x & 31);
        for (int base = 0; base < walk.total; base += 32) {
            int owner;
            const int t = walk.item(base, owner);
            if (t >= 0) red_add_s32(counts + t * R + (blockIdx.x % R), 1);
        }
    }
    // ---- the last CTA to get here scans the tile counters (threadfence reduction pattern)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(ctrl + CTRL_DONE, 1);
        s_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total = cta_exclusive_scan(counts, T * R, offsets, s_buf, s_scan);
    if (threadIdx.x == 0) {
        ctrl[CTRL_K] = total;
        ctrl[CTRL_DONE] = 0;  // the ticket is back at zero; scatter hands the counters back: the block cleans itself
        if (k_mapped) *k_mapped = total;  // mapped pinned host word: no D2H copy in the stream
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(kThreads)
scatter_kernel(const ushort4* __restrict__ rect, const float* __restrict__ depth, int N, int gx, int R,
               const int32_t* __restrict__ offsets, int32_t* __restrict__ counts,
               unsigned long long* __restrict__ keys, long long capacity) {
    gfb_pdl_launch_dependents();
    gfb_pdl_wait();  // rect / depth / offsets come from preprocess
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    unsigned int dbits = 0;
    if (i < N) {
        rc = rect[i];
        dbits = __float_as_uint(depth[i]);
    }
    const WarpTileWalk walk(rc.x, rc.y, rc.z - rc.x, rc.w - rc.y, gx, lane);
    for (int base = 0; base < walk.total; base += 32) {
        int owner;
        const int t = walk.item(base, owner);
        const unsigned int o_bits = __shfl_sync(0xffffffffu, dbits, owner);
        if (t >= 0) {
            const int slot = t * R + (blockIdx.x % R);
            const long long pos = (long long)offsets[slot] + (atomicSub(counts + slot, 1) - 1);
            const unsigned int id = (unsigned int)(i - lane + owner);
            if (pos >= 0 && pos < capacity) keys[pos] = ((unsigned long long)o_bits << 32) | id;
        }
    }
}

// record writers: records.cuh
using PackArgs = GfbPackArgs;

__global__ void __launch_bounds__(kSortThreads)
tile_sort_pack_kernel(const int32_t* __restrict__ offsets, int R, unsigned long long* __restrict__ keys,
                      int2* __restrict__ tile_range, int T, long long capacity, PackArgs pa) {
    __shared__ unsigned long long s_keys[kSortSmemSmall];
    gfb_pdl_launch_dependents();
    gfb_pdl_wait();  // keys come from scatter
    sort_tile_cta(offsets, R, keys, capacity, s_keys, tile_range,
                   [pa](long long pos, unsigned long long key) { gfb_write_record(pa, pos, (int)(unsigned int)key); });
}

// Fused geometry backward: grad_pack (from blend_bwd) -> parameter gradients.
__global__ void __launch_bounds__(kThreads)
geometry_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ scale, const float4* __restrict__ rotate,
                    const float* __restrict__ intr, const float* __restrict__ extr, int N, int W, int H, float nearest,
                    float extent, int C, float4* __restrict__ grad_pack, int clear_pack, float* __restrict__ d_xyz,
                    float* __restrict__ d_scale, float4* __restrict__ d_rotate, float* __restrict__ d_opacity,
                    float* __restrict__ d_feature, float* __restrict__ d_cam) {
    __shared__ float s_cam[16];
    load_camera(s_cam, intr, extr);
    const float* e = s_cam;
    const float* in = s_cam + 12;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    // Everything that does not need the gradient pack -- the Gaussian's own forward quantities -- comes BEFORE the wait on
    // blend_bwd: under programmatic dependent launch these CTAs take the SM slots blend_bwd's last wave leaves idle
    // (a third of its run time) and have their recomputation done when the pack is complete.
    float p[3] = {0.0f, 0.0f, 0.0f}, s[3] = {0.0f, 0.0f, 0.0f}, S[6];
    float4 q = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
    float u = 0.0f, v = 0.0f, xc = 0.0f, yc = 0.0f, zc = 0.0f;
    EwaMid m;
    bool seen = false, live = false;
    if (i < N) {
        p[0] = xyz[3 * i], p[1] = xyz[3 * i + 1], p[2] = xyz[3 * i + 2];
        seen = project_one(in, e, W, H, nearest, extent, p[0], p[1], p[2], u, v, xc, yc, zc);
        if (seen) {
            s[0] = scale[3 * i], s[1] = scale[3 * i + 1], s[2] = scale[3 * i + 2];
            q = rotate[i];
            cov3d_fwd_one(s, q, S);
            ewa_mid_eval(p, S, in, e, W, H, m);
            float rf;
            int x0, y0, x1, y1;
            live = ewa_live(m, u, v, gx, gy, rf, x0, y0, x1, y1);
        }
    }
    gfb_pdl_wait();  // grad_pack comes from blend_bwd
    if (i < N) {
        const float4 g0 = grad_pack[3 * (size_t)i], g1 = grad_pack[3 * (size_t)i + 1], g2 = grad_pack[3 * (size_t)i + 2];
        if (clear_pack) {  // a kept pack is handed back zeroed: the next backward accumulates into it without a memset
            const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            grad_pack[3 * (size_t)i] = z;
            grad_pack[3 * (size_t)i + 1] = z;
            grad_pack[3 * (size_t)i + 2] = z;
        }
        float dp[3] = {0.0f, 0.0f, 0.0f}, ds[3] = {0.0f, 0.0f, 0.0f};
        float4 dq = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (seen) {
            if (live) {
                float dS[6];
                ewa_bwd_one(m, p, S, in, e, g0.z, g0.w, g1.x, dp, dS, acc);
                cov3d_bwd_one(s, q, dS, ds, dq);
            }
            project_bwd_one(in, e, p[0], p[1], p[2], xc, yc, zc, g0.x, g0.y, 0.0f, dp, acc);
        }
        d_xyz[3 * i] = dp[0];
        d_xyz[3 * i + 1] = dp[1];
        d_xyz[3 * i + 2] = dp[2];
        d_scale[3 * i] = ds[0];
        d_scale[3 * i + 1] = ds[1];
        d_scale[3 * i + 2] = ds[2];
        d_rotate[i] = dq;
        d_opacity[i] = g1.y;
        float* f = d_feature + (size_t)i * C;
        f[0] = g1.z;
        if (C > 1) f[1] = g1.w;
        if (C > 2) f[2] = g2.x;
        if (C > 3) f[3] = g2.y;
    }
    block_reduce_atomic<16>(acc, d_cam);
}

}  // namespace

// Tile culling by the alpha >= 1/255 box in the fused pipeline and the native fit loop: their gaussian_ids_sorted /
// tile_range list only the (Gaussian, tile) pairs that can reach alpha 1/255 in that tile -- fewer than the 3-sigma
// rectangle msplat.sort_gaussian (and gfb_sort_gaussian) enumerates; images and gradients are unchanged.  Measured on
// a B200 (round 2): K 197 461 -> 148 565 at config 2, render step +5 %, native loop +2.5 %, every parity case green.
// GFB_TIGHT_TILES=0 restores the 3-sigma rule (the CPU-emulated suite runs both).
bool gfb_tight_tiles() {
    static const bool on = [] {
        const char* e = getenv("GFB_TIGHT_TILES");
        return !(e && e[0] == '0');
    }();
    return on;
}

// GFB_FUSE_SORT_BLEND=0: run the per-tile sort + pack and the forward blend as two kernels (the round-1 structure)
// instead of tile_sort_blend_fwd_kernel (blend.cu).
bool gfb_fuse_sort_blend() {
    static const bool on = [] {
        const char* e = getenv("GFB_FUSE_SORT_BLEND");
        return !(e && e[0] == '0');
    }();
    return on;
}

static int launch_scatter(const void* rect_ws, const float* depth, int N, int W, int H, void* control_ws, int64_t capacity,
                          void* keys_ws, void* stream, bool pdl) {
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    int32_t* counts = (int32_t*)control_ws;
    int32_t* tile_offsets = counts + (size_t)T * R + CTRL_WORDS;
    if (N > 0 && capacity > 0) {
        GFB_TRY(gfb_launch_pdl(scatter_kernel, dim3(gfb_div_up(N, kThreads)), dim3(kThreads), st, pdl,
                               reinterpret_cast<const ushort4*>(rect_ws), depth, N, gx, R, tile_offsets, counts,
                               reinterpret_cast<unsigned long long*>(keys_ws), (long long)capacity));
        GFB_CHECK_LAUNCH();
    }
    return 0;
}

// Second and third forward kernels for a caller that has run its own preprocess (gfb_render_forward
// above, the native fit iteration in fit.cu): claim slots + write keys, then per-tile sort + pack.
// control_ws is laid out as gfb_render_control_bytes() describes and holds the scanned offsets.
int gfb_internal_scatter_sort_pack(const void* rect_ws, const float* depth, int N, int W, int H, void* control_ws,
                                   int64_t capacity, void* keys_ws, int32_t* tile_range, const float* uv,
                                   const float* conic, const float* opacity, const float* feature, int C,
                                   int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, void* stream,
                                   bool pdl) {
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    int32_t* counts = (int32_t*)control_ws;
    int32_t* tile_offsets = counts + (size_t)T * R + CTRL_WORDS;
    float4* sA = reinterpret_cast<float4*>(geom_stream);
    int rc = launch_scatter(rect_ws, depth, N, W, H, control_ws, capacity, keys_ws, stream, pdl);
    if (rc) return rc;
    PackArgs pa{reinterpret_cast<const float2*>(uv), conic, opacity, feature, C, sA, sA + capacity,
                reinterpret_cast<float4*>(feat_stream), gaussian_ids_sorted};
    GFB_TRY(gfb_launch_pdl(tile_sort_pack_kernel, dim3(T), dim3(kSortThreads), st, pdl, tile_offsets, R,
                           reinterpret_cast<unsigned long long*>(keys_ws), reinterpret_cast<int2*>(tile_range), T,
                           (long long)capacity, pa));
    GFB_CHECK_LAUNCH();
    return 0;
}

// The same plus the forward blend of all C <= 4 channels: scatter, then ONE kernel that sorts, packs and blends each
// tile (or, with GFB_FUSE_SORT_BLEND=0, tile_sort_pack followed by blend_fwd).
int gfb_internal_scatter_sort_pack_blend(const void* rect_ws, const float* depth, int N, int W, int H, void* control_ws,
                                         int64_t capacity, void* keys_ws, int32_t* tile_range, const float* uv,
                                         const float* conic, const float* opacity, const float* feature, int C,
                                         int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, float bg,
                                         float* out, float* final_T, int32_t* n_contrib, void* stream, bool pdl) {
    if (!gfb_fuse_sort_blend() || C < 1 || C > 4) {
        int rc = gfb_internal_scatter_sort_pack(rect_ws, depth, N, W, H, control_ws, capacity, keys_ws, tile_range, uv, conic,
                                                opacity, feature, C, gaussian_ids_sorted, geom_stream, feat_stream, stream, pdl);
        if (rc) return rc;
        return gfb_internal_blend_fwd(geom_stream, feat_stream, capacity, tile_range, C, 0, C, bg, W, H, out, final_T,
                                      n_contrib, stream, pdl);
    }
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    const int32_t* tile_offsets = (const int32_t*)control_ws + (size_t)T * R + CTRL_WORDS;
    int rc = launch_scatter(rect_ws, depth, N, W, H, control_ws, capacity, keys_ws, stream, pdl);
    if (rc) return rc;
    return gfb_internal_sort_pack_blend_fwd(tile_offsets, R, keys_ws, tile_range, capacity, uv, conic, opacity, feature, C,
                                            gaussian_ids_sorted, geom_stream, feat_stream, bg, W, H, out, final_T, n_contrib,
                                            stream, pdl);
}

extern "C" {

size_t gfb_render_control_k_offset(int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    const size_t T = (size_t)((W + GFB_TILE - 1) / GFB_TILE) * ((H + GFB_TILE - 1) / GFB_TILE);
    return (T * (size_t)gfb_tile_replicas((int)T) + CTRL_K) * sizeof(int32_t);
}

size_t gfb_render_control_bytes(int W, int H) {
    if (W <= 0 || H <= 0) return 0;
    const size_t T = (size_t)((W + GFB_TILE - 1) / GFB_TILE) * ((H + GFB_TILE - 1) / GFB_TILE);
    const size_t R = (size_t)gfb_tile_replicas((int)T);
    // counts[T*R] | ctrl[4] | offsets[T*R + 1]
    return (2 * T * R + 1 + CTRL_WORDS) * sizeof(int32_t);
}

static int render_forward_impl(const float* xyz, const float* scale, const float* rotate, const float* opacity,
                               const float* feature, int C, const float* intr, const float* extr, int N, int W, int H,
                               float bg, float nearest, float extent, float* uv, float* depth, float* conic, int32_t* radius,
                               void* rect_ws, void* control_ws, int32_t* tile_range, int64_t capacity,
                               void* keys_ws, int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, float* out,
                               float* final_T, int32_t* n_contrib, int64_t* K_host, void* stream, bool keep) {
    if (N < 0 || W <= 0 || H <= 0 || C < 1 || C > 4 || capacity < 0) return GFB_E_BADARG;
    if (!intr || !extr || !control_ws || !tile_range || !out || !final_T || !n_contrib)
        return GFB_E_BADARG;
    if (N > 0 && (!xyz || !scale || !rotate || !opacity || !feature || !uv || !depth || !conic || !radius || !rect_ws))
        return GFB_E_BADARG;
    if (capacity > 0 && (!keys_ws || !gaussian_ids_sorted || !geom_stream || !feat_stream)) return GFB_E_BADARG;
    const GfbRange nvtx_range("gfb_render_forward");
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    int32_t* counts = (int32_t*)control_ws;
    int32_t* ctrl = counts + (size_t)T * R;
    int32_t* tile_offsets = ctrl + CTRL_WORDS;
    int32_t *pinned = nullptr, *mapped = nullptr;
    cudaEvent_t ev = nullptr;
    // Captured into a CUDA graph: no K hand-off (an event recorded inside a capture cannot be waited on, and the
    // ring slot would not be the replay's).  K still lands in the control block (gfb_render_control_k_offset), where
    // the owner of the graph reads it after a replay; kernels clamp at `capacity` as always.
    cudaStreamCaptureStatus cap_status = cudaStreamCaptureStatusNone;
    GFB_TRY(cudaStreamIsCapturing(st, &cap_status));
    const bool capturing = cap_status != cudaStreamCaptureStatusNone;
    if (capturing && K_host) return GFB_E_UNSUPPORTED;
    int rc = 0;
    if (!capturing) {
        rc = gfb_internal_host_sync(&pinned, &mapped, &ev);
        if (rc) return rc;
    }
    // keep: the caller's control block persists across calls and cleans itself (scatter returns every counter to zero,
    // the scan resets its ticket); scatter is skipped for N == 0 or capacity == 0, so those calls clear it here
    if (!keep || N == 0 || capacity == 0)
        GFB_TRY(cudaMemsetAsync(control_ws, 0, ((size_t)T * R + CTRL_WORDS) * sizeof(int32_t), st));
    // N == 0 still runs one CTA so the scan zeroes the offsets
    if (gfb_tight_tiles())
        preprocess_kernel<true><<<max(1, gfb_div_up(N, kThreads)), kThreads, 0, st>>>(
            xyz, scale, reinterpret_cast<const float4*>(rotate), intr, extr, N, W, H, nearest, extent,
            reinterpret_cast<float2*>(uv), depth, conic, radius, reinterpret_cast<ushort4*>(rect_ws), counts,
            tile_offsets, ctrl, mapped, T, R, opacity);
    else
        preprocess_kernel<false><<<max(1, gfb_div_up(N, kThreads)), kThreads, 0, st>>>(
            xyz, scale, reinterpret_cast<const float4*>(rotate), intr, extr, N, W, H, nearest, extent,
            reinterpret_cast<float2*>(uv), depth, conic, radius, reinterpret_cast<ushort4*>(rect_ws), counts,
            tile_offsets, ctrl, mapped, T, R, opacity);
    GFB_CHECK_LAUNCH();
    if (!capturing) GFB_TRY(cudaEventRecord(ev, st));
    // speculative part: enqueued before K is known on the host
    rc = gfb_internal_scatter_sort_pack_blend(rect_ws, depth, N, W, H, control_ws, capacity, keys_ws, tile_range, uv, conic,
                                              opacity, feature, C, gaussian_ids_sorted, geom_stream, feat_stream, bg, out,
                                              final_T, n_contrib, stream, true);
    if (rc) return rc;
    if (!K_host) return 0;  // the caller overlaps host work, then calls gfb_wait_k()
    GFB_TRY(cudaEventSynchronize(ev));  // waits for `preprocess` only
    *K_host = (int64_t)pinned[0];
    return (*K_host > capacity) ? GFB_E_CAPACITY : 0;
}

int gfb_render_forward(const float* xyz, const float* scale, const float* rotate, const float* opacity,
                       const float* feature, int C, const float* intr, const float* extr, int N, int W, int H,
                       float bg, float nearest, float extent, float* uv, float* depth, float* conic, int32_t* radius,
                       void* rect_ws, void* control_ws, int32_t* tile_range, int64_t capacity,
                       void* keys_ws, int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, float* out,
                       float* final_T, int32_t* n_contrib, int64_t* K_host, void* stream) {
    return render_forward_impl(xyz, scale, rotate, opacity, feature, C, intr, extr, N, W, H, bg, nearest, extent, uv, depth,
                               conic, radius, rect_ws, control_ws, tile_range, capacity, keys_ws, gaussian_ids_sorted,
                               geom_stream, feat_stream, out, final_T, n_contrib, K_host, stream, false);
}

int gfb_render_forward_keep(const float* xyz, const float* scale, const float* rotate, const float* opacity,
                            const float* feature, int C, const float* intr, const float* extr, int N, int W, int H,
                            float bg, float nearest, float extent, float* uv, float* depth, float* conic, int32_t* radius,
                            void* rect_ws, void* control_ws, int32_t* tile_range, int64_t capacity,
                            void* keys_ws, int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, float* out,
                            float* final_T, int32_t* n_contrib, int64_t* K_host, void* stream) {
    return render_forward_impl(xyz, scale, rotate, opacity, feature, C, intr, extr, N, W, H, bg, nearest, extent, uv, depth,
                               conic, radius, rect_ws, control_ws, tile_range, capacity, keys_ws, gaussian_ids_sorted,
                               geom_stream, feat_stream, out, final_T, n_contrib, K_host, stream, true);
}

size_t gfb_render_grad_bytes(int N) { return N < 0 ? 0 : ((size_t)N * 12 + 16) * sizeof(float); }

static int render_backward_impl(const float* xyz, const float* scale, const float* rotate, const float* intr,
                                const float* extr, int N, int W, int H, int C, float bg, float nearest, float extent,
                                const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int64_t capacity,
                                const void* geom_stream, const void* feat_stream, const float* final_T,
                                const int32_t* n_contrib, const float* g_out, float* grad_pack, float* d_cam, bool keep,
                                float* d_xyz, float* d_scale, float* d_rotate, float* d_opacity, float* d_feature,
                                void* stream) {
    if (N < 0 || W <= 0 || H <= 0 || C < 1 || C > 4 || capacity < 0 || !grad_pack || !d_cam) return GFB_E_BADARG;
    const GfbRange nvtx_range("gfb_render_backward");
    cudaStream_t st = (cudaStream_t)stream;
    if (!keep) {
        GFB_TRY(cudaMemsetAsync(grad_pack, 0, gfb_render_grad_bytes(N), st));  // pack and d_cam are one block here
    } else if (N == 0 || capacity == 0) {
        GFB_TRY(cudaMemsetAsync(d_cam, 0, 16 * sizeof(float), st));  // no blend backward will run to clear it
    }
    if (N == 0) return 0;
    if (!xyz || !scale || !rotate || !intr || !extr || !tile_range || !final_T || !n_contrib || !g_out || !d_xyz ||
        !d_scale || !d_rotate || !d_opacity || !d_feature)
        return GFB_E_BADARG;
    if (capacity > 0) {
        int rc = gfb_internal_blend_bwd(geom_stream, feat_stream, capacity, gaussian_ids_sorted, tile_range, C, 0, C, bg, W,
                                        H, final_T, n_contrib, g_out, grad_pack, stream, false, keep ? d_cam : nullptr);
        if (rc) return rc;
    }
    GFB_TRY(gfb_launch_pdl(geometry_bwd_kernel, dim3(gfb_div_up(N, kThreads)), dim3(kThreads), st, capacity > 0, xyz,
                           scale, reinterpret_cast<const float4*>(rotate), intr, extr, N, W, H, nearest, extent, C,
                           reinterpret_cast<float4*>(grad_pack), keep ? 1 : 0, d_xyz, d_scale,
                           reinterpret_cast<float4*>(d_rotate), d_opacity, d_feature, d_cam));
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_render_backward(const float* xyz, const float* scale, const float* rotate, const float* intr,
                        const float* extr, int N, int W, int H, int C, float bg, float nearest, float extent,
                        const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int64_t capacity,
                        const void* geom_stream, const void* feat_stream, const float* final_T,
                        const int32_t* n_contrib, const float* g_out, void* grad_ws, float* d_xyz, float* d_scale,
                        float* d_rotate, float* d_opacity, float* d_feature, void* stream) {
    if (N < 0 || !grad_ws) return GFB_E_BADARG;
    float* grad_pack = (float*)grad_ws;
    return render_backward_impl(xyz, scale, rotate, intr, extr, N, W, H, C, bg, nearest, extent, gaussian_ids_sorted,
                                tile_range, capacity, geom_stream, feat_stream, final_T, n_contrib, g_out, grad_pack,
                                grad_pack + (size_t)N * 12, false, d_xyz, d_scale, d_rotate, d_opacity, d_feature, stream);
}

int gfb_render_backward_keep(const float* xyz, const float* scale, const float* rotate, const float* intr,
                             const float* extr, int N, int W, int H, int C, float bg, float nearest, float extent,
                             const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int64_t capacity,
                             const void* geom_stream, const void* feat_stream, const float* final_T,
                             const int32_t* n_contrib, const float* g_out, void* grad_pack_keep, float* d_cam, float* d_xyz,
                             float* d_scale, float* d_rotate, float* d_opacity, float* d_feature, void* stream) {
    return render_backward_impl(xyz, scale, rotate, intr, extr, N, W, H, C, bg, nearest, extent, gaussian_ids_sorted,
                                tile_range, capacity, geom_stream, feat_stream, final_T, n_contrib, g_out,
                                (float*)grad_pack_keep, d_cam, true, d_xyz, d_scale, d_rotate, d_opacity, d_feature, stream);
}

}  // extern "C"
