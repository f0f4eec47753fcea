// common.cuh -- shared helpers for the sm_100a splat kernels (gflow_b200).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gflow_b200.h"

#define GFB_ALPHA_MIN (1.0f / 255.0f)
#define GFB_ALPHA_MAX 0.99f
#define GFB_T_EPS 1e-4f
#define GFB_COV_BLUR 0.3f
#define GFB_FRUSTUM_CLAMP 1.3f

// host-side counter of kernels launched by this library (gfb_kernel_launch_count)
void gfb_internal_count_launch();
// per-device mapped pinned int32[4] (host + device view) and event used to hand K to the host
// without draining the stream or putting a copy into it
int gfb_internal_host_sync(int32_t** pinned, int32_t** mapped, cudaEvent_t* ev);

// follows every kernel launch: error check + launch accounting
#define GFB_CHECK_LAUNCH()                      \
    do {                                        \
        cudaError_t e__ = cudaGetLastError();   \
        if (e__ != cudaSuccess) return (int)e__; \
        gfb_internal_count_launch();            \
    } while (0)

#define GFB_TRY(expr)                            \
    do {                                         \
        cudaError_t e__ = (expr);                \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

static inline int gfb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// NVTX ranges around the multi-kernel entry points (render forward / backward, one fit iteration), so that
// `nsys` timelines and `ncu --nvtx --nvtx-include "gfb_fit_iteration/"` can select them.  Header-only NVTX v3: no
// link dependency, a no-op unless a profiler injects itself; set GFB_NVTX=0 to skip even the calls.
struct GfbRange {
    bool on;
    explicit GfbRange(const char* name);
    ~GfbRange();
    GfbRange(const GfbRange&) = delete;
    GfbRange& operator=(const GfbRange&) = delete;
};

// control buffer of the fused pipelines (gfb_render_control_bytes): counts[T*R] | ctrl[4] | offsets[T*R + 1]
enum { GFB_CTRL_DONE = 0, GFB_CTRL_K = 1, GFB_CTRL_WORDS = 4 };

// ---- programmatic dependent launch (PDL): a kernel launched with gfb_launch_pdl may start while
// its predecessor in the stream is still draining; it must call gfb_pdl_wait() before touching
// anything the predecessor wrote.  gfb_pdl_launch_dependents() in the predecessor lets the
// successor's CTAs take SM slots as soon as every predecessor CTA has started.  Both are no-ops in a
// kernel launched the ordinary way.
__device__ __forceinline__ void gfb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gfb_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t gfb_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, bool pdl,
                                         Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// camera-space point with the oracle's operation order: ((e0*x + e1*y) + e2*z) + e3
// (geometry.cu is compiled with -fmad=false so no multiply-add is fused).
__device__ __forceinline__ void gfb_cam_point(const float* __restrict__ e, float x, float y, float z, float& xc,
                                              float& yc, float& zc) {
    xc = ((e[0] * x + e[1] * y) + e[2] * z) + e[3];
    yc = ((e[4] * x + e[5] * y) + e[6] * z) + e[7];
    zc = ((e[8] * x + e[9] * y) + e[10] * z) + e[11];
}

// 3DGS getRect rule on the 16x16 tile grid.
__device__ __forceinline__ void gfb_tile_rect(float u, float v, float r, int gx, int gy, int& x0, int& y0, int& x1,
                                              int& y1) {
    x0 = min(gx, max(0, (int)((u - r) / (float)GFB_TILE)));
    y0 = min(gy, max(0, (int)((v - r) / (float)GFB_TILE)));
    x1 = min(gx, max(0, (int)(((u + r) + (float)(GFB_TILE - 1)) / (float)GFB_TILE)));
    y1 = min(gy, max(0, (int)(((v + r) + (float)(GFB_TILE - 1)) / (float)GFB_TILE)));
}

// Record stream A of the blend kernels = {u, v, extents, id}: `extents` holds the half extents (hx, hy) of the
// alpha >= 1/255 box as two bfloat16 values rounded UP in magnitude (the box only grows, so the bbox cull stays
// conservative; +/-inf survive), `id` the Gaussian index as raw bits.  One LDS.128 per surviving record then
// carries everything but conic / opacity / features, and the backward needs no gaussian_ids_sorted lookup.
__device__ __forceinline__ float4 gfb_pack_record_a(float u, float v, float hx, float hy, int id) {
    const unsigned int bx = (__float_as_uint(hx) + 0xffffu) & 0xffff0000u;
    const unsigned int by = (__float_as_uint(hy) + 0xffffu) & 0xffff0000u;
    return make_float4(u, v, __uint_as_float(bx | (by >> 16)), __int_as_float(id));
}
__device__ __forceinline__ void gfb_unpack_extents(float packed, float& hx, float& hy) {
    const unsigned int w = __float_as_uint(packed);
    hx = __uint_as_float(w & 0xffff0000u);
    hy = __uint_as_float(w << 16);
}

__device__ __forceinline__ float gfb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
