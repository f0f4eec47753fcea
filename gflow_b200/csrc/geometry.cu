// geometry.cu -- per-Gaussian geometry ops of the msplat surface for sm_100a:
//   project_point, compute_cov3d, ewa_project, compute_sh (forward + backward).
//
// One thread per Gaussian; these are streaming kernels bounded by HBM/L2 latency
// (24..93 B per Gaussian).  This translation unit is compiled with
// -fmad=false so every multiply / add rounds once, in the operation order documented
// in oracle/splat_ref.py: the float32 forward results (depth bits, radius,
// tiles_touched) are then bit-identical to the CPU oracle, which is what makes
// gaussian_ids_sorted / tile_range bit-exact downstream.
//
// Reference interface: msplat.project_point / compute_cov3d / ewa_project /
// compute_sh as called from /root/reference/gflow/utils/render.py:21-49,116-135 and
// /root/reference/gflow/trainer.py:955.
#include "splat_math.cuh"

namespace {

using namespace gfbm;

// ------------------------------------------------------------------ project_point
__global__ void __launch_bounds__(kThreads)
project_point_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ intr,
                         const float* __restrict__ extr, int N, int W, int H, float nearest, float extent,
                         float2* __restrict__ uv, float* __restrict__ depth) {
    __shared__ float s_cam[16];
    load_camera(s_cam, intr, extr);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float u, v, xc, yc, zc;
    const bool ok = project_one(s_cam + 12, s_cam, W, H, nearest, extent, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2],
                                u, v, xc, yc, zc);
    uv[i] = ok ? make_float2(u, v) : make_float2(0.0f, 0.0f);
    depth[i] = ok ? zc : 0.0f;
}

__global__ void __launch_bounds__(kThreads)
project_point_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ intr,
                         const float* __restrict__ extr, int N, int W, int H, float nearest, float extent,
                         const float2* __restrict__ g_uv, const float* __restrict__ g_depth,
                         float* __restrict__ d_xyz, float* __restrict__ d_cam /* 12 extr + 4 intr */) {
    __shared__ float s_cam[16];
    load_camera(s_cam, intr, extr);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    if (i < N) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        float u, v, xc, yc, zc;
        float dp[3] = {0.0f, 0.0f, 0.0f};
        if (project_one(s_cam + 12, s_cam, W, H, nearest, extent, x, y, z, u, v, xc, yc, zc)) {
            const float2 g = g_uv[i];
            project_bwd_one(s_cam + 12, s_cam, x, y, z, xc, yc, zc, g.x, g.y, g_depth ? g_depth[i] : 0.0f, dp, acc);
        }
        d_xyz[3 * i] = dp[0];
        d_xyz[3 * i + 1] = dp[1];
        d_xyz[3 * i + 2] = dp[2];
    }
    block_reduce_atomic<16>(acc, d_cam);
}

// ------------------------------------------------------------------ compute_cov3d
__global__ void __launch_bounds__(kThreads)
compute_cov3d_fwd_kernel(const float* __restrict__ scale, const float4* __restrict__ rotate,
                         const uint8_t* __restrict__ visible, int N, float* __restrict__ cov3d) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float o[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (!visible || visible[i]) {
        const float s[3] = {scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]};
        cov3d_fwd_one(s, rotate[i], o);
    }
    float2* dst = reinterpret_cast<float2*>(cov3d + 6 * (size_t)i);  // 24 B records are 8 B aligned
    dst[0] = make_float2(o[0], o[1]);
    dst[1] = make_float2(o[2], o[3]);
    dst[2] = make_float2(o[4], o[5]);
}

__global__ void __launch_bounds__(kThreads)
compute_cov3d_bwd_kernel(const float* __restrict__ scale, const float4* __restrict__ rotate,
                         const uint8_t* __restrict__ visible, int N, const float* __restrict__ g_cov3d,
                         float* __restrict__ d_scale, float4* __restrict__ d_rotate) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float ds[3] = {0.0f, 0.0f, 0.0f};
    float4 dq = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (!visible || visible[i]) {
        const float s[3] = {scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]};
        float g[6];
        load_cov3d(g_cov3d, i, g);
        cov3d_bwd_one(s, rotate[i], g, ds, dq);
    }
    d_scale[3 * i] = ds[0];
    d_scale[3 * i + 1] = ds[1];
    d_scale[3 * i + 2] = ds[2];
    d_rotate[i] = dq;
}

// ------------------------------------------------------------------ ewa_project
__global__ void __launch_bounds__(kThreads)
ewa_project_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ cov3d,
                       const float* __restrict__ intr, const float* __restrict__ extr,
                       const float2* __restrict__ uv, int N, int W, int H, const uint8_t* __restrict__ visible,
                       float* __restrict__ conic, int32_t* __restrict__ radius, int32_t* __restrict__ tiles_touched) {
    __shared__ float s_cam[16];
    load_camera(s_cam, intr, extr);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float ca = 0.0f, cb = 0.0f, cc = 0.0f;
    int rad = 0, tiles = 0;
    if (!visible || visible[i]) {
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        float S[6];
        load_cov3d(cov3d, i, S);
        EwaMid m;
        ewa_mid_eval(p, S, s_cam + 12, s_cam, W, H, m);
        float rf;
        int x0, y0, x1, y1;
        const float2 q = uv[i];
        if (ewa_live(m, q.x, q.y, (W + GFB_TILE - 1) / GFB_TILE, (H + GFB_TILE - 1) / GFB_TILE, rf, x0, y0, x1, y1)) {
            const float dinv = 1.0f / m.det;
            ca = m.c * dinv;
            cb = -m.b * dinv;
            cc = m.a * dinv;
            rad = (int)rf;
            tiles = (x1 - x0) * (y1 - y0);
        }
    }
    conic[3 * i] = ca;
    conic[3 * i + 1] = cb;
    conic[3 * i + 2] = cc;
    radius[i] = rad;
    tiles_touched[i] = tiles;
}

__global__ void __launch_bounds__(kThreads)
ewa_project_bwd_kernel(const float* __restrict__ xyz, const float* __restrict__ cov3d,
                       const float* __restrict__ intr, const float* __restrict__ extr,
                       const float2* __restrict__ uv, int N, int W, int H, const uint8_t* __restrict__ visible,
                       const float* __restrict__ g_conic, float* __restrict__ d_xyz, float* __restrict__ d_cov3d,
                       float* __restrict__ d_cam /* 12 extr + 4 intr (only [12],[13] used) */) {
    __shared__ float s_cam[16];
    load_camera(s_cam, intr, extr);
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float acc[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = 0.0f;
    if (i < N) {
        float dp[3] = {0.0f, 0.0f, 0.0f};
        float dS[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        if (!visible || visible[i]) {
            const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
            float S[6];
            load_cov3d(cov3d, i, S);
            EwaMid m;
            ewa_mid_eval(p, S, s_cam + 12, s_cam, W, H, m);
            float rf;
            int x0, y0, x1, y1;
            const float2 q = uv[i];
            if (ewa_live(m, q.x, q.y, (W + GFB_TILE - 1) / GFB_TILE, (H + GFB_TILE - 1) / GFB_TILE, rf, x0, y0, x1,
                         y1))
                ewa_bwd_one(m, p, S, s_cam + 12, s_cam, g_conic[3 * i], g_conic[3 * i + 1], g_conic[3 * i + 2], dp,
                            dS, acc);
        }
        d_xyz[3 * i] = dp[0];
        d_xyz[3 * i + 1] = dp[1];
        d_xyz[3 * i + 2] = dp[2];
        float2* dst = reinterpret_cast<float2*>(d_cov3d + 6 * (size_t)i);
        dst[0] = make_float2(dS[0], dS[1]);
        dst[1] = make_float2(dS[2], dS[3]);
        dst[2] = make_float2(dS[4], dS[5]);
    }
    block_reduce_atomic<14>(acc, d_cam);
}

// ------------------------------------------------------------------ compute_sh
__constant__ float kShC2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
__constant__ float kShC3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};
constexpr float kShC0 = 0.28209479177387814f;
constexpr float kShC1 = 0.4886025119029199f;

template <bool kDeriv>
__device__ __forceinline__ void sh_eval(float x, float y, float z, int K, float* Y, float* Yx, float* Yy, float* Yz) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        Y[k] = 0.0f;
        if (kDeriv) Yx[k] = Yy[k] = Yz[k] = 0.0f;
    }
    Y[0] = kShC0;
    if (K > 1) {
        Y[1] = -kShC1 * y;
        Y[2] = kShC1 * z;
        Y[3] = -kShC1 * x;
        if (kDeriv) { Yy[1] = -kShC1; Yz[2] = kShC1; Yx[3] = -kShC1; }
    }
    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    if (K > 4) {
        Y[4] = kShC2[0] * xy;
        Y[5] = kShC2[1] * yz;
        Y[6] = kShC2[2] * (2.0f * zz - xx - yy);
        Y[7] = kShC2[3] * xz;
        Y[8] = kShC2[4] * (xx - yy);
        if (kDeriv) {
            Yx[4] = kShC2[0] * y; Yy[4] = kShC2[0] * x;
            Yy[5] = kShC2[1] * z; Yz[5] = kShC2[1] * y;
            Yx[6] = kShC2[2] * -2.0f * x; Yy[6] = kShC2[2] * -2.0f * y; Yz[6] = kShC2[2] * 4.0f * z;
            Yx[7] = kShC2[3] * z; Yz[7] = kShC2[3] * x;
            Yx[8] = kShC2[4] * 2.0f * x; Yy[8] = kShC2[4] * -2.0f * y;
        }
    }
    if (K > 9) {
        Y[9] = kShC3[0] * y * (3.0f * xx - yy);
        Y[10] = kShC3[1] * xy * z;
        Y[11] = kShC3[2] * y * (4.0f * zz - xx - yy);
        Y[12] = kShC3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        Y[13] = kShC3[4] * x * (4.0f * zz - xx - yy);
        Y[14] = kShC3[5] * z * (xx - yy);
        Y[15] = kShC3[6] * x * (xx - 3.0f * yy);
        if (kDeriv) {
            Yx[9] = kShC3[0] * 6.0f * xy; Yy[9] = kShC3[0] * (3.0f * xx - 3.0f * yy);
            Yx[10] = kShC3[1] * yz; Yy[10] = kShC3[1] * xz; Yz[10] = kShC3[1] * xy;
            Yx[11] = kShC3[2] * -2.0f * xy; Yy[11] = kShC3[2] * (4.0f * zz - xx - 3.0f * yy);
            Yz[11] = kShC3[2] * 8.0f * yz;
            Yx[12] = kShC3[3] * -6.0f * xz; Yy[12] = kShC3[3] * -6.0f * yz;
            Yz[12] = kShC3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
            Yx[13] = kShC3[4] * (4.0f * zz - 3.0f * xx - yy); Yy[13] = kShC3[4] * -2.0f * xy;
            Yz[13] = kShC3[4] * 8.0f * xz;
            Yx[14] = kShC3[5] * 2.0f * xz; Yy[14] = kShC3[5] * -2.0f * yz; Yz[14] = kShC3[5] * (xx - yy);
            Yx[15] = kShC3[6] * (3.0f * xx - 3.0f * yy); Yy[15] = kShC3[6] * -6.0f * xy;
        }
    }
}

// One warp per group of 32 Gaussians would leave the 4*C*K-byte SH rows strided; instead
// each thread owns one Gaussian and walks its contiguous C*K coefficients (192 B at
// degree 3), which the L1 turns into full-line requests across the warp.
__global__ void __launch_bounds__(kThreads)
compute_sh_fwd_kernel(const float* __restrict__ shs, const float* __restrict__ dirs,
                      const uint8_t* __restrict__ visible, int N, int C, int K, float* __restrict__ out) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float* o = out + (size_t)i * C;
    if (visible && !visible[i]) {
        for (int c = 0; c < C; ++c) o[c] = 0.0f;
        return;
    }
    const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
    float Y[16];
    sh_eval<false>(x * inv, y * inv, z * inv, K, Y, nullptr, nullptr, nullptr);
    for (int c = 0; c < C; ++c) {
        const float* s = shs + ((size_t)i * C + c) * K;
        float acc = 0.0f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) acc += s[k] * Y[k];
        o[c] = acc;
    }
}

__global__ void __launch_bounds__(kThreads)
compute_sh_bwd_kernel(const float* __restrict__ shs, const float* __restrict__ dirs,
                      const uint8_t* __restrict__ visible, int N, int C, int K, const float* __restrict__ g_out,
                      float* __restrict__ d_shs, float* __restrict__ d_dirs) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float dd[3] = {0.0f, 0.0f, 0.0f};
    if (visible && !visible[i]) {
        for (int k = 0; k < C * K; ++k) d_shs[(size_t)i * C * K + k] = 0.0f;
    } else {
        const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
        const float nx = x * inv, ny = y * inv, nz = z * inv;
        float Y[16], Yx[16], Yy[16], Yz[16];
        sh_eval<true>(nx, ny, nz, K, Y, Yx, Yy, Yz);
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float* s = shs + ((size_t)i * C + c) * K;
            float* ds = d_shs + ((size_t)i * C + c) * K;
            const float g = g_out[(size_t)i * C + c];
#pragma unroll 4
            for (int k = 0; k < K; ++k) {
                ds[k] = g * Y[k];
                const float gs = g * s[k];
                gx += gs * Yx[k];
                gy += gs * Yy[k];
                gz += gs * Yz[k];
            }
        }
        const float dot = gx * nx + gy * ny + gz * nz;
        dd[0] = (gx - nx * dot) * inv;
        dd[1] = (gy - ny * dot) * inv;
        dd[2] = (gz - nz * dot) * inv;
    }
    d_dirs[3 * i] = dd[0];
    d_dirs[3 * i + 1] = dd[1];
    d_dirs[3 * i + 2] = dd[2];
}

// ---- coalesced variants for C == 3 (rgb): the (N, 3, K) coefficient block of a warp's 32 Gaussians is contiguous in
// memory, so the warp streams it through shared memory with 128-bit loads / stores and every lane then works on its
// own row there (odd pitch: conflict-free).  The per-thread walk above touches 32 different lines per load
// instruction; at 200 000 Gaussians / degree 3 (38 MB in, 38 MB of gradients out) this kernel is the one purely
// HBM-bound piece of the path, so the access pattern is the whole story.  Arithmetic and its order are
// those of the kernels above.
constexpr int kShWarps = 4;
// row pitch of the tile in floats: for CK % 4 == 0 an ODD number of float4s, so that rows read / written with 128-bit
// accesses by consecutive lanes fall into disjoint bank groups; otherwise an odd number of floats (scalar accesses)
template <int CK>
struct ShPitch {
    static constexpr int value = (CK % 4 == 0) ? 4 * ((CK / 4) | 1) : (CK | 1);
};

template <int CK>
__device__ __forceinline__ void sh_tile_load(const float* __restrict__ src, int total, float* tile, int lane) {
    constexpr int PITCH = ShPitch<CK>::value;
    if constexpr (CK % 4 == 0) {
        // all CK / 4 loads of the lane are issued before the first one is consumed: one round trip to HBM per warp
        // tile instead of CK / 4 dependent ones
        const float4* src4 = reinterpret_cast<const float4*>(src);
        float4 v[CK / 4];
#pragma unroll
        for (int j = 0; j < CK / 4; ++j) {
            const int q = lane + 32 * j;
            v[j] = (q < total / 4) ? src4[q] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int j = 0; j < CK / 4; ++j) {
            const int q = lane + 32 * j;
            const int r = (4 * q) / CK, c = (4 * q) % CK;  // CK % 4 == 0: the four values stay in one row
            *reinterpret_cast<float4*>(tile + r * PITCH + c) = v[j];
        }
    } else {
        float v[CK];
#pragma unroll
        for (int j = 0; j < CK; ++j) {
            const int idx = lane + 32 * j;
            v[j] = (idx < total) ? src[idx] : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < CK; ++j) {
            const int idx = lane + 32 * j;
            tile[(idx / CK) * PITCH + idx % CK] = v[j];
        }
    }
    __syncwarp();
}

template <int CK>
__device__ __forceinline__ void sh_tile_store(float* __restrict__ dst, int total, const float* tile, int lane) {
    constexpr int PITCH = ShPitch<CK>::value;
    __syncwarp();
    if constexpr (CK % 4 == 0) {
        float4* dst4 = reinterpret_cast<float4*>(dst);
#pragma unroll
        for (int j = 0; j < CK / 4; ++j) {
            const int q = lane + 32 * j;
            const int r = (4 * q) / CK, c = (4 * q) % CK;
            if (q < total / 4) dst4[q] = *reinterpret_cast<const float4*>(tile + r * PITCH + c);
        }
    } else {
#pragma unroll
        for (int j = 0; j < CK; ++j) {
            const int idx = lane + 32 * j;
            if (idx < total) dst[idx] = tile[(idx / CK) * PITCH + idx % CK];
        }
    }
}

// the lane's row of the tile into registers / back (128-bit accesses when the pitch allows)
template <int CK>
__device__ __forceinline__ void sh_row_read(const float* src, float (&row)[CK]) {
    if constexpr (CK % 4 == 0) {
#pragma unroll
        for (int j = 0; j < CK / 4; ++j) {
            const float4 v = reinterpret_cast<const float4*>(src)[j];
            row[4 * j] = v.x, row[4 * j + 1] = v.y, row[4 * j + 2] = v.z, row[4 * j + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < CK; ++j) row[j] = src[j];
    }
}
template <int CK>
__device__ __forceinline__ void sh_row_write(float* dst, const float (&row)[CK]) {
    if constexpr (CK % 4 == 0) {
#pragma unroll
        for (int j = 0; j < CK / 4; ++j)
            reinterpret_cast<float4*>(dst)[j] = make_float4(row[4 * j], row[4 * j + 1], row[4 * j + 2], row[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < CK; ++j) dst[j] = row[j];
    }
}

template <int K>
__global__ void __launch_bounds__(32 * kShWarps)
compute_sh3_fwd_kernel(const float* __restrict__ shs, const float* __restrict__ dirs, const uint8_t* __restrict__ visible,
                       int N, float* __restrict__ out) {
    constexpr int C = 3, CK = C * K, PITCH = ShPitch<CK>::value;
    __shared__ __align__(16) float s_tile[kShWarps][32 * PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (blockIdx.x * kShWarps + warp) * 32;
    if (base >= N) return;
    const int rows = min(32, N - base);
    float* tile = s_tile[warp];
    sh_tile_load<CK>(shs + (size_t)base * CK, rows * CK, tile, lane);
    const int i = base + lane;
    if (i >= N) return;
    float* o = out + (size_t)i * C;
    if (visible && !visible[i]) {
        o[0] = o[1] = o[2] = 0.0f;
        return;
    }
    const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
    const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
    float Y[16];
    sh_eval<false>(x * inv, y * inv, z * inv, K, Y, nullptr, nullptr, nullptr);
    float row[CK];
    sh_row_read<CK>(tile + lane * PITCH, row);
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) acc += row[c * K + k] * Y[k];
        o[c] = acc;
    }
}

template <int K>
__global__ void __launch_bounds__(32 * kShWarps)
compute_sh3_bwd_kernel(const float* __restrict__ shs, const float* __restrict__ dirs, const uint8_t* __restrict__ visible,
                       int N, const float* __restrict__ g_out, float* __restrict__ d_shs, float* __restrict__ d_dirs) {
    constexpr int C = 3, CK = C * K, PITCH = ShPitch<CK>::value;
    __shared__ __align__(16) float s_tile[kShWarps][32 * PITCH];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int base = (blockIdx.x * kShWarps + warp) * 32;
    if (base >= N) return;
    const int rows = min(32, N - base);
    float* tile = s_tile[warp];
    sh_tile_load<CK>(shs + (size_t)base * CK, rows * CK, tile, lane);
    const int i = base + lane;
    if (i < N) {
        float dd[3] = {0.0f, 0.0f, 0.0f};
        float row[CK];  // coefficients in, their gradients out (each lane owns its row of the tile)
        if (visible && !visible[i]) {
#pragma unroll
            for (int k = 0; k < CK; ++k) row[k] = 0.0f;
        } else {
            sh_row_read<CK>(tile + lane * PITCH, row);
            const float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
            const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
            const float nx = x * inv, ny = y * inv, nz = z * inv;
            float Y[16], Yx[16], Yy[16], Yz[16];
            sh_eval<true>(nx, ny, nz, K, Y, Yx, Yy, Yz);
            float gx = 0.0f, gy = 0.0f, gz = 0.0f;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float g = g_out[(size_t)i * C + c];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const float gs = g * row[c * K + k];
                    row[c * K + k] = g * Y[k];
                    gx += gs * Yx[k];
                    gy += gs * Yy[k];
                    gz += gs * Yz[k];
                }
            }
            const float dot = gx * nx + gy * ny + gz * nz;
            dd[0] = (gx - nx * dot) * inv;
            dd[1] = (gy - ny * dot) * inv;
            dd[2] = (gz - nz * dot) * inv;
        }
        sh_row_write<CK>(tile + lane * PITCH, row);
        d_dirs[3 * i] = dd[0];
        d_dirs[3 * i + 1] = dd[1];
        d_dirs[3 * i + 2] = dd[2];
    }
    sh_tile_store<CK>(d_shs + (size_t)base * CK, rows * CK, tile, lane);
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

int gfb_project_point_fwd(const float* xyz, const float* intr, const float* extr, int N, int W, int H, float nearest,
                          float extent, float* uv, float* depth, void* stream) {
    if (N < 0 || W <= 0 || H <= 0) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!xyz || !intr || !extr || !uv || !depth) return GFB_E_BADARG;
    project_point_fwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        xyz, intr, extr, N, W, H, nearest, extent, reinterpret_cast<float2*>(uv), depth);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_project_point_bwd(const float* xyz, const float* intr, const float* extr, int N, int W, int H, float nearest,
                          float extent, const float* g_uv, const float* g_depth, float* d_xyz, float* d_cam,
                          void* stream) {
    if (N < 0 || W <= 0 || H <= 0 || !d_cam) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    GFB_TRY(cudaMemsetAsync(d_cam, 0, 16 * sizeof(float), st));
    if (N == 0) return 0;
    if (!xyz || !intr || !extr || !g_uv || !d_xyz) return GFB_E_BADARG;
    project_point_bwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(
        xyz, intr, extr, N, W, H, nearest, extent, reinterpret_cast<const float2*>(g_uv), g_depth, d_xyz, d_cam);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_compute_cov3d_fwd(const float* scale, const float* rotate, const uint8_t* visible, int N, float* cov3d,
                          void* stream) {
    if (N < 0) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!scale || !rotate || !cov3d) return GFB_E_BADARG;
    compute_cov3d_fwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        scale, reinterpret_cast<const float4*>(rotate), visible, N, cov3d);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_compute_cov3d_bwd(const float* scale, const float* rotate, const uint8_t* visible, int N,
                          const float* g_cov3d, float* d_scale, float* d_rotate, void* stream) {
    if (N < 0) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!scale || !rotate || !g_cov3d || !d_scale || !d_rotate) return GFB_E_BADARG;
    compute_cov3d_bwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        scale, reinterpret_cast<const float4*>(rotate), visible, N, g_cov3d, d_scale,
        reinterpret_cast<float4*>(d_rotate));
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_ewa_project_fwd(const float* xyz, const float* cov3d, const float* intr, const float* extr, const float* uv,
                        int N, int W, int H, const uint8_t* visible, float* conic, int32_t* radius,
                        int32_t* tiles_touched, void* stream) {
    if (N < 0 || W <= 0 || H <= 0) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!xyz || !cov3d || !intr || !extr || !uv || !conic || !radius || !tiles_touched) return GFB_E_BADARG;
    ewa_project_fwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(
        xyz, cov3d, intr, extr, reinterpret_cast<const float2*>(uv), N, W, H, visible, conic, radius, tiles_touched);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_ewa_project_bwd(const float* xyz, const float* cov3d, const float* intr, const float* extr, const float* uv,
                        int N, int W, int H, const uint8_t* visible, const float* g_conic, float* d_xyz,
                        float* d_cov3d, float* d_cam, void* stream) {
    if (N < 0 || W <= 0 || H <= 0 || !d_cam) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    GFB_TRY(cudaMemsetAsync(d_cam, 0, 16 * sizeof(float), st));
    if (N == 0) return 0;
    if (!xyz || !cov3d || !intr || !extr || !uv || !g_conic || !d_xyz || !d_cov3d) return GFB_E_BADARG;
    ewa_project_bwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(
        xyz, cov3d, intr, extr, reinterpret_cast<const float2*>(uv), N, W, H, visible, g_conic, d_xyz, d_cov3d,
        d_cam);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_compute_sh_fwd(const float* shs, const float* dirs, const uint8_t* visible, int N, int C, int K, float* out,
                       void* stream) {
    if (N < 0 || C <= 0 || !(K == 1 || K == 4 || K == 9 || K == 16)) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!shs || !dirs || !out) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3 && ((uintptr_t)shs & 15u) == 0) {  // coalesced path (float4 tile loads need a 16-byte aligned base)
        const int grid = gfb_div_up(N, 32 * kShWarps), block = 32 * kShWarps;
        switch (K) {
            case 1: compute_sh3_fwd_kernel<1><<<grid, block, 0, st>>>(shs, dirs, visible, N, out); break;
            case 4: compute_sh3_fwd_kernel<4><<<grid, block, 0, st>>>(shs, dirs, visible, N, out); break;
            case 9: compute_sh3_fwd_kernel<9><<<grid, block, 0, st>>>(shs, dirs, visible, N, out); break;
            default: compute_sh3_fwd_kernel<16><<<grid, block, 0, st>>>(shs, dirs, visible, N, out); break;
        }
    } else {
        compute_sh_fwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(shs, dirs, visible, N, C, K, out);
    }
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_compute_sh_bwd(const float* shs, const float* dirs, const uint8_t* visible, int N, int C, int K,
                       const float* g_out, float* d_shs, float* d_dirs, void* stream) {
    if (N < 0 || C <= 0 || !(K == 1 || K == 4 || K == 9 || K == 16)) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!shs || !dirs || !g_out || !d_shs || !d_dirs) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3 && (((uintptr_t)shs | (uintptr_t)d_shs) & 15u) == 0) {
        const int grid = gfb_div_up(N, 32 * kShWarps), block = 32 * kShWarps;
        switch (K) {
            case 1: compute_sh3_bwd_kernel<1><<<grid, block, 0, st>>>(shs, dirs, visible, N, g_out, d_shs, d_dirs); break;
            case 4: compute_sh3_bwd_kernel<4><<<grid, block, 0, st>>>(shs, dirs, visible, N, g_out, d_shs, d_dirs); break;
            case 9: compute_sh3_bwd_kernel<9><<<grid, block, 0, st>>>(shs, dirs, visible, N, g_out, d_shs, d_dirs); break;
            default: compute_sh3_bwd_kernel<16><<<grid, block, 0, st>>>(shs, dirs, visible, N, g_out, d_shs, d_dirs); break;
        }
    } else {
        compute_sh_bwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, st>>>(shs, dirs, visible, N, C, K, g_out, d_shs, d_dirs);
    }
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
