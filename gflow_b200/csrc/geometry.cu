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
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// Block-wide sum of NV per-thread values, then one atomicAdd per value per block.
template <int NV>
__device__ __forceinline__ void block_reduce_atomic(float (&v)[NV], float* __restrict__ dst) {
    __shared__ float s_part[kThreads / 32][NV];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = gfb_warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) s_part[warp][k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float acc = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) acc += s_part[w][threadIdx.x];
        if (acc != 0.0f) atomicAdd(dst + threadIdx.x, acc);
    }
}

// ------------------------------------------------------------------ project_point
__device__ __forceinline__ bool project_one(const float* __restrict__ intr, const float* __restrict__ extr, int W,
                                            int H, float nearest, float extent, float x, float y, float z, float& u,
                                            float& v, float& xc, float& yc, float& zc) {
    gfb_cam_point(extr, x, y, z, xc, yc, zc);
    if (!(zc > nearest)) return false;
    u = (intr[0] * xc) / zc + intr[2];
    v = (intr[1] * yc) / zc + intr[3];
    const float xn = u / (0.5f * (float)W) - 1.0f;
    const float yn = v / (0.5f * (float)H) - 1.0f;
    return (fabsf(xn) <= extent) && (fabsf(yn) <= extent);
}

__global__ void __launch_bounds__(kThreads)
project_point_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ intr,
                         const float* __restrict__ extr, int N, int W, int H, float nearest, float extent,
                         float2* __restrict__ uv, float* __restrict__ depth) {
    __shared__ float s_cam[16];
    if (threadIdx.x < 12) s_cam[threadIdx.x] = extr[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) s_cam[threadIdx.x] = intr[threadIdx.x - 12];
    __syncthreads();
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
    if (threadIdx.x < 12) s_cam[threadIdx.x] = extr[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) s_cam[threadIdx.x] = intr[threadIdx.x - 12];
    __syncthreads();
    const float* e = s_cam;
    const float* in = s_cam + 12;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    if (i < N) {
        const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        float u, v, xc, yc, zc;
        const bool ok = project_one(in, e, W, H, nearest, extent, x, y, z, u, v, xc, yc, zc);
        float dx = 0.0f, dy = 0.0f, dz = 0.0f;
        if (ok) {
            const float2 g = g_uv[i];
            const float gd = g_depth ? g_depth[i] : 0.0f;
            const float iz = 1.0f / zc;
            const float gx = in[0] * iz * g.x;
            const float gy = in[1] * iz * g.y;
            const float gz = -(in[0] * xc * iz * iz) * g.x - (in[1] * yc * iz * iz) * g.y + gd;
            dx = e[0] * gx + e[4] * gy + e[8] * gz;
            dy = e[1] * gx + e[5] * gy + e[9] * gz;
            dz = e[2] * gx + e[6] * gy + e[10] * gz;
            acc[0] = gx * x; acc[1] = gx * y; acc[2] = gx * z; acc[3] = gx;
            acc[4] = gy * x; acc[5] = gy * y; acc[6] = gy * z; acc[7] = gy;
            acc[8] = gz * x; acc[9] = gz * y; acc[10] = gz * z; acc[11] = gz;
            acc[12] = g.x * xc * iz;
            acc[13] = g.y * yc * iz;
            acc[14] = g.x;
            acc[15] = g.y;
        }
        d_xyz[3 * i] = dx;
        d_xyz[3 * i + 1] = dy;
        d_xyz[3 * i + 2] = dz;
    }
    block_reduce_atomic<16>(acc, d_cam);
}

// ------------------------------------------------------------------ compute_cov3d
__device__ __forceinline__ void quat_rot(float w, float x, float y, float z, float* R) {
    R[0] = 1.0f - 2.0f * (y * y + z * z);
    R[1] = 2.0f * (x * y - w * z);
    R[2] = 2.0f * (x * z + w * y);
    R[3] = 2.0f * (x * y + w * z);
    R[4] = 1.0f - 2.0f * (x * x + z * z);
    R[5] = 2.0f * (y * z - w * x);
    R[6] = 2.0f * (x * z - w * y);
    R[7] = 2.0f * (y * z + w * x);
    R[8] = 1.0f - 2.0f * (x * x + y * y);
}

__global__ void __launch_bounds__(kThreads)
compute_cov3d_fwd_kernel(const float* __restrict__ scale, const float4* __restrict__ rotate,
                         const uint8_t* __restrict__ visible, int N, float* __restrict__ cov3d) {
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= N) return;
    float o[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (!visible || visible[i]) {
        const float4 q = rotate[i];
        const float s[3] = {scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]};
        float R[9], M[9];
        quat_rot(q.x, q.y, q.z, q.w, R);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * s[c];
        o[0] = (M[0] * M[0] + M[1] * M[1]) + M[2] * M[2];
        o[1] = (M[0] * M[3] + M[1] * M[4]) + M[2] * M[5];
        o[2] = (M[0] * M[6] + M[1] * M[7]) + M[2] * M[8];
        o[3] = (M[3] * M[3] + M[4] * M[4]) + M[5] * M[5];
        o[4] = (M[3] * M[6] + M[4] * M[7]) + M[5] * M[8];
        o[5] = (M[6] * M[6] + M[7] * M[7]) + M[8] * M[8];
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
        const float4 q4 = rotate[i];
        const float w = q4.x, x = q4.y, y = q4.z, z = q4.w;
        const float s[3] = {scale[3 * i], scale[3 * i + 1], scale[3 * i + 2]};
        const float2* gp = reinterpret_cast<const float2*>(g_cov3d + 6 * (size_t)i);
        const float2 g01 = gp[0], g23 = gp[1], g45 = gp[2];
        const float g[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
        float R[9], M[9], Gs[9], dM[9], D[9];
        quat_rot(w, x, y, z, R);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * s[c];
        Gs[0] = 2.0f * g[0]; Gs[1] = g[1]; Gs[2] = g[2];
        Gs[3] = g[1]; Gs[4] = 2.0f * g[3]; Gs[5] = g[4];
        Gs[6] = g[2]; Gs[7] = g[4]; Gs[8] = 2.0f * g[5];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                dM[3 * r + c] = Gs[3 * r] * M[c] + Gs[3 * r + 1] * M[3 + c] + Gs[3 * r + 2] * M[6 + c];
#pragma unroll
        for (int c = 0; c < 3; ++c) ds[c] = dM[c] * R[c] + dM[3 + c] * R[3 + c] + dM[6 + c] * R[6 + c];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) D[3 * r + c] = dM[3 * r + c] * s[c];
        dq.x = 2.0f * (-z * D[1] + y * D[2] + z * D[3] - x * D[5] - y * D[6] + x * D[7]);
        dq.y = 2.0f * (y * D[1] + z * D[2] + y * D[3] - 2.0f * x * D[4] - w * D[5] + z * D[6] + w * D[7] -
                       2.0f * x * D[8]);
        dq.z = 2.0f * (-2.0f * y * D[0] + x * D[1] + w * D[2] + x * D[3] + z * D[5] - w * D[6] + z * D[7] -
                       2.0f * y * D[8]);
        dq.w = 2.0f * (-2.0f * z * D[0] - w * D[1] + x * D[2] + w * D[3] - 2.0f * z * D[4] + y * D[5] + x * D[6] +
                       y * D[7]);
    }
    d_scale[3 * i] = ds[0];
    d_scale[3 * i + 1] = ds[1];
    d_scale[3 * i + 2] = ds[2];
    d_rotate[i] = dq;
}

// ------------------------------------------------------------------ ewa_project
struct EwaMid {
    float tx, ty, tz, txc, tyc, j00, j02, j11, j12, T0[3], T1[3], a, b, c, det;
    bool clampx, clampy;
};

__device__ __forceinline__ void ewa_mid_eval(const float* p, const float* S, const float* __restrict__ intr,
                                             const float* __restrict__ e, int W, int H, EwaMid& m) {
    const float fx = intr[0], fy = intr[1];
    gfb_cam_point(e, p[0], p[1], p[2], m.tx, m.ty, m.tz);
    const float limx = GFB_FRUSTUM_CLAMP * ((float)W / (2.0f * fx));
    const float limy = GFB_FRUSTUM_CLAMP * ((float)H / (2.0f * fy));
    const float rx = m.tx / m.tz, ry = m.ty / m.tz;
    m.clampx = (rx < -limx) || (rx > limx);
    m.clampy = (ry < -limy) || (ry > limy);
    m.txc = fminf(limx, fmaxf(-limx, rx)) * m.tz;
    m.tyc = fminf(limy, fmaxf(-limy, ry)) * m.tz;
    m.j00 = fx / m.tz;
    m.j02 = -(fx * m.txc) / (m.tz * m.tz);
    m.j11 = fy / m.tz;
    m.j12 = -(fy * m.tyc) / (m.tz * m.tz);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        m.T0[k] = m.j00 * e[k] + m.j02 * e[8 + k];
        m.T1[k] = m.j11 * e[4 + k] + m.j12 * e[8 + k];
    }
    const float* T0 = m.T0;
    const float* T1 = m.T1;
    const float a0 = (S[0] * T0[0] + S[1] * T0[1]) + S[2] * T0[2];
    const float a1 = (S[1] * T0[0] + S[3] * T0[1]) + S[4] * T0[2];
    const float a2 = (S[2] * T0[0] + S[4] * T0[1]) + S[5] * T0[2];
    const float b0 = (S[0] * T1[0] + S[1] * T1[1]) + S[2] * T1[2];
    const float b1 = (S[1] * T1[0] + S[3] * T1[1]) + S[4] * T1[2];
    const float b2 = (S[2] * T1[0] + S[4] * T1[1]) + S[5] * T1[2];
    m.a = ((T0[0] * a0 + T0[1] * a1) + T0[2] * a2) + GFB_COV_BLUR;
    m.b = (T1[0] * a0 + T1[1] * a1) + T1[2] * a2;
    m.c = ((T1[0] * b0 + T1[1] * b1) + T1[2] * b2) + GFB_COV_BLUR;
    m.det = m.a * m.c - m.b * m.b;
}

__device__ __forceinline__ void load_cov3d(const float* __restrict__ cov3d, int i, float* S) {
    const float2* sp = reinterpret_cast<const float2*>(cov3d + 6 * (size_t)i);
    const float2 s01 = sp[0], s23 = sp[1], s45 = sp[2];
    S[0] = s01.x; S[1] = s01.y; S[2] = s23.x; S[3] = s23.y; S[4] = s45.x; S[5] = s45.y;
}

// Returns true when the Gaussian is live (visible, det != 0, touches >= 1 tile).
__device__ __forceinline__ bool ewa_live(const EwaMid& m, float2 uv, int gx, int gy, float& rf, int& area) {
    if (m.det == 0.0f) return false;
    const float mid = 0.5f * (m.a + m.c);
    const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - m.det));
    rf = ceilf(3.0f * sqrtf(lam));
    int x0, y0, x1, y1;
    gfb_tile_rect(uv.x, uv.y, rf, gx, gy, x0, y0, x1, y1);
    area = (x1 - x0) * (y1 - y0);
    return area > 0;
}

__global__ void __launch_bounds__(kThreads)
ewa_project_fwd_kernel(const float* __restrict__ xyz, const float* __restrict__ cov3d,
                       const float* __restrict__ intr, const float* __restrict__ extr,
                       const float2* __restrict__ uv, int N, int W, int H, const uint8_t* __restrict__ visible,
                       float* __restrict__ conic, int32_t* __restrict__ radius, int32_t* __restrict__ tiles_touched) {
    __shared__ float s_cam[16];
    if (threadIdx.x < 12) s_cam[threadIdx.x] = extr[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) s_cam[threadIdx.x] = intr[threadIdx.x - 12];
    __syncthreads();
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
        int area;
        if (ewa_live(m, uv[i], (W + GFB_TILE - 1) / GFB_TILE, (H + GFB_TILE - 1) / GFB_TILE, rf, area)) {
            const float dinv = 1.0f / m.det;
            ca = m.c * dinv;
            cb = -m.b * dinv;
            cc = m.a * dinv;
            rad = (int)rf;
            tiles = area;
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
    if (threadIdx.x < 12) s_cam[threadIdx.x] = extr[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) s_cam[threadIdx.x] = intr[threadIdx.x - 12];
    __syncthreads();
    const float* e = s_cam;
    const float* in = s_cam + 12;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    float acc[14];
#pragma unroll
    for (int k = 0; k < 14; ++k) acc[k] = 0.0f;
    if (i < N) {
        float dp[3] = {0.0f, 0.0f, 0.0f};
        float dS[6] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
        bool live = (!visible || visible[i]);
        EwaMid m;
        float p[3], S[6];
        if (live) {
            p[0] = xyz[3 * i]; p[1] = xyz[3 * i + 1]; p[2] = xyz[3 * i + 2];
            load_cov3d(cov3d, i, S);
            ewa_mid_eval(p, S, in, e, W, H, m);
            float rf;
            int area;
            live = ewa_live(m, uv[i], (W + GFB_TILE - 1) / GFB_TILE, (H + GFB_TILE - 1) / GFB_TILE, rf, area);
        }
        if (live) {
            const float gA = g_conic[3 * i], gB = g_conic[3 * i + 1], gC = g_conic[3 * i + 2];
            const float a = m.a, b = m.b, c = m.c, dinv = 1.0f / m.det, d2 = dinv * dinv;
            const float ga = d2 * (-c * c * gA + b * c * gB) + gC * (dinv - a * c * d2);
            const float gb = 2.0f * b * c * d2 * gA + gB * (-dinv - 2.0f * b * b * d2) + 2.0f * a * b * d2 * gC;
            const float gc = gA * (dinv - a * c * d2) + a * b * d2 * gB - a * a * d2 * gC;
            const float* T0 = m.T0;
            const float* T1 = m.T1;
            dS[0] = ga * T0[0] * T0[0] + gb * T0[0] * T1[0] + gc * T1[0] * T1[0];
            dS[3] = ga * T0[1] * T0[1] + gb * T0[1] * T1[1] + gc * T1[1] * T1[1];
            dS[5] = ga * T0[2] * T0[2] + gb * T0[2] * T1[2] + gc * T1[2] * T1[2];
            dS[1] = 2.0f * ga * T0[0] * T0[1] + gb * (T0[0] * T1[1] + T0[1] * T1[0]) + 2.0f * gc * T1[0] * T1[1];
            dS[2] = 2.0f * ga * T0[0] * T0[2] + gb * (T0[0] * T1[2] + T0[2] * T1[0]) + 2.0f * gc * T1[0] * T1[2];
            dS[4] = 2.0f * ga * T0[1] * T0[2] + gb * (T0[1] * T1[2] + T0[2] * T1[1]) + 2.0f * gc * T1[1] * T1[2];
            float ST0[3], ST1[3], dT0[3], dT1[3];
            ST0[0] = S[0] * T0[0] + S[1] * T0[1] + S[2] * T0[2];
            ST0[1] = S[1] * T0[0] + S[3] * T0[1] + S[4] * T0[2];
            ST0[2] = S[2] * T0[0] + S[4] * T0[1] + S[5] * T0[2];
            ST1[0] = S[0] * T1[0] + S[1] * T1[1] + S[2] * T1[2];
            ST1[1] = S[1] * T1[0] + S[3] * T1[1] + S[4] * T1[2];
            ST1[2] = S[2] * T1[0] + S[4] * T1[1] + S[5] * T1[2];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dT0[k] = 2.0f * ga * ST0[k] + gb * ST1[k];
                dT1[k] = 2.0f * gc * ST1[k] + gb * ST0[k];
            }
            float dj00 = 0.0f, dj02 = 0.0f, dj11 = 0.0f, dj12 = 0.0f, dR[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dj00 += dT0[k] * e[k];
                dj02 += dT0[k] * e[8 + k];
                dj11 += dT1[k] * e[4 + k];
                dj12 += dT1[k] * e[8 + k];
                dR[k] = dT0[k] * m.j00;
                dR[3 + k] = dT1[k] * m.j11;
                dR[6 + k] = dT0[k] * m.j02 + dT1[k] * m.j12;
            }
            const float fx = in[0], fy = in[1], iz = 1.0f / m.tz, iz2 = iz * iz, iz3 = iz2 * iz;
            acc[12] = dj00 * iz - dj02 * m.txc * iz2;
            acc[13] = dj11 * iz - dj12 * m.tyc * iz2;
            const float dtxc = -dj02 * fx * iz2, dtyc = -dj12 * fy * iz2;
            float dtz = -dj00 * fx * iz2 + 2.0f * dj02 * fx * m.txc * iz3 - dj11 * fy * iz2 +
                        2.0f * dj12 * fy * m.tyc * iz3;
            float dtx = 0.0f, dty = 0.0f;
            if (m.clampx) dtz += dtxc * (m.txc * iz); else dtx = dtxc;
            if (m.clampy) dtz += dtyc * (m.tyc * iz); else dty = dtyc;
            const float dt[3] = {dtx, dty, dtz};
#pragma unroll
            for (int k = 0; k < 3; ++k) dp[k] = e[k] * dt[0] + e[4 + k] * dt[1] + e[8 + k] * dt[2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int k = 0; k < 3; ++k) acc[4 * r + k] = dR[3 * r + k] + dt[r] * p[k];
                acc[4 * r + 3] = dt[r];
            }
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
    compute_sh_fwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(shs, dirs, visible, N, C, K,
                                                                                           out);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_compute_sh_bwd(const float* shs, const float* dirs, const uint8_t* visible, int N, int C, int K,
                       const float* g_out, float* d_shs, float* d_dirs, void* stream) {
    if (N < 0 || C <= 0 || !(K == 1 || K == 4 || K == 9 || K == 16)) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!shs || !dirs || !g_out || !d_shs || !d_dirs) return GFB_E_BADARG;
    compute_sh_bwd_kernel<<<gfb_div_up(N, kThreads), kThreads, 0, (cudaStream_t)stream>>>(shs, dirs, visible, N, C, K,
                                                                                           g_out, d_shs, d_dirs);
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
