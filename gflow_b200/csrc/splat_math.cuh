// splat_math.cuh -- per-Gaussian device math shared by the op-level kernels (geometry.cu) and the
// fused render pipeline (pipeline.cu).  Both translation units are compiled with -fmad=false, and
// every expression keeps the operation order of oracle/splat_ref.py, so the float32 forward results
// are bit-identical between the two paths and against the CPU oracle.
#pragma once
#include "common.cuh"

namespace gfbm {

constexpr int kThreads = 256;

// Block-wide sum of NV per-thread values, then one atomicAdd per value per CTA.
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

// camera block in shared memory: [0..11] extr (3x4 row-major), [12..15] intr (fx fy cx cy)
__device__ __forceinline__ void load_camera(float* s_cam, const float* __restrict__ intr,
                                            const float* __restrict__ extr) {
    if (threadIdx.x < 12) s_cam[threadIdx.x] = extr[threadIdx.x];
    if (threadIdx.x >= 12 && threadIdx.x < 16) s_cam[threadIdx.x] = intr[threadIdx.x - 12];
    __syncthreads();
}

// ------------------------------------------------------------------ project_point
__device__ __forceinline__ bool project_one(const float* intr, const float* extr, int W, int H, float nearest,
                                            float extent, float x, float y, float z, float& u, float& v, float& xc,
                                            float& yc, float& zc) {
    gfb_cam_point(extr, x, y, z, xc, yc, zc);
    if (!(zc > nearest)) return false;
    u = (intr[0] * xc) / zc + intr[2];
    v = (intr[1] * yc) / zc + intr[3];
    const float xn = u / (0.5f * (float)W) - 1.0f;
    const float yn = v / (0.5f * (float)H) - 1.0f;
    return (fabsf(xn) <= extent) && (fabsf(yn) <= extent);
}

// Adds this point's contribution to d_xyz (dp) and to the 16 camera-gradient partials (acc).
__device__ __forceinline__ void project_bwd_one(const float* in, const float* e, float x, float y, float z, float xc,
                                                float yc, float zc, float gu, float gv, float gd, float* dp,
                                                float* acc) {
    const float iz = 1.0f / zc;
    const float gx = in[0] * iz * gu;
    const float gy = in[1] * iz * gv;
    const float gz = -(in[0] * xc * iz * iz) * gu - (in[1] * yc * iz * iz) * gv + gd;
    dp[0] += e[0] * gx + e[4] * gy + e[8] * gz;
    dp[1] += e[1] * gx + e[5] * gy + e[9] * gz;
    dp[2] += e[2] * gx + e[6] * gy + e[10] * gz;
    acc[0] += gx * x; acc[1] += gx * y; acc[2] += gx * z; acc[3] += gx;
    acc[4] += gy * x; acc[5] += gy * y; acc[6] += gy * z; acc[7] += gy;
    acc[8] += gz * x; acc[9] += gz * y; acc[10] += gz * z; acc[11] += gz;
    acc[12] += gu * xc * iz;
    acc[13] += gv * yc * iz;
    acc[14] += gu;
    acc[15] += gv;
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

__device__ __forceinline__ void cov3d_fwd_one(const float* s, float4 q, float* o) {
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

__device__ __forceinline__ void cov3d_bwd_one(const float* s, float4 q4, const float* g, float* ds, float4& dq) {
    const float w = q4.x, x = q4.y, y = q4.z, z = q4.w;
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

// ------------------------------------------------------------------ ewa_project
struct EwaMid {
    float tx, ty, tz, txc, tyc, j00, j02, j11, j12, T0[3], T1[3], a, b, c, det;
    bool clampx, clampy;
};

__device__ __forceinline__ void ewa_mid_eval(const float* p, const float* S, const float* intr, const float* e, int W,
                                             int H, EwaMid& m) {
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

// true when the Gaussian is live (det != 0 and its rect covers >= 1 tile); rect and radius returned
__device__ __forceinline__ bool ewa_live(const EwaMid& m, float u, float v, int gx, int gy, float& rf, int& x0,
                                         int& y0, int& x1, int& y1) {
    if (m.det == 0.0f) return false;
    const float mid = 0.5f * (m.a + m.c);
    const float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - m.det));
    rf = ceilf(3.0f * sqrtf(lam));
    gfb_tile_rect(u, v, rf, gx, gy, x0, y0, x1, y1);
    return (x1 - x0) * (y1 - y0) > 0;
}

// dL/dconic (gA,gB,gC) -> dL/dxyz (added to dp), dL/dcov3d (dS, overwritten), camera partials
// (added to acc: [0..11] extr, [12..13] fx fy).
__device__ __forceinline__ void ewa_bwd_one(const EwaMid& m, const float* p, const float* S, const float* in,
                                            const float* e, float gA, float gB, float gC, float* dp, float* dS,
                                            float* acc) {
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
    acc[12] += dj00 * iz - dj02 * m.txc * iz2;
    acc[13] += dj11 * iz - dj12 * m.tyc * iz2;
    const float dtxc = -dj02 * fx * iz2, dtyc = -dj12 * fy * iz2;
    float dtz = -dj00 * fx * iz2 + 2.0f * dj02 * fx * m.txc * iz3 - dj11 * fy * iz2 + 2.0f * dj12 * fy * m.tyc * iz3;
    float dtx = 0.0f, dty = 0.0f;
    if (m.clampx) dtz += dtxc * (m.txc * iz); else dtx = dtxc;
    if (m.clampy) dtz += dtyc * (m.tyc * iz); else dty = dtyc;
    const float dt[3] = {dtx, dty, dtz};
#pragma unroll
    for (int k = 0; k < 3; ++k) dp[k] += e[k] * dt[0] + e[4 + k] * dt[1] + e[8 + k] * dt[2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) acc[4 * r + k] += dR[3 * r + k] + dt[r] * p[k];
        acc[4 * r + 3] += dt[r];
    }
}

// ------------------------------------------------------------------ blend support
// Half extents of the bbox of {alpha = o exp(-q/2) >= 1/255}: an ellipse d^T Q d <= 2 ln(255 o).
// -inf = the Gaussian can never reach 1/255 (always rejected); +inf = conic not positive definite
// or NaN opacity (never culled).  The margins make the box conservative under rounding.
__device__ __forceinline__ void splat_bbox(float a, float b, float c, float o, float& hx, float& hy) {
    hx = -INFINITY;
    hy = -INFINITY;
    const float o255 = 255.0f * o;
    if (o255 >= 1.0f) {
        const float det = a * c - b * b;
        if (a > 0.0f && c > 0.0f && det > 0.0f) {
            const float s = 2.0f * logf(o255) / det;
            hx = sqrtf(s * c) * 1.0005f + 0.01f;
            hy = sqrtf(s * a) * 1.0005f + 0.01f;
        } else {
            hx = INFINITY;
            hy = INFINITY;
        }
    } else if (!(o255 < 1.0f)) {
        hx = INFINITY;
        hy = INFINITY;
    }
}

// Opt-in tile culling for the paths whose intersection list is internal (fused pipeline, native fit loop): shrink
// the 3-sigma tile rectangle [x0,x1) x [y0,y1) to the tiles the alpha >= 1/255 box of splat_bbox() reaches.  A pair
// dropped here has no pixel with alpha >= 1/255, so images and gradients are unchanged; the operator-level
// sort_gaussian keeps the reference's 3-sigma rule.  Returns false when no tile is left.
__device__ __forceinline__ bool tighten_rect(float u, float v, float a, float b, float c, float o, int& x0, int& y0,
                                             int& x1, int& y1) {
    float hx, hy;
    splat_bbox(a, b, c, o, hx, hy);
    if (hx < 0.0f) return false;             // -inf: the Gaussian never reaches 1/255
    if (hx < 3.0e38f && hy < 3.0e38f) {      // +inf: conic not positive definite -> keep the 3-sigma rectangle
        x0 = max(x0, (int)floorf((u - hx) / (float)GFB_TILE));
        x1 = min(x1, (int)floorf((u + hx) / (float)GFB_TILE) + 1);
        y0 = max(y0, (int)floorf((v - hy) / (float)GFB_TILE));
        y1 = min(y1, (int)floorf((v + hy) / (float)GFB_TILE) + 1);
    }
    return x1 > x0 && y1 > y0;
}

}  // namespace gfbm
