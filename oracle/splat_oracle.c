/*
 * splat_oracle.c -- plain-C restatement of the msplat operator path GFlow calls.
 *
 * TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 * The product (gflow_b200) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party package
 * `msplat` (github.com/pointrix-project/msplat, version unpinned by
 * /root/reference/README.md:27), which is absent from /root/reference; the
 * reference has no tests or golden vectors.  Each function
 * follows the call-site contract cited above it plus the published 3DGS-lineage
 * algorithm.  It is cross-checked against the independent
 * PyTorch restatement oracle/splat_ref.py (forward bit-for-bit on the integer
 * outputs, backward against autograd).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  No fused
 * multiply-add is allowed in the per-Gaussian geometry so float32 results are
 * reproducible bit for bit across this file, splat_ref.py and the CUDA kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_EPS 1e-4f
#define COV_BLUR 0.3f
#define FRUSTUM_CLAMP 1.3f

static inline void cam_point(const float *e, float x, float y, float z, float *xc, float *yc, float *zc) {
    *xc = ((e[0] * x + e[1] * y) + e[2] * z) + e[3];
    *yc = ((e[4] * x + e[5] * y) + e[6] * z) + e[7];
    *zc = ((e[8] * x + e[9] * y) + e[10] * z) + e[11];
}

/* ------------------------------------------------------------------ a1: project_point
 * call sites /root/reference/gflow/utils/render.py:21-24,116-119, trainer.py:955 */
static inline int project_one(const float *intr, const float *extr, int W, int H, float nearest, float extent,
                              float x, float y, float z, float *u, float *v, float *xc, float *yc, float *zc) {
    cam_point(extr, x, y, z, xc, yc, zc);
    if (!(*zc > nearest)) return 0;
    *u = (intr[0] * *xc) / *zc + intr[2];
    *v = (intr[1] * *yc) / *zc + intr[3];
    float xn = *u / (0.5f * (float)W) - 1.0f;
    float yn = *v / (0.5f * (float)H) - 1.0f;
    if (!(fabsf(xn) <= extent) || !(fabsf(yn) <= extent)) return 0;
    return 1;
}

void orc_project_point_fwd(const float *xyz, const float *intr, const float *extr, int N, int W, int H,
                           float nearest, float extent, float *uv, float *depth) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float u, v, xc, yc, zc;
        int ok = project_one(intr, extr, W, H, nearest, extent, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], &u, &v,
                             &xc, &yc, &zc);
        uv[2 * i] = ok ? u : 0.0f;
        uv[2 * i + 1] = ok ? v : 0.0f;
        depth[i] = ok ? zc : 0.0f;
    }
}

/* d_extr (12) and d_intr (4) are accumulated in double and written (not added). */
void orc_project_point_bwd(const float *xyz, const float *intr, const float *extr, int N, int W, int H,
                           float nearest, float extent, const float *g_uv, const float *g_depth, float *d_xyz,
                           float *d_intr, float *d_extr) {
    double ae[12] = {0}, ai[4] = {0};
#pragma omp parallel for schedule(static) reduction(+ : ae[:12], ai[:4])
    for (int i = 0; i < N; ++i) {
        float u, v, xc, yc, zc;
        float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        int ok = project_one(intr, extr, W, H, nearest, extent, x, y, z, &u, &v, &xc, &yc, &zc);
        if (!ok) {
            d_xyz[3 * i] = d_xyz[3 * i + 1] = d_xyz[3 * i + 2] = 0.0f;
            continue;
        }
        float gu = g_uv[2 * i], gv = g_uv[2 * i + 1], gd = g_depth ? g_depth[i] : 0.0f;
        float iz = 1.0f / zc;
        float gx = intr[0] * iz * gu;
        float gy = intr[1] * iz * gv;
        float gz = -(intr[0] * xc * iz * iz) * gu - (intr[1] * yc * iz * iz) * gv + gd;
        d_xyz[3 * i] = extr[0] * gx + extr[4] * gy + extr[8] * gz;
        d_xyz[3 * i + 1] = extr[1] * gx + extr[5] * gy + extr[9] * gz;
        d_xyz[3 * i + 2] = extr[2] * gx + extr[6] * gy + extr[10] * gz;
        float g[3] = {gx, gy, gz}, p[3] = {x, y, z};
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) ae[4 * r + c] += (double)g[r] * p[c];
            ae[4 * r + 3] += g[r];
        }
        ai[0] += (double)gu * xc * iz;
        ai[1] += (double)gv * yc * iz;
        ai[2] += gu;
        ai[3] += gv;
    }
    for (int k = 0; k < 12; ++k) d_extr[k] = (float)ae[k];
    for (int k = 0; k < 4; ++k) d_intr[k] = (float)ai[k];
}

/* ------------------------------------------------------------------ a3: compute_cov3d
 * call site /root/reference/gflow/utils/render.py:37-41; quaternion (w,x,y,z) */
static inline void quat_rot(const float *q, float *R) {
    float w = q[0], x = q[1], y = q[2], z = q[3];
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

void orc_compute_cov3d_fwd(const float *scale, const float *rotate, const uint8_t *visible, int N, float *cov3d) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float *o = cov3d + 6 * i;
        if (visible && !visible[i]) {
            for (int k = 0; k < 6; ++k) o[k] = 0.0f;
            continue;
        }
        float R[9], M[9];
        quat_rot(rotate + 4 * i, R);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * scale[3 * i + c];
        o[0] = (M[0] * M[0] + M[1] * M[1]) + M[2] * M[2];
        o[1] = (M[0] * M[3] + M[1] * M[4]) + M[2] * M[5];
        o[2] = (M[0] * M[6] + M[1] * M[7]) + M[2] * M[8];
        o[3] = (M[3] * M[3] + M[4] * M[4]) + M[5] * M[5];
        o[4] = (M[3] * M[6] + M[4] * M[7]) + M[5] * M[8];
        o[5] = (M[6] * M[6] + M[7] * M[7]) + M[8] * M[8];
    }
}

void orc_compute_cov3d_bwd(const float *scale, const float *rotate, const uint8_t *visible, int N,
                           const float *g_cov3d, float *d_scale, float *d_rotate) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float *ds = d_scale + 3 * i, *dq = d_rotate + 4 * i;
        if (visible && !visible[i]) {
            ds[0] = ds[1] = ds[2] = 0.0f;
            dq[0] = dq[1] = dq[2] = dq[3] = 0.0f;
            continue;
        }
        const float *q = rotate + 4 * i, *s = scale + 3 * i, *g = g_cov3d + 6 * i;
        float R[9], M[9], Gs[9], dM[9], D[9];
        quat_rot(q, R);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r + c] * s[c];
        Gs[0] = 2.0f * g[0]; Gs[1] = g[1]; Gs[2] = g[2];
        Gs[3] = g[1]; Gs[4] = 2.0f * g[3]; Gs[5] = g[4];
        Gs[6] = g[2]; Gs[7] = g[4]; Gs[8] = 2.0f * g[5];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                dM[3 * r + c] = Gs[3 * r] * M[c] + Gs[3 * r + 1] * M[3 + c] + Gs[3 * r + 2] * M[6 + c];
        for (int c = 0; c < 3; ++c) ds[c] = dM[c] * R[c] + dM[3 + c] * R[3 + c] + dM[6 + c] * R[6 + c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) D[3 * r + c] = dM[3 * r + c] * s[c];
        float w = q[0], x = q[1], y = q[2], z = q[3];
        dq[0] = 2.0f * (-z * D[1] + y * D[2] + z * D[3] - x * D[5] - y * D[6] + x * D[7]);
        dq[1] = 2.0f * (y * D[1] + z * D[2] + y * D[3] - 2.0f * x * D[4] - w * D[5] + z * D[6] + w * D[7] - 2.0f * x * D[8]);
        dq[2] = 2.0f * (-2.0f * y * D[0] + x * D[1] + w * D[2] + x * D[3] + z * D[5] - w * D[6] + z * D[7] - 2.0f * y * D[8]);
        dq[3] = 2.0f * (-2.0f * z * D[0] - w * D[1] + x * D[2] + w * D[3] - 2.0f * z * D[4] + y * D[5] + x * D[6] + y * D[7]);
    }
}

/* ------------------------------------------------------------------ a4: ewa_project
 * call site /root/reference/gflow/utils/render.py:44-49 */
static inline void tile_rect(float u, float v, float r, int gx, int gy, int *x0, int *y0, int *x1, int *y1) {
    int a;
    a = (int)((u - r) / (float)TILE); *x0 = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)((v - r) / (float)TILE); *y0 = a < 0 ? 0 : (a > gy ? gy : a);
    a = (int)(((u + r) + (float)(TILE - 1)) / (float)TILE); *x1 = a < 0 ? 0 : (a > gx ? gx : a);
    a = (int)(((v + r) + (float)(TILE - 1)) / (float)TILE); *y1 = a < 0 ? 0 : (a > gy ? gy : a);
}

typedef struct {
    float tx, ty, tz, txc, tyc, j00, j02, j11, j12, T0[3], T1[3], a, b, c, det;
    int clampx, clampy;
} ewa_mid;

static inline void ewa_mid_eval(const float *p, const float *S, const float *intr, const float *e, int W, int H,
                                ewa_mid *m) {
    float fx = intr[0], fy = intr[1];
    cam_point(e, p[0], p[1], p[2], &m->tx, &m->ty, &m->tz);
    float limx = FRUSTUM_CLAMP * ((float)W / (2.0f * fx));
    float limy = FRUSTUM_CLAMP * ((float)H / (2.0f * fy));
    float rx = m->tx / m->tz, ry = m->ty / m->tz;
    m->clampx = (rx < -limx) || (rx > limx);
    m->clampy = (ry < -limy) || (ry > limy);
    m->txc = fminf(limx, fmaxf(-limx, rx)) * m->tz;
    m->tyc = fminf(limy, fmaxf(-limy, ry)) * m->tz;
    m->j00 = fx / m->tz;
    m->j02 = -(fx * m->txc) / (m->tz * m->tz);
    m->j11 = fy / m->tz;
    m->j12 = -(fy * m->tyc) / (m->tz * m->tz);
    for (int k = 0; k < 3; ++k) {
        m->T0[k] = m->j00 * e[k] + m->j02 * e[8 + k];
        m->T1[k] = m->j11 * e[4 + k] + m->j12 * e[8 + k];
    }
    const float *T0 = m->T0, *T1 = m->T1;
    float a0 = (S[0] * T0[0] + S[1] * T0[1]) + S[2] * T0[2];
    float a1 = (S[1] * T0[0] + S[3] * T0[1]) + S[4] * T0[2];
    float a2 = (S[2] * T0[0] + S[4] * T0[1]) + S[5] * T0[2];
    float b0 = (S[0] * T1[0] + S[1] * T1[1]) + S[2] * T1[2];
    float b1 = (S[1] * T1[0] + S[3] * T1[1]) + S[4] * T1[2];
    float b2 = (S[2] * T1[0] + S[4] * T1[1]) + S[5] * T1[2];
    m->a = ((T0[0] * a0 + T0[1] * a1) + T0[2] * a2) + COV_BLUR;
    m->b = (T1[0] * a0 + T1[1] * a1) + T1[2] * a2;
    m->c = ((T1[0] * b0 + T1[1] * b1) + T1[2] * b2) + COV_BLUR;
    m->det = m->a * m->c - m->b * m->b;
}

void orc_ewa_project_fwd(const float *xyz, const float *cov3d, const float *intr, const float *extr, const float *uv,
                         int N, int W, int H, const uint8_t *visible, float *conic, int32_t *radius,
                         int32_t *tiles_touched) {
    int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        conic[3 * i] = conic[3 * i + 1] = conic[3 * i + 2] = 0.0f;
        radius[i] = 0;
        tiles_touched[i] = 0;
        if (visible && !visible[i]) continue;
        ewa_mid m;
        ewa_mid_eval(xyz + 3 * i, cov3d + 6 * i, intr, extr, W, H, &m);
        if (m.det == 0.0f) continue;
        float dinv = 1.0f / m.det;
        float mid = 0.5f * (m.a + m.c);
        float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - m.det));
        float rf = ceilf(3.0f * sqrtf(lam));
        int x0, y0, x1, y1;
        tile_rect(uv[2 * i], uv[2 * i + 1], rf, gx, gy, &x0, &y0, &x1, &y1);
        int area = (x1 - x0) * (y1 - y0);
        if (area <= 0) continue;
        conic[3 * i] = m.c * dinv;
        conic[3 * i + 1] = -m.b * dinv;
        conic[3 * i + 2] = m.a * dinv;
        radius[i] = (int32_t)rf;
        tiles_touched[i] = area;
    }
}

/* `live` = the forward's tiles_touched > 0 predicate, recomputed here from uv. */
void orc_ewa_project_bwd(const float *xyz, const float *cov3d, const float *intr, const float *extr, const float *uv,
                         int N, int W, int H, const uint8_t *visible, const float *g_conic, float *d_xyz,
                         float *d_cov3d, float *d_intr, float *d_extr) {
    int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    double ae[12] = {0}, ai[4] = {0};
    const float *e = extr;
#pragma omp parallel for schedule(static) reduction(+ : ae[:12], ai[:4])
    for (int i = 0; i < N; ++i) {
        float *dp = d_xyz + 3 * i, *dS = d_cov3d + 6 * i;
        dp[0] = dp[1] = dp[2] = 0.0f;
        for (int k = 0; k < 6; ++k) dS[k] = 0.0f;
        if (visible && !visible[i]) continue;
        const float *p = xyz + 3 * i, *S = cov3d + 6 * i;
        ewa_mid m;
        ewa_mid_eval(p, S, intr, extr, W, H, &m);
        if (m.det == 0.0f) continue;
        float mid = 0.5f * (m.a + m.c);
        float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - m.det));
        float rf = ceilf(3.0f * sqrtf(lam));
        int x0, y0, x1, y1;
        tile_rect(uv[2 * i], uv[2 * i + 1], rf, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) <= 0) continue;
        float gA = g_conic[3 * i], gB = g_conic[3 * i + 1], gC = g_conic[3 * i + 2];
        float a = m.a, b = m.b, c = m.c, dinv = 1.0f / m.det, d2 = dinv * dinv;
        float ga = d2 * (-c * c * gA + b * c * gB) + gC * (dinv - a * c * d2);
        float gb = 2.0f * b * c * d2 * gA + gB * (-dinv - 2.0f * b * b * d2) + 2.0f * a * b * d2 * gC;
        float gc = gA * (dinv - a * c * d2) + a * b * d2 * gB - a * a * d2 * gC;
        const float *T0 = m.T0, *T1 = m.T1;
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
        for (int k = 0; k < 3; ++k) {
            dT0[k] = 2.0f * ga * ST0[k] + gb * ST1[k];
            dT1[k] = 2.0f * gc * ST1[k] + gb * ST0[k];
        }
        float dj00 = 0, dj02 = 0, dj11 = 0, dj12 = 0, dR[9];
        for (int k = 0; k < 3; ++k) {
            dj00 += dT0[k] * e[k];
            dj02 += dT0[k] * e[8 + k];
            dj11 += dT1[k] * e[4 + k];
            dj12 += dT1[k] * e[8 + k];
            dR[k] = dT0[k] * m.j00;
            dR[3 + k] = dT1[k] * m.j11;
            dR[6 + k] = dT0[k] * m.j02 + dT1[k] * m.j12;
        }
        float fx = intr[0], fy = intr[1], tz = m.tz, iz = 1.0f / tz, iz2 = iz * iz, iz3 = iz2 * iz;
        ai[0] += (double)(dj00 * iz - dj02 * m.txc * iz2);
        ai[1] += (double)(dj11 * iz - dj12 * m.tyc * iz2);
        float dtxc = -dj02 * fx * iz2, dtyc = -dj12 * fy * iz2;
        float dtz = -dj00 * fx * iz2 + 2.0f * dj02 * fx * m.txc * iz3 - dj11 * fy * iz2 + 2.0f * dj12 * fy * m.tyc * iz3;
        float dtx = 0.0f, dty = 0.0f;
        if (m.clampx) dtz += dtxc * (m.txc * iz); else dtx = dtxc;
        if (m.clampy) dtz += dtyc * (m.tyc * iz); else dty = dtyc;
        float dt[3] = {dtx, dty, dtz};
        for (int k = 0; k < 3; ++k) dp[k] = e[k] * dt[0] + e[4 + k] * dt[1] + e[8 + k] * dt[2];
        for (int r = 0; r < 3; ++r) {
            for (int k = 0; k < 3; ++k) ae[4 * r + k] += (double)dR[3 * r + k] + (double)dt[r] * p[k];
            ae[4 * r + 3] += dt[r];
        }
    }
    for (int k = 0; k < 12; ++k) d_extr[k] = (float)ae[k];
    for (int k = 0; k < 4; ++k) d_intr[k] = (float)ai[k];
}

/* ------------------------------------------------------------------ a5: sort_gaussian
 * call site /root/reference/gflow/utils/render.py:52-54.
 * Returns K; ids_sorted must hold sum(tiles_touched) entries (call with
 * ids_sorted == NULL to get K only). */
typedef struct { uint64_t key; int32_t id; } sort_ent;

static void radix_sort_ent(sort_ent *a, sort_ent *tmp, int64_t n, int key_bits) {
    for (int shift = 0; shift < key_bits; shift += 8) {
        int64_t cnt[257] = {0};
        for (int64_t i = 0; i < n; ++i) cnt[((a[i].key >> shift) & 0xff) + 1]++;
        for (int k = 0; k < 256; ++k) cnt[k + 1] += cnt[k];
        for (int64_t i = 0; i < n; ++i) tmp[cnt[(a[i].key >> shift) & 0xff]++] = a[i];
        sort_ent *t = a; a = tmp; tmp = t;
    }
    /* key_bits is padded to a multiple of 16 by the caller => even pass count => result in `a` */
}

int64_t orc_sort_gaussian(const float *uv, const float *depth, int N, int W, int H, const int32_t *radius,
                          const int32_t *tiles_touched, int32_t *ids_sorted, int32_t *tile_range) {
    int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE, T = gx * gy;
    int64_t K = 0;
    for (int i = 0; i < N; ++i)
        if (radius[i] > 0 && tiles_touched[i] > 0) K += tiles_touched[i];
    if (!ids_sorted) return K;
    memset(tile_range, 0, sizeof(int32_t) * 2 * (size_t)T);
    if (K == 0) return 0;
    sort_ent *a = (sort_ent *)malloc(sizeof(sort_ent) * (size_t)K), *tmp = (sort_ent *)malloc(sizeof(sort_ent) * (size_t)K);
    int64_t k = 0;
    for (int i = 0; i < N; ++i) {
        if (!(radius[i] > 0 && tiles_touched[i] > 0)) continue;
        int x0, y0, x1, y1;
        tile_rect(uv[2 * i], uv[2 * i + 1], (float)radius[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t db;
        memcpy(&db, depth + i, 4);
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                if (k >= K) break; /* tiles_touched inconsistent with the rect: never overrun */
                a[k].key = ((uint64_t)(y * gx + x) << 32) | db;
                a[k].id = i;
                ++k;
            }
    }
    K = k;
    radix_sort_ent(a, tmp, K, 64);
    for (int64_t i = 0; i < K; ++i) {
        ids_sorted[i] = a[i].id;
        int t = (int)(a[i].key >> 32);
        if (i == 0 || (int)(a[i - 1].key >> 32) != t) tile_range[2 * t] = (int32_t)i;
        if (i == K - 1 || (int)(a[i + 1].key >> 32) != t) tile_range[2 * t + 1] = (int32_t)(i + 1);
    }
    free(a);
    free(tmp);
    return K;
}

/* ------------------------------------------------------------------ a6: alpha_blending forward
 * call sites /root/reference/gflow/utils/render.py:58-64,68-74,84-90,99-105,148-154 */
void orc_alpha_blending_fwd(const float *uv, const float *conic, const float *opacity, const float *feature, int C,
                            const int32_t *ids_sorted, const int32_t *tile_range, float bg, int W, int H, float *out,
                            float *final_T, int32_t *n_contrib) {
    int gx = (W + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 64)
    for (int pix = 0; pix < W * H; ++pix) {
        int px = pix % W, py = pix / W;
        int t = (py / TILE) * gx + (px / TILE);
        int s = tile_range[2 * t], e = tile_range[2 * t + 1];
        float T = 1.0f;
        float acc[64];
        for (int c = 0; c < C; ++c) acc[c] = 0.0f;
        int last = 0;
        for (int j = s; j < e; ++j) {
            int g = ids_sorted[j];
            float dx = uv[2 * g] - (float)px, dy = uv[2 * g + 1] - (float)py;
            float power = -0.5f * (conic[3 * g] * dx * dx + conic[3 * g + 2] * dy * dy) - conic[3 * g + 1] * dx * dy;
            if (power > 0.0f) continue;
            float alpha = fminf(ALPHA_MAX, opacity[g] * expf(power));
            if (alpha < ALPHA_MIN) continue;
            float test_T = T * (1.0f - alpha);
            if (test_T < T_EPS) break;
            for (int c = 0; c < C; ++c) acc[c] += feature[(size_t)g * C + c] * alpha * T;
            T = test_T;
            last = j - s + 1;
        }
        for (int c = 0; c < C; ++c) out[(size_t)c * W * H + pix] = acc[c] + T * bg;
        if (final_T) final_T[pix] = T;
        if (n_contrib) n_contrib[pix] = last;
    }
}

/* ------------------------------------------------------------------ a7: alpha_blending backward
 * (autograd, triggered by /root/reference/gflow/trainer.py:533).  Serial over
 * pixels; gradient buffers are written (zeroed first), accumulated in float. */
void orc_alpha_blending_bwd(const float *uv, const float *conic, const float *opacity, const float *feature, int C,
                            int N, const int32_t *ids_sorted, const int32_t *tile_range, float bg, int W, int H,
                            const float *final_T, const int32_t *n_contrib, const float *g_out, float *d_uv,
                            float *d_conic, float *d_opacity, float *d_feature) {
    int gx = (W + TILE - 1) / TILE;
    memset(d_uv, 0, sizeof(float) * 2 * (size_t)N);
    memset(d_conic, 0, sizeof(float) * 3 * (size_t)N);
    memset(d_opacity, 0, sizeof(float) * (size_t)N);
    memset(d_feature, 0, sizeof(float) * (size_t)C * N);
    /* double accumulators keep the oracle independent of pixel order */
    double *A = (double *)calloc((size_t)N * (6 + C), sizeof(double));
#pragma omp parallel
    {
    double *Ath = (double *)calloc((size_t)N * (6 + C), sizeof(double));
#pragma omp for schedule(dynamic, 256)
    for (int pix = 0; pix < W * H; ++pix) {
        int px = pix % W, py = pix / W;
        int t = (py / TILE) * gx + (px / TILE);
        int s = tile_range[2 * t];
        float Tf = final_T[pix], T = Tf;
        int last = n_contrib[pix];
        float go[64], accum[64], lastc[64];
        float bgdot = 0.0f;
        for (int c = 0; c < C; ++c) {
            go[c] = g_out[(size_t)c * W * H + pix];
            accum[c] = 0.0f;
            lastc[c] = 0.0f;
            bgdot += bg * go[c];
        }
        float last_alpha = 0.0f;
        for (int j = s + last - 1; j >= s; --j) {
            int g = ids_sorted[j];
            float dx = uv[2 * g] - (float)px, dy = uv[2 * g + 1] - (float)py;
            float ca = conic[3 * g], cb = conic[3 * g + 1], cc = conic[3 * g + 2];
            float power = -0.5f * (ca * dx * dx + cc * dy * dy) - cb * dx * dy;
            if (power > 0.0f) continue;
            float G = expf(power);
            float alpha = fminf(ALPHA_MAX, opacity[g] * G);
            if (alpha < ALPHA_MIN) continue;
            T = T / (1.0f - alpha);
            float w = alpha * T;
            float dalpha = 0.0f;
            double *acc = Ath + (size_t)g * (6 + C);
            for (int c = 0; c < C; ++c) {
                float f = feature[(size_t)g * C + c];
                accum[c] = last_alpha * lastc[c] + (1.0f - last_alpha) * accum[c];
                lastc[c] = f;
                dalpha += (f - accum[c]) * go[c];
                acc[6 + c] += (double)(w * go[c]);
            }
            dalpha *= T;
            last_alpha = alpha;
            dalpha += (-Tf / (1.0f - alpha)) * bgdot;
            float dG = opacity[g] * dalpha;
            float gdx = G * dx, gdy = G * dy;
            acc[0] += (double)(dG * (-gdx * ca - gdy * cb));
            acc[1] += (double)(dG * (-gdy * cc - gdx * cb));
            acc[2] += (double)(-0.5f * gdx * dx * dG);
            acc[3] += (double)(-gdx * dy * dG);
            acc[4] += (double)(-0.5f * gdy * dy * dG);
            acc[5] += (double)(G * dalpha);
        }
    }
#pragma omp critical
    for (size_t k = 0; k < (size_t)N * (6 + C); ++k) A[k] += Ath[k];
    free(Ath);
    }
    for (int g = 0; g < N; ++g) {
        double *acc = A + (size_t)g * (6 + C);
        d_uv[2 * g] = (float)acc[0];
        d_uv[2 * g + 1] = (float)acc[1];
        d_conic[3 * g] = (float)acc[2];
        d_conic[3 * g + 1] = (float)acc[3];
        d_conic[3 * g + 2] = (float)acc[4];
        d_opacity[g] = (float)acc[5];
        for (int c = 0; c < C; ++c) d_feature[(size_t)g * C + c] = (float)acc[6 + c];
    }
    free(A);
}

/* ------------------------------------------------------------------ a8: compute_sh
 * (no GFlow call site; north_star operator surface, BASELINE config 5).
 * shs (N,C,K) with K in {1,4,9,16}; dirs (N,3) un-normalised. */
static const float SH_C0 = 0.28209479177387814f, SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

static void sh_eval(float x, float y, float z, int K, float *Y, float *Yx, float *Yy, float *Yz) {
    for (int k = 0; k < 16; ++k) Y[k] = Yx[k] = Yy[k] = Yz[k] = 0.0f;
    Y[0] = SH_C0;
    if (K > 1) {
        Y[1] = -SH_C1 * y; Yy[1] = -SH_C1;
        Y[2] = SH_C1 * z;  Yz[2] = SH_C1;
        Y[3] = -SH_C1 * x; Yx[3] = -SH_C1;
    }
    float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
    if (K > 4) {
        Y[4] = SH_C2[0] * xy; Yx[4] = SH_C2[0] * y; Yy[4] = SH_C2[0] * x;
        Y[5] = SH_C2[1] * yz; Yy[5] = SH_C2[1] * z; Yz[5] = SH_C2[1] * y;
        Y[6] = SH_C2[2] * (2.0f * zz - xx - yy);
        Yx[6] = SH_C2[2] * -2.0f * x; Yy[6] = SH_C2[2] * -2.0f * y; Yz[6] = SH_C2[2] * 4.0f * z;
        Y[7] = SH_C2[3] * xz; Yx[7] = SH_C2[3] * z; Yz[7] = SH_C2[3] * x;
        Y[8] = SH_C2[4] * (xx - yy); Yx[8] = SH_C2[4] * 2.0f * x; Yy[8] = SH_C2[4] * -2.0f * y;
    }
    if (K > 9) {
        Y[9] = SH_C3[0] * y * (3.0f * xx - yy);
        Yx[9] = SH_C3[0] * 6.0f * xy; Yy[9] = SH_C3[0] * (3.0f * xx - 3.0f * yy);
        Y[10] = SH_C3[1] * xy * z;
        Yx[10] = SH_C3[1] * yz; Yy[10] = SH_C3[1] * xz; Yz[10] = SH_C3[1] * xy;
        Y[11] = SH_C3[2] * y * (4.0f * zz - xx - yy);
        Yx[11] = SH_C3[2] * -2.0f * xy; Yy[11] = SH_C3[2] * (4.0f * zz - xx - 3.0f * yy); Yz[11] = SH_C3[2] * 8.0f * yz;
        Y[12] = SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
        Yx[12] = SH_C3[3] * -6.0f * xz; Yy[12] = SH_C3[3] * -6.0f * yz; Yz[12] = SH_C3[3] * (6.0f * zz - 3.0f * xx - 3.0f * yy);
        Y[13] = SH_C3[4] * x * (4.0f * zz - xx - yy);
        Yx[13] = SH_C3[4] * (4.0f * zz - 3.0f * xx - yy); Yy[13] = SH_C3[4] * -2.0f * xy; Yz[13] = SH_C3[4] * 8.0f * xz;
        Y[14] = SH_C3[5] * z * (xx - yy);
        Yx[14] = SH_C3[5] * 2.0f * xz; Yy[14] = SH_C3[5] * -2.0f * yz; Yz[14] = SH_C3[5] * (xx - yy);
        Y[15] = SH_C3[6] * x * (xx - 3.0f * yy);
        Yx[15] = SH_C3[6] * (3.0f * xx - 3.0f * yy); Yy[15] = SH_C3[6] * -6.0f * xy;
    }
}

void orc_compute_sh_fwd(const float *shs, const float *dirs, const uint8_t *visible, int N, int C, int K, float *out) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float *o = out + (size_t)i * C;
        if (visible && !visible[i]) {
            for (int c = 0; c < C; ++c) o[c] = 0.0f;
            continue;
        }
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float inv = 1.0f / sqrtf(x * x + y * y + z * z);
        float Y[16], Yx[16], Yy[16], Yz[16];
        sh_eval(x * inv, y * inv, z * inv, K, Y, Yx, Yy, Yz);
        for (int c = 0; c < C; ++c) {
            const float *s = shs + ((size_t)i * C + c) * K;
            float acc = 0.0f;
            for (int k = 0; k < K; ++k) acc += s[k] * Y[k];
            o[c] = acc;
        }
    }
}

void orc_compute_sh_bwd(const float *shs, const float *dirs, const uint8_t *visible, int N, int C, int K,
                        const float *g_out, float *d_shs, float *d_dirs) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N; ++i) {
        float *dd = d_dirs + 3 * i;
        dd[0] = dd[1] = dd[2] = 0.0f;
        if (visible && !visible[i]) {
            for (int k = 0; k < C * K; ++k) d_shs[(size_t)i * C * K + k] = 0.0f;
            continue;
        }
        float x = dirs[3 * i], y = dirs[3 * i + 1], z = dirs[3 * i + 2];
        float inv = 1.0f / sqrtf(x * x + y * y + z * z);
        float nx = x * inv, ny = y * inv, nz = z * inv;
        float Y[16], Yx[16], Yy[16], Yz[16];
        sh_eval(nx, ny, nz, K, Y, Yx, Yy, Yz);
        float gx = 0.0f, gy = 0.0f, gz = 0.0f;
        for (int c = 0; c < C; ++c) {
            const float *s = shs + ((size_t)i * C + c) * K;
            float g = g_out[(size_t)i * C + c];
            for (int k = 0; k < K; ++k) {
                d_shs[((size_t)i * C + c) * K + k] = g * Y[k];
                gx += g * s[k] * Yx[k];
                gy += g * s[k] * Yy[k];
                gz += g * s[k] * Yz[k];
            }
        }
        float dot = gx * nx + gy * ny + gz * nz;
        dd[0] = (gx - nx * dot) * inv;
        dd[1] = (gy - ny * dot) * inv;
        dd[2] = (gz - nz * dot) * inv;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
