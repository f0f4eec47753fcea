// fit.cu -- the per-frame optimisation iteration of GFlow as one stream of sm_100a kernels.
//
// Reference: the inner loop of SimpleGaussian.train, /root/reference/gflow/trainer.py:387-558, driven per
// frame by /root/reference/gflow/fit_video.py:119-142,256-315.  In the reference every iteration is
// ~150 PyTorch kernel launches plus several host synchronisations (loss.item(), colormap round trip)
// around five msplat operators; the render itself is a small part of the iteration.  Here the whole
// iteration is eight launches (ten with SSIM) with no host synchronisation:
//
//   fit_preprocess      raw attributes -> activations (trainer.py:62-69) + project + cov3d + EWA + tile
//                       counting + scan (the fused pipeline's preprocess, fed from raw parameters); also
//                       the per-Gaussian regularisers' loss sums (trainer.py:490-503)
//   scatter, tile_sort_pack   (pipeline.cu, unchanged)
//   blend_fwd<C>        (blend.cu, unchanged) C = 4: rgb and the depth map in ONE tile walk -- render.py:58-74
//                       issues two blends over the same sort; depth is simply the fourth feature channel
//   ssim_stats / ssim_grad    optional, pytorch_ssim.py:17-37 forward + its analytic backward
//   fit_loss            mean squared error + scale/shift invariant depth loss -> dL/d(out), d depth_a/b
//   blend_bwd<C>        (blend.cu, unchanged)
//   fit_geometry_bwd_adam     packed gradients -> EWA / cov3d / projection backward -> activation backward
//                       -> gradient masks (trainer.py:535-551) -> Adam update of all 14 scalars of the
//                       Gaussian in registers (torch.optim.Adam semantics), camera gradients block-reduced
//   fit_finish          d(extr) -> d(pose) through the unit-quaternion map (trainer.py:115-121), Adam on the
//                       pose and depth_a / depth_b, loss history, camera of the next iteration
//
// All of it is HBM / latency bound per-Gaussian and per-pixel streaming around the two blend kernels;
// nothing here is GEMM shaped.  Compiled with -fmad=false like geometry.cu / pipeline.cu so the
// per-Gaussian geometry is bit-identical to the operator path for the same activated inputs.
#include <cstdlib>

#include "sort_network.cuh"
#include "splat_math.cuh"

// pipeline.cu / blend.cu
bool gfb_tight_tiles();
int gfb_internal_scatter_sort_pack_blend(const void*, const float*, int, int, int, void*, int64_t, void*, int32_t*, const float*,
                                         const float*, const float*, const float*, int, int32_t*, void*, void*, float, float*,
                                         float*, int32_t*, void*, bool pdl);
int gfb_internal_scatter_sort_pack(const void*, const float*, int, int, int, void*, int64_t, void*, int32_t*,
                                   const float*, const float*, const float*, const float*, int, int32_t*, void*, void*,
                                   void*, bool);
int gfb_internal_blend_fwd(const void*, const void*, int64_t, const int32_t*, int, int, int, float, int, int, float*,
                           float*, int32_t*, void*, bool pdl);
int gfb_internal_blend_bwd(const void*, const void*, int64_t, const int32_t*, const int32_t*, int, int, int, float, int, int,
                           const float*, const int32_t*, const float*, float*, void*, bool no_rgb, float* zero16);

namespace {

using namespace gfbm;

enum { ST_ITER = 0, ST_K_LAST = 1, ST_K_MAX = 2, ST_SUB_K_MAX = 3, ST_WORDS = 16 };
// loss accumulators (sums; fit_finish turns them into means)
enum { LA_SQ = 0, LA_DEPTH = 1, LA_DA = 2, LA_DB = 3, LA_SSIM = 4, LA_VAR = 5, LA_SCALE = 6, LA_NSCALE = 7, LA_STILL = 8,
       LA_FLOW = 9, LA_WORDS = 12 };
enum { HIST_WORDS = 8 };

// per-launch Adam constants, computed on the host in double: step = lr * LinearLR factor / (1 - beta1^t)
struct AdamStep {
    float step, inv_sqrt_bc2, b1, b2, eps;
};

__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamStep& a) {
    m = a.b1 * m + (1.0f - a.b1) * g;
    v = a.b2 * v + (1.0f - a.b2) * g * g;
    const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
    p = p - a.step * (m / denom);
}

__device__ __forceinline__ float sigmoid1(float x) { return 1.0f / (1.0f + expf(-x)); }

// F.normalize: q / max(|q|, 1e-12)
__device__ __forceinline__ float4 normalize4(float4 q, float& n) {
    n = fmaxf(sqrtf(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w), 1e-12f);
    return make_float4(q.x / n, q.y / n, q.z / n, q.w / n);
}

// pose (qx qy qz qw tx ty tz) -> [R | t] row-major 3x4, roma.RigidUnitQuat(...).normalize().to_homogeneous()[:3]
__device__ __forceinline__ void pose_to_extr(const float* pose, float* e) {
    float n;
    const float4 q = normalize4(make_float4(pose[0], pose[1], pose[2], pose[3]), n);
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    e[0] = 1.0f - 2.0f * (y * y + z * z); e[1] = 2.0f * (x * y - w * z); e[2] = 2.0f * (x * z + w * y); e[3] = pose[4];
    e[4] = 2.0f * (x * y + w * z); e[5] = 1.0f - 2.0f * (x * x + z * z); e[6] = 2.0f * (y * z - w * x); e[7] = pose[5];
    e[8] = 2.0f * (x * z - w * y); e[9] = 2.0f * (y * z + w * x); e[10] = 1.0f - 2.0f * (x * x + y * y); e[11] = pose[6];
}

// dL/d(extr) (3x4 row-major) -> dL/d(pose) through the normalised quaternion
__device__ __forceinline__ void extr_grad_to_pose(const float* pose, const float* d, float* dp) {
    float n;
    const float4 q = normalize4(make_float4(pose[0], pose[1], pose[2], pose[3]), n);
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    const float r00 = d[0], r01 = d[1], r02 = d[2], r10 = d[4], r11 = d[5], r12 = d[6], r20 = d[8], r21 = d[9], r22 = d[10];
    const float gx = 2.0f * (y * r01 + z * r02 + y * r10 - 2.0f * x * r11 - w * r12 + z * r20 + w * r21 - 2.0f * x * r22);
    const float gy = 2.0f * (-2.0f * y * r00 + x * r01 + w * r02 + x * r10 + z * r12 - w * r20 + z * r21 - 2.0f * y * r22);
    const float gz = 2.0f * (-2.0f * z * r00 - w * r01 + x * r02 + w * r10 - 2.0f * z * r11 + y * r12 + x * r20 + y * r21);
    const float gw = 2.0f * (-z * r01 + y * r02 + z * r10 - x * r12 - y * r20 + x * r21);
    const float dot = x * gx + y * gy + z * gz + w * gw;
    dp[0] = (gx - x * dot) / n;
    dp[1] = (gy - y * dot) / n;
    dp[2] = (gz - z * dot) / n;
    dp[3] = (gw - w * dot) / n;
    dp[4] = d[3];
    dp[5] = d[7];
    dp[6] = d[11];
}

// per-Gaussian terms of trainer.py:490-530 that do not go through the image
struct FitRegs {
    float lambda_var, lambda_scale;
    const uint8_t* scale_sel;   // loss_scale only over these (NULL = every in-image Gaussian)
    const float* still_ref;     // loss_still = mean_sel |xyz - still_ref|
    const uint8_t* still_sel;
    int n_still_ref;
    float w_still;              // lambda_still / still_count
    const float* flow_target;   // loss_flow = mean_sel,2 (uv - flow_target)^2
    const uint8_t* flow_sel;
    int n_flow;
    float w_flow;               // lambda_flow / (2 flow_count)
    __device__ __forceinline__ bool any() const {
        return lambda_var != 0.0f || lambda_scale != 0.0f || w_still != 0.0f || w_flow != 0.0f;
    }
};

// ------------------------------------------------------------------ init
__global__ void fit_init_kernel(const float* __restrict__ pose, const float* __restrict__ intr, float* __restrict__ cam,
                                int32_t* __restrict__ status, float* __restrict__ loss_acc) {
    const int t = threadIdx.x;
    if (t < ST_WORDS) status[t] = 0;
    if (t < LA_WORDS) loss_acc[t] = 0.0f;
    if (t == 0) pose_to_extr(pose, cam);
    if (t < 4) cam[12 + t] = intr[t];
}

// ------------------------------------------------------------------ forward: activations + geometry + binning
__global__ void __launch_bounds__(kThreads)
fit_preprocess_kernel(const float* __restrict__ xyz, const float* __restrict__ scale_raw,
                      const float4* __restrict__ rot_raw, const float* __restrict__ op_raw,
                      const float* __restrict__ rgb_raw, const float* __restrict__ cam, int N, int W, int H,
                      float nearest, float extent, int C, float2* __restrict__ uv, float* __restrict__ depth,
                      float* __restrict__ conic, int32_t* __restrict__ radius, ushort4* __restrict__ rect,
                      float* __restrict__ op_act, float* __restrict__ feat, int32_t* __restrict__ counts,
                      int32_t* __restrict__ offsets, int32_t* __restrict__ ctrl, int T, int R, FitRegs rg,
                      float* __restrict__ loss_acc, float* __restrict__ dbg_act, int tight) {
    __shared__ float s_cam[16];
    __shared__ int s_scan[34];
    __shared__ int s_buf[kScanSmemInts];
    static_assert(kThreads == kScanThreads, "the last CTA of fit_preprocess runs the scan");
    __shared__ bool s_last;
    gfb_pdl_launch_dependents();  // with GFB_FIT_PDL=1 scatter may take SM slots while the counting tail drains
    load_camera(s_cam, cam + 12, cam);
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    ushort4 rc = make_ushort4(0, 0, 0, 0);
    const float lambda_var = rg.lambda_var, lambda_scale = rg.lambda_scale;
    // sums of: std(scale), |scale| / depth, count of the latter, |xyz - still_ref|, (uv - flow_target)^2
    float reg[5] = {0.0f, 0.0f, 0.0f, 0.0f, 0.0f};
    if (i < N) {
        const float p[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const float s[3] = {fabsf(scale_raw[3 * i]), fabsf(scale_raw[3 * i + 1]), fabsf(scale_raw[3 * i + 2])};
        float qn;
        const float4 q = normalize4(rot_raw[i], qn);
        const float o = sigmoid1(10.0f * op_raw[i]);
        const float cr = sigmoid1(rgb_raw[3 * i]), cg = sigmoid1(rgb_raw[3 * i + 1]), cb = sigmoid1(rgb_raw[3 * i + 2]);
        float u, v, xc, yc, zc;
        const bool ok = project_one(s_cam + 12, s_cam, W, H, nearest, extent, p[0], p[1], p[2], u, v, xc, yc, zc);
        float ca = 0.0f, cb2 = 0.0f, cc = 0.0f;
        int rad = 0;
        if (ok) {
            float S[6];
            cov3d_fwd_one(s, q, S);
            EwaMid m;
            ewa_mid_eval(p, S, s_cam + 12, s_cam, W, H, m);
            float rf;
            int x0, y0, x1, y1;
            if (ewa_live(m, u, v, gx, gy, rf, x0, y0, x1, y1)) {
                const float dinv = 1.0f / m.det;
                ca = m.c * dinv;
                cb2 = -m.b * dinv;
                cc = m.a * dinv;
                rad = (int)rf;
                if (!tight || tighten_rect(u, v, ca, cb2, cc, o, x0, y0, x1, y1))  // GFB_TIGHT_TILES (pipeline.cu)
                    rc = make_ushort4((unsigned short)x0, (unsigned short)y0, (unsigned short)x1, (unsigned short)y1);
            }
        }
        uv[i] = ok ? make_float2(u, v) : make_float2(0.0f, 0.0f);
        depth[i] = ok ? zc : 0.0f;
        conic[3 * i] = ca;
        conic[3 * i + 1] = cb2;
        conic[3 * i + 2] = cc;
        radius[i] = rad;
        rect[i] = rc;
        op_act[i] = o;
        float* f = feat + (size_t)i * C;
        f[0] = cr;
        f[1] = cg;
        f[2] = cb;
        if (C > 3) f[3] = ok ? zc : 0.0f;  // the depth map is the fourth blended channel (render.py:68-74)
        if (lambda_var != 0.0f) {  // trainer.py:490-492: mean over Gaussians of the unbiased std of the 3 scales
            const float mu = (s[0] + s[1] + s[2]) / 3.0f;
            const float d0 = s[0] - mu, d1 = s[1] - mu, d2 = s[2] - mu;
            reg[0] = sqrtf((d0 * d0 + d1 * d1 + d2 * d2) * 0.5f);
        }
        if (lambda_scale != 0.0f && ok && u > 0.0f && u < (float)(W - 1) && v > 0.0f && v < (float)(H - 1) &&
            (!rg.scale_sel || rg.scale_sel[i])) {
            reg[1] = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]) / zc;  // trainer.py:495-501
            reg[2] = 1.0f;
        }
        if (rg.w_still != 0.0f && i < rg.n_still_ref && rg.still_sel[i]) {  // trainer.py:504-508
            const float d0 = p[0] - rg.still_ref[3 * i], d1 = p[1] - rg.still_ref[3 * i + 1], d2 = p[2] - rg.still_ref[3 * i + 2];
            reg[3] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
        }
        if (rg.w_flow != 0.0f && i < rg.n_flow && rg.flow_sel[i]) {  // trainer.py:510-530 (uv of a culled point is 0)
            const float d0 = (ok ? u : 0.0f) - rg.flow_target[2 * i], d1 = (ok ? v : 0.0f) - rg.flow_target[2 * i + 1];
            reg[4] = d0 * d0 + d1 * d1;
        }
        if (dbg_act) {
            float* a = dbg_act + (size_t)i * 14;
            a[0] = p[0]; a[1] = p[1]; a[2] = p[2]; a[3] = s[0]; a[4] = s[1]; a[5] = s[2];
            a[6] = q.x; a[7] = q.y; a[8] = q.z; a[9] = q.w; a[10] = o; a[11] = cr; a[12] = cg; a[13] = cb;
        }
    }
    static_assert(LA_SCALE == LA_VAR + 1 && LA_NSCALE == LA_VAR + 2 && LA_STILL == LA_VAR + 3 && LA_FLOW == LA_VAR + 4,
                  "the five per-Gaussian sums are reduced as one block");
    if (rg.any()) block_reduce_atomic<5>(reg, loss_acc + LA_VAR);
    {   // per-tile counting, 32 (Gaussian, tile) pairs per warp round (as in pipeline.cu preprocess)
        const WarpTileWalk walk(rc.x, rc.y, rc.z - rc.x, rc.w - rc.y, gx, threadIdx.x & 31);
        for (int base = 0; base < walk.total; base += 32) {
            int owner;
            const int t = walk.item(base, owner);
            if (t >= 0) red_add_s32(counts + t * R + (blockIdx.x % R), 1);
        }
    }
    // the last CTA to get here scans the tile counters
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(ctrl + GFB_CTRL_DONE, 1);
        s_last = (ticket == (int)gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int total = cta_exclusive_scan(counts, T * R, offsets, s_buf, s_scan);
    if (threadIdx.x == 0) ctrl[GFB_CTRL_K] = total;
}

// ------------------------------------------------------------------ SSIM (pytorch_ssim.py:7-37)
// 11-tap Gaussian, sigma 1.5, the float32 values torch produces for gaussian(11, 1.5)
__constant__ float kSsimW[11] = {0.001028380123898387f, 0.0075987582094967365f, 0.036000773310661316f,
                                 0.10936068743467331f,  0.21300552785396576f,   0.26601171493530273f,
                                 0.21300552785396576f,  0.10936068743467331f,   0.036000773310661316f,
                                 0.0075987582094967365f, 0.001028380123898387f};
constexpr int kSsimHalo = 5;
constexpr int kSsimIn = GFB_TILE + 2 * kSsimHalo;  // 26
constexpr float kSsimC1 = 0.01f * 0.01f, kSsimC2 = 0.03f * 0.03f;

// One CTA per 16x16 tile and channel.  x = rendered * mask, y = target * mask (zero outside the image =
// conv2d's zero padding).  Writes the per-pixel partial derivatives of the SSIM map with respect to the
// three windowed moments that depend on x: dS/dmu1, dS/dE[x^2], dS/dE[xy] (maps (3,3,H,W)).
__global__ void __launch_bounds__(256)
ssim_stats_kernel(const float* __restrict__ out, const float* __restrict__ gt_image, const uint8_t* __restrict__ mask,
                  int W, int H, float* __restrict__ maps, float* __restrict__ loss_acc) {
    __shared__ float s_x[kSsimIn][kSsimIn + 1], s_y[kSsimIn][kSsimIn + 1];
    __shared__ float s_h[5][kSsimIn][GFB_TILE];
    const int ch = blockIdx.z;
    const size_t P = (size_t)W * H;
    const int ox = blockIdx.x * GFB_TILE - kSsimHalo, oy = blockIdx.y * GFB_TILE - kSsimHalo;
    for (int idx = threadIdx.x; idx < kSsimIn * kSsimIn; idx += 256) {
        const int r = idx / kSsimIn, c = idx - r * kSsimIn;
        const int px = ox + c, py = oy + r;
        float x = 0.0f, y = 0.0f;
        if (px >= 0 && px < W && py >= 0 && py < H) {
            const size_t pix = (size_t)py * W + px;
            const float m = mask ? (mask[pix] ? 1.0f : 0.0f) : 1.0f;
            x = out[ch * P + pix] * m;
            y = gt_image[pix * 3 + ch] * m;
        }
        s_x[r][c] = x;
        s_y[r][c] = y;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kSsimIn * GFB_TILE; idx += 256) {
        const int r = idx / GFB_TILE, c = idx - r * GFB_TILE;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f, a4 = 0.0f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = kSsimW[k], x = s_x[r][c + k], y = s_y[r][c + k];
            a0 += w * x;
            a1 += w * y;
            a2 += w * (x * x);
            a3 += w * (y * y);
            a4 += w * (x * y);
        }
        s_h[0][r][c] = a0; s_h[1][r][c] = a1; s_h[2][r][c] = a2; s_h[3][r][c] = a3; s_h[4][r][c] = a4;
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int px = blockIdx.x * GFB_TILE + lx, py = blockIdx.y * GFB_TILE + ly;
    float acc[1] = {0.0f};
    if (px < W && py < H) {
        float mu1 = 0.0f, mu2 = 0.0f, e11 = 0.0f, e22 = 0.0f, e12 = 0.0f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = kSsimW[k];
            mu1 += w * s_h[0][ly + k][lx];
            mu2 += w * s_h[1][ly + k][lx];
            e11 += w * s_h[2][ly + k][lx];
            e22 += w * s_h[3][ly + k][lx];
            e12 += w * s_h[4][ly + k][lx];
        }
        const float A1 = 2.0f * mu1 * mu2 + kSsimC1;
        const float A2 = 2.0f * (e12 - mu1 * mu2) + kSsimC2;
        const float B1 = mu1 * mu1 + mu2 * mu2 + kSsimC1;
        const float B2 = (e11 - mu1 * mu1) + (e22 - mu2 * mu2) + kSsimC2;
        const float inv = 1.0f / (B1 * B2);
        acc[0] = A1 * A2 * inv;
        // d/dmu1 with E[x^2], E[xy] held fixed: A1' = 2 mu2, A2' = -2 mu2, B1' = 2 mu1, B2' = -2 mu1
        const float dmu = (2.0f * mu2 * (A2 - A1)) * inv - (A1 * A2) * (2.0f * mu1 * (B2 - B1)) * inv * inv;
        const float d11 = -(A1 * A2) * inv / B2;
        const float d12 = 2.0f * A1 * inv;
        const size_t pix = (size_t)py * W + px;
        maps[(0 * 3 + ch) * P + pix] = dmu;
        maps[(1 * 3 + ch) * P + pix] = d11;
        maps[(2 * 3 + ch) * P + pix] = d12;
    }
    block_reduce_atomic<1>(acc, loss_acc + LA_SSIM);
}

// dL/dx(p) = -w_ssim * sum_q win(q - p) [dS/dmu1(q) + 2 x(p) dS/dE11(q) + y(p) dS/dE12(q)]   (L = w_ssim * sum(1 - S))
__global__ void __launch_bounds__(256)
ssim_grad_kernel(const float* __restrict__ out, const float* __restrict__ gt_image, const uint8_t* __restrict__ mask,
                 int W, int H, const float* __restrict__ maps, float w_ssim, float* __restrict__ ssim_grad) {
    __shared__ float s_m[3][kSsimIn][kSsimIn + 1];
    __shared__ float s_h[3][kSsimIn][GFB_TILE];
    const int ch = blockIdx.z;
    const size_t P = (size_t)W * H;
    const int ox = blockIdx.x * GFB_TILE - kSsimHalo, oy = blockIdx.y * GFB_TILE - kSsimHalo;
    for (int idx = threadIdx.x; idx < kSsimIn * kSsimIn; idx += 256) {
        const int r = idx / kSsimIn, c = idx - r * kSsimIn;
        const int px = ox + c, py = oy + r;
        const bool in = px >= 0 && px < W && py >= 0 && py < H;
        const size_t pix = in ? (size_t)py * W + px : 0;
#pragma unroll
        for (int j = 0; j < 3; ++j) s_m[j][r][c] = in ? maps[(j * 3 + ch) * P + pix] : 0.0f;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kSsimIn * GFB_TILE; idx += 256) {
        const int r = idx / GFB_TILE, c = idx - r * GFB_TILE;
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const float w = kSsimW[k];
            a0 += w * s_m[0][r][c + k];
            a1 += w * s_m[1][r][c + k];
            a2 += w * s_m[2][r][c + k];
        }
        s_h[0][r][c] = a0; s_h[1][r][c] = a1; s_h[2][r][c] = a2;
    }
    __syncthreads();
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int px = blockIdx.x * GFB_TILE + lx, py = blockIdx.y * GFB_TILE + ly;
    if (px >= W || py >= H) return;
    float c0 = 0.0f, c1 = 0.0f, c2 = 0.0f;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
        const float w = kSsimW[k];
        c0 += w * s_h[0][ly + k][lx];
        c1 += w * s_h[1][ly + k][lx];
        c2 += w * s_h[2][ly + k][lx];
    }
    const size_t pix = (size_t)py * W + px;
    const float m = mask ? (mask[pix] ? 1.0f : 0.0f) : 1.0f;
    const float x = out[ch * P + pix] * m, y = gt_image[pix * 3 + ch] * m;
    ssim_grad[ch * P + pix] = -w_ssim * (c0 + 2.0f * x * c1 + y * c2);
}

// ------------------------------------------------------------------ camera-only stage: moving-subset mask
// trainer.py:446-451: grey = 0.299 r + 0.587 g + 0.114 b of the moving Gaussians' render; grey > 0 removes the
// pixel from the losses from now on (move_mask = move_gs_mask | move_mask).  Also tracks the subset's max K.
__global__ void __launch_bounds__(256)
fit_move_mask_kernel(const float* __restrict__ sub_out, int P, uint8_t* __restrict__ dyn_mask,
                     const int32_t* __restrict__ sub_ctrl, int32_t* __restrict__ status) {
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix == 0) status[ST_SUB_K_MAX] = max(status[ST_SUB_K_MAX], sub_ctrl[GFB_CTRL_K]);
    if (pix >= P) return;
    const float grey = 0.299f * sub_out[pix] + 0.587f * sub_out[(size_t)P + pix] + 0.114f * sub_out[2 * (size_t)P + pix];
    if (grey > 0.0f) dyn_mask[pix] = 0;
}

// ------------------------------------------------------------------ pixel losses -> dL/d(out)
// trainer.py:452-464 (mse over (H,W,3)) and trainer.py:476-488 ((aD+b - Dgt)^2 / (aD+b + Dgt), mean);
// w_rgb = lambda_rgb / (3 H W), w_depth = lambda_depth / (H W).  ssim_grad (3,H,W) is added when present.
__global__ void __launch_bounds__(kThreads)
fit_loss_kernel(const float* __restrict__ out, int C, const float* __restrict__ gt_image,
                const float* __restrict__ gt_depth, const uint8_t* __restrict__ mask,
                const float* __restrict__ depth_ab, int W, int H, float w_rgb, float w_depth, float den_min,
                const float* __restrict__ ssim_grad, float* __restrict__ g_out, float* __restrict__ loss_acc,
                uint32_t* __restrict__ zero_a, size_t n_zero_a, uint32_t* __restrict__ zero_b, size_t n_zero_b,
                size_t zero_b_keep) {
    const size_t P = (size_t)W * H;
    const size_t pix = (size_t)blockIdx.x * kThreads + threadIdx.x;
    // this kernel sits between the forward and the backward of the iteration, so it also clears what the backward
    // accumulates into (zero_a: packed gradients + camera gradients) and what the NEXT forward counts into (zero_b:
    // tile counters + control words; this iteration's binning is over) -- two memset launches less per iteration.
    // Word zero_b_keep (this iteration's K, read by fit_finish later and overwritten by the next scan) is left alone.
    for (size_t k = pix; k < n_zero_a; k += (size_t)gridDim.x * kThreads) zero_a[k] = 0u;
    for (size_t k = pix; k < n_zero_b; k += (size_t)gridDim.x * kThreads)
        if (k != zero_b_keep) zero_b[k] = 0u;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // sum r^2, sum depth term, d depth_a, d depth_b
    if (pix < P) {
        const float m = mask ? (mask[pix] ? 1.0f : 0.0f) : 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = out[c * P + pix] * m, y = gt_image[pix * 3 + c] * m;
            const float r = x - y;
            acc[0] += r * r;
            float g = 2.0f * w_rgb * r;
            if (ssim_grad) g += ssim_grad[c * P + pix];
            g_out[c * P + pix] = g * m;
        }
        if (C > 3) {
            const float a = depth_ab[0], b = depth_ab[1];
            const float D = out[3 * P + pix], gd = gt_depth[pix];
            const float d = a * D + b, e = d - gd, den = d + gd;
            float l, dl;
            if (den_min <= 0.0f || den >= den_min) {  // den_min = 0: the reference's unclamped quotient, whatever its sign
                l = e * e / den;
                dl = (2.0f * e * den - e * e) / (den * den);
            } else {
                l = e * e / den_min;
                dl = 2.0f * e / den_min;
            }
            acc[1] = l * m;
            const float gl = w_depth * dl * m;
            g_out[3 * P + pix] = gl * a;
            acc[2] = gl * D;
            acc[3] = gl;
        }
    }
    block_reduce_atomic<4>(acc, loss_acc + LA_SQ);
}

// ------------------------------------------------------------------ backward + Adam per Gaussian
struct FitMasks {
    const uint8_t* still_mask;
    int n_still, camera_only, freeze_rgb;
};

// two CTAs per SM (112 registers, no spill): this kernel waits on ~30 global loads per thread
__global__ void __launch_bounds__(kThreads, 2)
fit_geometry_bwd_adam_kernel(float* __restrict__ xyz, float* __restrict__ scale_raw, float4* __restrict__ rot_raw,
                             float* __restrict__ op_raw, float* __restrict__ rgb_raw, const float* __restrict__ cam,
                             int N, int W, int H, float nearest, float extent, int C,
                             const float4* __restrict__ grad_pack, FitMasks mk, FitRegs rg,
                             const float* __restrict__ loss_acc, float* __restrict__ adam_m,
                             float* __restrict__ adam_v, AdamStep a, float* __restrict__ d_cam,
                             float* __restrict__ dbg_grads) {
    __shared__ float s_cam[16];
    load_camera(s_cam, cam + 12, cam);  // written by the previous iteration's fit_finish: long complete
    const float* e = s_cam;
    const float* in = s_cam + 12;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const float lambda_var = rg.lambda_var, lambda_scale = rg.lambda_scale;
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0f;
    // The Gaussian's own forward quantities are recomputed BEFORE the wait on blend_bwd (programmatic dependent launch):
    // these CTAs move into the SM slots blend_bwd's last wave leaves idle and are ready when the gradient pack is.
    // Nothing upstream still reads the raw parameters: the blend kernels work from the packed record streams.
    float p[3] = {0.0f, 0.0f, 0.0f}, sr[3] = {0.0f, 0.0f, 0.0f}, s[3] = {0.0f, 0.0f, 0.0f}, c_raw[3] = {0.0f, 0.0f, 0.0f}, S[6];
    float4 qr = make_float4(0.0f, 0.0f, 0.0f, 1.0f), q = qr;
    float o_raw = 0.0f, qn = 1.0f, u = 0.0f, v = 0.0f, xc = 0.0f, yc = 0.0f, zc = 1.0f;
    EwaMid m;
    bool seen = false, live = false;
    if (i < N) {
        p[0] = xyz[3 * i], p[1] = xyz[3 * i + 1], p[2] = xyz[3 * i + 2];
        sr[0] = scale_raw[3 * i], sr[1] = scale_raw[3 * i + 1], sr[2] = scale_raw[3 * i + 2];
        qr = rot_raw[i];
        o_raw = op_raw[i];
        c_raw[0] = rgb_raw[3 * i], c_raw[1] = rgb_raw[3 * i + 1], c_raw[2] = rgb_raw[3 * i + 2];
        s[0] = fabsf(sr[0]), s[1] = fabsf(sr[1]), s[2] = fabsf(sr[2]);
        q = normalize4(qr, qn);
        seen = project_one(in, e, W, H, nearest, extent, p[0], p[1], p[2], u, v, xc, yc, zc);
        if (seen) {
            cov3d_fwd_one(s, q, S);
            ewa_mid_eval(p, S, in, e, W, H, m);
            float rf;
            int x0, y0, x1, y1;
            live = ewa_live(m, u, v, gx, gy, rf, x0, y0, x1, y1);
        }
    }
    gfb_pdl_wait();  // grad_pack comes from blend_bwd (no-op without the PDL attribute)
    if (i < N) {
        const float4 g0 = grad_pack[3 * (size_t)i], g1 = grad_pack[3 * (size_t)i + 1], g2 = grad_pack[3 * (size_t)i + 2];
        float dp[3] = {0.0f, 0.0f, 0.0f}, ds[3] = {0.0f, 0.0f, 0.0f};
        float4 dq = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float gd = (C > 3) ? g2.y : 0.0f;  // dL/d(depth_i): the depth map's feature gradient
        if (seen) {
            if (lambda_scale != 0.0f && u > 0.0f && u < (float)(W - 1) && v > 0.0f && v < (float)(H - 1) &&
                (!rg.scale_sel || rg.scale_sel[i])) {
                const float nrm = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
                const float wgt = lambda_scale / loss_acc[LA_NSCALE];
                if (nrm > 0.0f) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) ds[k] = wgt * s[k] / (nrm * zc);
                }
                gd += -wgt * nrm / (zc * zc);
            }
            if (live) {
                float dS[6], ds2[3];
                ewa_bwd_one(m, p, S, in, e, g0.z, g0.w, g1.x, dp, dS, acc);
                cov3d_bwd_one(s, q, dS, ds2, dq);
#pragma unroll
                for (int k = 0; k < 3; ++k) ds[k] += ds2[k];
            }
            float gu = g0.x, gv = g0.y;
            if (rg.w_flow != 0.0f && i < rg.n_flow && rg.flow_sel[i]) {  // d/duv of w_flow |uv - target|^2
                gu += 2.0f * rg.w_flow * (u - rg.flow_target[2 * i]);
                gv += 2.0f * rg.w_flow * (v - rg.flow_target[2 * i + 1]);
            }
            project_bwd_one(in, e, p[0], p[1], p[2], xc, yc, zc, gu, gv, gd, dp, acc);
        }
        if (rg.w_still != 0.0f && i < rg.n_still_ref && rg.still_sel[i]) {
            const float d0 = p[0] - rg.still_ref[3 * i], d1 = p[1] - rg.still_ref[3 * i + 1], d2 = p[2] - rg.still_ref[3 * i + 2];
            const float nrm = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
            if (nrm > 0.0f) {  // torch.norm's backward is 0 at 0
                dp[0] += rg.w_still * d0 / nrm;
                dp[1] += rg.w_still * d1 / nrm;
                dp[2] += rg.w_still * d2 / nrm;
            }
        }
        if (lambda_var != 0.0f) {
            const float mu = (s[0] + s[1] + s[2]) / 3.0f;
            const float d0 = s[0] - mu, d1 = s[1] - mu, d2 = s[2] - mu;
            const float sd = sqrtf((d0 * d0 + d1 * d1 + d2 * d2) * 0.5f);
            if (sd > 0.0f) {
                const float wgt = lambda_var / ((float)N * 2.0f * sd);
                ds[0] += wgt * d0;
                ds[1] += wgt * d1;
                ds[2] += wgt * d2;
            }
        }
        // activation backward (trainer.py:62-69): abs, normalize, sigmoid(10 x), sigmoid(x)
        float g_s[3], g_c[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) g_s[k] = sr[k] > 0.0f ? ds[k] : (sr[k] < 0.0f ? -ds[k] : 0.0f);
        const float qdot = q.x * dq.x + q.y * dq.y + q.z * dq.z + q.w * dq.w;
        const float4 g_q = make_float4((dq.x - q.x * qdot) / qn, (dq.y - q.y * qdot) / qn, (dq.z - q.z * qdot) / qn,
                                       (dq.w - q.w * qdot) / qn);
        const float so = sigmoid1(10.0f * o_raw);
        const float g_o = g1.y * 10.0f * so * (1.0f - so);
        const float d_c[3] = {g1.z, g1.w, g2.x};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float sc = sigmoid1(c_raw[k]);
            g_c[k] = d_c[k] * sc * (1.0f - sc);
        }
        if (dbg_grads) {
            float* g = dbg_grads + (size_t)i * 14;
            g[0] = dp[0]; g[1] = dp[1]; g[2] = dp[2]; g[3] = g_s[0]; g[4] = g_s[1]; g[5] = g_s[2];
            g[6] = g_q.x; g[7] = g_q.y; g[8] = g_q.z; g[9] = g_q.w; g[10] = g_o; g[11] = g_c[0]; g[12] = g_c[1]; g[13] = g_c[2];
        }
        // gradient masks (trainer.py:535-551).  The reference zeroes the gradient and still calls Adam; with the
        // fresh optimiser state of every train() call and a mask that is constant over the call, a zero gradient
        // gives a zero update, so a masked attribute is simply left alone.
        if (!mk.camera_only) {
            const size_t n3 = 3 * (size_t)N;
            const bool still = mk.still_mask && i < mk.n_still && mk.still_mask[i];
            if (!still) {
                float* m = adam_m + 3 * (size_t)i;
                float* vv = adam_v + 3 * (size_t)i;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    adam_update(p[k], m[k], vv[k], dp[k], a);
                    xyz[3 * i + k] = p[k];
                }
            }
            {
                float* m = adam_m + n3 + 3 * (size_t)i;
                float* vv = adam_v + n3 + 3 * (size_t)i;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    adam_update(sr[k], m[k], vv[k], g_s[k], a);
                    scale_raw[3 * i + k] = sr[k];
                }
            }
            {   // (N,4) block at float offset 6 N: only 8-byte aligned for odd N, so scalar accesses
                float* m = adam_m + 2 * n3 + 4 * (size_t)i;
                float* vv = adam_v + 2 * n3 + 4 * (size_t)i;
                adam_update(qr.x, m[0], vv[0], g_q.x, a);
                adam_update(qr.y, m[1], vv[1], g_q.y, a);
                adam_update(qr.z, m[2], vv[2], g_q.z, a);
                adam_update(qr.w, m[3], vv[3], g_q.w, a);
                rot_raw[i] = qr;
            }
            {
                float* m = adam_m + 2 * n3 + 4 * (size_t)N + i;
                float* vv = adam_v + 2 * n3 + 4 * (size_t)N + i;
                adam_update(o_raw, *m, *vv, g_o, a);
                op_raw[i] = o_raw;
            }
            if (!mk.freeze_rgb) {
                float* m = adam_m + 2 * n3 + 5 * (size_t)N + 3 * (size_t)i;
                float* vv = adam_v + 2 * n3 + 5 * (size_t)N + 3 * (size_t)i;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    adam_update(c_raw[k], m[k], vv[k], g_c[k], a);
                    rgb_raw[3 * i + k] = c_raw[k];
                }
            }
        }
    }
    block_reduce_atomic<16>(acc, d_cam);
}

// ------------------------------------------------------------------ end of iteration (one warp)
struct FitFinish {
    int iter, use_depth, use_ssim, N, P, freeze_camera;
    float lambda_rgb, lambda_depth, lambda_var, lambda_scale, lambda_still, lambda_flow;
    float inv_still_count, inv_flow_count2;  // 1 / still_count, 1 / (2 flow_count); 0 when the term is off
};

__global__ void fit_finish_kernel(float* __restrict__ pose, float* __restrict__ depth_ab, const float* __restrict__ intr,
                                  float* __restrict__ cam, const float* __restrict__ d_cam, float* __restrict__ loss_acc,
                                  float* __restrict__ loss_hist, int32_t* __restrict__ status,
                                  const int32_t* __restrict__ ctrl, float* __restrict__ m_tail,
                                  float* __restrict__ v_tail, AdamStep a_pose, AdamStep a_ab, FitFinish f) {
    if (threadIdx.x != 0) return;
    const float mse = loss_acc[LA_SQ] / (3.0f * (float)f.P);
    const float ssim = f.use_ssim ? loss_acc[LA_SSIM] / (3.0f * (float)f.P) : 0.0f;
    const float ld = f.use_depth ? loss_acc[LA_DEPTH] / (float)f.P : 0.0f;
    const float lv = f.lambda_var != 0.0f ? loss_acc[LA_VAR] / (float)f.N : 0.0f;
    const float ls = (f.lambda_scale != 0.0f && loss_acc[LA_NSCALE] > 0.0f) ? loss_acc[LA_SCALE] / loss_acc[LA_NSCALE] : 0.0f;
    float* h = loss_hist + (size_t)f.iter * HIST_WORDS;
    const float lst = loss_acc[LA_STILL] * f.inv_still_count, lfl = loss_acc[LA_FLOW] * f.inv_flow_count2;
    h[0] = f.lambda_rgb * (mse + (f.use_ssim ? 1.0f - ssim : 0.0f)) + f.lambda_depth * ld + f.lambda_var * lv +
           f.lambda_scale * ls + f.lambda_still * lst + f.lambda_flow * lfl;
    h[1] = mse; h[2] = ssim; h[3] = ld; h[4] = lv; h[5] = ls; h[6] = lst; h[7] = lfl;
    float dp[7];
    extr_grad_to_pose(pose, d_cam, dp);
    float* diag = reinterpret_cast<float*>(status + 8);  // status[8..14]: dL/d(pose) of this iteration (float bits)
#pragma unroll
    for (int k = 0; k < 7; ++k) diag[k] = dp[k];
    if (!f.freeze_camera) {
#pragma unroll
        for (int k = 0; k < 7; ++k) adam_update(pose[k], m_tail[k], v_tail[k], dp[k], a_pose);
    }
    if (f.use_depth && !f.freeze_camera) {
        adam_update(depth_ab[0], m_tail[7], v_tail[7], loss_acc[LA_DA], a_ab);
        adam_update(depth_ab[1], m_tail[8], v_tail[8], loss_acc[LA_DB], a_ab);
    }
    pose_to_extr(pose, cam);
#pragma unroll
    for (int k = 0; k < 4; ++k) cam[12 + k] = intr[k];
#pragma unroll
    for (int k = 0; k < LA_WORDS; ++k) loss_acc[k] = 0.0f;
    const int K = ctrl[GFB_CTRL_K];
    status[ST_ITER] = f.iter + 1;
    status[ST_K_LAST] = K;
    status[ST_K_MAX] = max(status[ST_K_MAX], K);
}

// ------------------------------------------------------------------ workspace layout
struct Layout {
    gfb_fit_layout pub;
    size_t loss_acc, rect, op_act, feat, control, keys, geom, fstream, final_T, n_contrib, grad_ws, ssim_maps, ssim_grad;
};

bool make_layout(int N, int W, int H, int64_t capacity, int max_iters, Layout& L) {
    if (N <= 0 || W <= 0 || H <= 0 || capacity < 0 || max_iters <= 0) return false;
    const size_t P = (size_t)W * H, n = (size_t)N, cap = (size_t)(capacity > 0 ? capacity : 1);
    const size_t T = (size_t)((W + GFB_TILE - 1) / GFB_TILE) * ((H + GFB_TILE - 1) / GFB_TILE);
    const size_t R = (size_t)gfb_tile_replicas((int)T);
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    L.pub.status = take(ST_WORDS * 4);
    L.loss_acc = take(LA_WORDS * 4);
    L.pub.cam = take(16 * 4);
    L.pub.loss_hist = take((size_t)max_iters * HIST_WORDS * 4);
    L.pub.adam_m = take((14 * n + 12) * 4);
    L.pub.adam_v = take((14 * n + 12) * 4);
    L.pub.uv = take(n * 8);
    L.pub.depth = take(n * 4);
    L.pub.conic = take(n * 12);
    L.pub.radius = take(n * 4);
    L.rect = take(n * 8);
    L.op_act = take(n * 4);
    L.feat = take(n * 16);
    L.control = take((2 * T * R + 1 + GFB_CTRL_WORDS) * 4);
    L.pub.tile_range = take(T * 8);
    L.keys = take(cap * 8);
    L.pub.ids = take(cap * 4);
    L.geom = take(cap * 32);
    L.fstream = take(cap * 16);
    L.pub.out = take(4 * P * 4);
    L.final_T = take(P * 4);
    L.n_contrib = take(P * 4);
    L.pub.g_out = take(4 * P * 4);
    L.grad_ws = take((12 * n + 16) * 4);
    L.ssim_maps = take(9 * P * 4);
    L.ssim_grad = take(3 * P * 4);
    L.pub.total = off;
    return true;
}

struct SubLayout {
    size_t uv, depth, conic, radius, rect, op_act, feat, control, tile_range, keys, ids, geom, fstream, out, final_T,
        n_contrib, total;
};

bool make_sub_layout(int n_, int W, int H, int64_t capacity, SubLayout& L) {
    if (n_ <= 0 || W <= 0 || H <= 0 || capacity <= 0) return false;
    const size_t P = (size_t)W * H, n = (size_t)n_, cap = (size_t)capacity;
    const size_t T = (size_t)((W + GFB_TILE - 1) / GFB_TILE) * ((H + GFB_TILE - 1) / GFB_TILE);
    const size_t R = (size_t)gfb_tile_replicas((int)T);
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    L.uv = take(n * 8);
    L.depth = take(n * 4);
    L.conic = take(n * 12);
    L.radius = take(n * 4);
    L.rect = take(n * 8);
    L.op_act = take(n * 4);
    L.feat = take(n * 12);
    L.control = take((2 * T * R + 1 + GFB_CTRL_WORDS) * 4);
    L.tile_range = take(T * 8);
    L.keys = take(cap * 8);
    L.ids = take(cap * 4);
    L.geom = take(cap * 32);
    L.fstream = take(cap * 16);
    L.out = take(3 * P * 4);
    L.final_T = take(P * 4);
    L.n_contrib = take(P * 4);
    L.total = off;
    return true;
}

bool problem_ok(const gfb_fit_problem* p) {
    return p && p->xyz && p->scale && p->rotate && p->opacity && p->rgb && p->pose && p->depth_ab && p->intr &&
           p->gt_image && p->N > 0 && p->W > 0 && p->H > 0 && p->total_iters > 0 && p->n_still >= 0 &&
           (p->n_still == 0 || p->still_mask) && p->n_still <= p->N && p->n_still_ref >= 0 && p->n_still_ref <= p->N &&
           p->n_flow >= 0 && p->n_flow <= p->N && p->still_count >= 0 && p->flow_count >= 0 && p->adam_t0 >= 0 &&
           p->sub_N >= 0 &&
           (p->sub_N == 0 || (p->sub_xyz && p->sub_scale && p->sub_rotate && p->sub_opacity && p->sub_rgb && p->dyn_mask &&
                              p->sub_workspace && p->sub_capacity > 0));
}

AdamStep adam_step(const gfb_fit_problem* p, double lr, int iter) {
    const int total = p->total_iters;
    const double factor =  // LinearLR 1 -> 0.1; constant after a densification re-created the optimiser
        p->constant_lr ? 1.0 : 1.0 + (0.1 - 1.0) * (double)(iter < total ? iter : total) / (double)total;
    const double t = (double)(iter - p->adam_t0) + 1.0;
    const double bc1 = 1.0 - pow((double)p->beta1, t), bc2 = 1.0 - pow((double)p->beta2, t);
    AdamStep a;
    a.step = (float)(lr * factor / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.b1 = p->beta1;
    a.b2 = p->beta2;
    a.eps = p->eps;
    return a;
}

// Programmatic dependent launch between the iteration's kernels, the way pipeline.cu chains its own
// (preprocess -> scatter -> sort -> blend, blend_bwd -> geometry_bwd).  Off until measured on hardware.
bool fit_skip_rgb_grad() {
    static const bool on = [] {
        const char* e = getenv("GFB_FIT_SKIP_RGB_GRAD");
        return !(e && e[0] == '0');
    }();
    return on;
}

// programmatic dependent launch between the kernels of one iteration (GFB_FIT_PDL=0 switches it off): measured
// +7 % iterations/s on a B200 with every parity case green (round 2)
bool fit_pdl() {
    static const bool on = [] {
        const char* e = getenv("GFB_FIT_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

}  // namespace

extern "C" {

int gfb_fit_get_layout(int N, int W, int H, int64_t capacity, int max_iters, gfb_fit_layout* layout) {
    Layout L;
    if (!layout || !make_layout(N, W, H, capacity, max_iters, L)) return GFB_E_BADARG;
    *layout = L.pub;
    return 0;
}

size_t gfb_fit_sub_workspace_bytes(int sub_N, int W, int H, int64_t sub_capacity) {
    SubLayout L;
    return make_sub_layout(sub_N, W, H, sub_capacity, L) ? L.total : 0;
}

int gfb_fit_init(const gfb_fit_problem* p, void* workspace, int64_t capacity, int max_iters, void* stream) {
    Layout L;
    if (!problem_ok(p) || !workspace || !make_layout(p->N, p->W, p->H, capacity, max_iters, L)) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    GFB_TRY(cudaMemsetAsync(ws + L.pub.loss_hist, 0, (size_t)max_iters * HIST_WORDS * 4, st));
    GFB_TRY(cudaMemsetAsync(ws + L.pub.adam_m, 0, (14 * (size_t)p->N + 12) * 4, st));
    GFB_TRY(cudaMemsetAsync(ws + L.pub.adam_v, 0, (14 * (size_t)p->N + 12) * 4, st));
    fit_init_kernel<<<1, 32, 0, st>>>(p->pose, p->intr, (float*)(ws + L.pub.cam), (int32_t*)(ws + L.pub.status),
                                      (float*)(ws + L.loss_acc));
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_fit_iterate(const gfb_fit_problem* p, void* workspace, int64_t capacity, int max_iters, int first_iter,
                    int n_iters, void* stream) {
    Layout L;
    if (!problem_ok(p) || !workspace || !make_layout(p->N, p->W, p->H, capacity, max_iters, L)) return GFB_E_BADARG;
    if (first_iter < 0 || n_iters < 0 || first_iter + n_iters > max_iters || capacity <= 0 || first_iter < p->adam_t0)
        return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const int N = p->N, W = p->W, H = p->H;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE, T = gx * gy;
    const int R = gfb_tile_replicas(T);
    const int P = W * H;
    const bool use_depth = p->gt_depth != nullptr && p->lambda_depth > 0.0f;
    const int C = use_depth ? 4 : 3;
    float* cam = (float*)(ws + L.pub.cam);
    int32_t* status = (int32_t*)(ws + L.pub.status);
    float* loss_acc = (float*)(ws + L.loss_acc);
    float* loss_hist = (float*)(ws + L.pub.loss_hist);
    float* adam_m = (float*)(ws + L.pub.adam_m);
    float* adam_v = (float*)(ws + L.pub.adam_v);
    float* uv = (float*)(ws + L.pub.uv);
    float* depth = (float*)(ws + L.pub.depth);
    float* conic = (float*)(ws + L.pub.conic);
    int32_t* radius = (int32_t*)(ws + L.pub.radius);
    void* rect = ws + L.rect;
    float* op_act = (float*)(ws + L.op_act);
    float* feat = (float*)(ws + L.feat);
    int32_t* counts = (int32_t*)(ws + L.control);
    int32_t* ctrl = counts + (size_t)T * R;
    int32_t* offsets = ctrl + GFB_CTRL_WORDS;
    int32_t* tile_range = (int32_t*)(ws + L.pub.tile_range);
    void* keys = ws + L.keys;
    int32_t* ids = (int32_t*)(ws + L.pub.ids);
    void* geom = ws + L.geom;
    void* fstream = ws + L.fstream;
    float* out = (float*)(ws + L.pub.out);
    float* final_T = (float*)(ws + L.final_T);
    int32_t* n_contrib = (int32_t*)(ws + L.n_contrib);
    float* g_out = (float*)(ws + L.pub.g_out);
    float* grad_ws = (float*)(ws + L.grad_ws);
    float* d_cam = grad_ws + (size_t)N * 12;
    float* ssim_maps = (float*)(ws + L.ssim_maps);
    float* ssim_grad = (float*)(ws + L.ssim_grad);
    const float w_rgb = p->lambda_rgb / (3.0f * (float)P), w_depth = p->lambda_depth / (float)P;
    const FitMasks mk{p->still_mask, p->n_still, p->camera_only, p->freeze_rgb};
    const bool use_still = p->lambda_still != 0.0f && p->still_ref && p->still_sel && p->still_count > 0;
    const bool use_flow = p->lambda_flow != 0.0f && p->flow_target && p->flow_sel && p->flow_count > 0;
    const float inv_still = use_still ? 1.0f / (float)p->still_count : 0.0f;
    const float inv_flow2 = use_flow ? 1.0f / (2.0f * (float)p->flow_count) : 0.0f;
    const FitRegs rg{p->lambda_var,  p->lambda_scale, p->scale_sel, p->still_ref,   p->still_sel, use_still ? p->n_still_ref : 0,
                     p->lambda_still * inv_still, p->flow_target, p->flow_sel,  use_flow ? p->n_flow : 0,
                     p->lambda_flow * inv_flow2};
    const int nblk = gfb_div_up(N, kThreads);
    const bool pdl = fit_pdl();
    const int tight = gfb_tight_tiles() ? 1 : 0;
    // frames >= 1 freeze the colours (trainer.py:537-540) and a camera-only stage freezes every attribute: the blend
    // backward then does not reduce the rgb channels' own gradient (GFB_FIT_SKIP_RGB_GRAD=0 keeps the full backward)
    const bool no_rgb = C == 4 && (p->freeze_rgb || p->camera_only) && fit_skip_rgb_grad();
    // camera-only stage: the moving subset is rendered every iteration and its footprint leaves the losses
    const bool use_sub = p->sub_N > 0 && p->camera_only;
    SubLayout S;
    if (use_sub && !make_sub_layout(p->sub_N, W, H, p->sub_capacity, S)) return GFB_E_BADARG;
    char* sw = (char*)p->sub_workspace;
    const uint8_t* loss_mask = use_sub ? p->dyn_mask : p->pixel_mask;
    const FitRegs no_regs{0.0f, 0.0f, nullptr, nullptr, nullptr, 0, 0.0f, nullptr, nullptr, 0, 0.0f};
    int rc;
    const size_t n_ctrl_words = (size_t)T * R + GFB_CTRL_WORDS, n_grad_words = (size_t)N * 12 + 16;
    // the tile counters are cleared by fit_loss of the previous iteration; before the first one of a call, here
    GFB_TRY(cudaMemsetAsync(counts, 0, n_ctrl_words * sizeof(int32_t), st));
    for (int it = first_iter; it < first_iter + n_iters; ++it) {
        const GfbRange nvtx_range("gfb_fit_iteration");
        fit_preprocess_kernel<<<nblk, kThreads, 0, st>>>(
            p->xyz, p->scale, reinterpret_cast<const float4*>(p->rotate), p->opacity, p->rgb, cam, N, W, H, p->nearest,
            p->extent, C, reinterpret_cast<float2*>(uv), depth, conic, radius, reinterpret_cast<ushort4*>(rect), op_act,
            feat, counts, offsets, ctrl, T, R, rg, loss_acc, p->dbg_act, tight);
        GFB_CHECK_LAUNCH();
        rc = gfb_internal_scatter_sort_pack_blend(rect, depth, N, W, H, counts, capacity, keys, tile_range, uv, conic, op_act,
                                                  feat, C, ids, geom, fstream, p->bg, out, final_T, n_contrib, stream, pdl);
        if (rc) return rc;
        if (use_sub) {
            int32_t* s_counts = (int32_t*)(sw + S.control);
            int32_t* s_ctrl = s_counts + (size_t)T * R;
            GFB_TRY(cudaMemsetAsync(s_counts, 0, ((size_t)T * R + GFB_CTRL_WORDS) * sizeof(int32_t), st));
            fit_preprocess_kernel<<<gfb_div_up(p->sub_N, kThreads), kThreads, 0, st>>>(
                p->sub_xyz, p->sub_scale, reinterpret_cast<const float4*>(p->sub_rotate), p->sub_opacity, p->sub_rgb, cam,
                p->sub_N, W, H, p->nearest, p->extent, 3, reinterpret_cast<float2*>(sw + S.uv), (float*)(sw + S.depth),
                (float*)(sw + S.conic), (int32_t*)(sw + S.radius), reinterpret_cast<ushort4*>(sw + S.rect),
                (float*)(sw + S.op_act), (float*)(sw + S.feat), s_counts, s_ctrl + GFB_CTRL_WORDS, s_ctrl, T, R, no_regs,
                loss_acc, nullptr, tight);
            GFB_CHECK_LAUNCH();
            rc = gfb_internal_scatter_sort_pack_blend(sw + S.rect, (float*)(sw + S.depth), p->sub_N, W, H, s_counts,
                                                      p->sub_capacity, sw + S.keys, (int32_t*)(sw + S.tile_range),
                                                      (float*)(sw + S.uv), (float*)(sw + S.conic), (float*)(sw + S.op_act),
                                                      (float*)(sw + S.feat), 3, (int32_t*)(sw + S.ids), sw + S.geom,
                                                      sw + S.fstream, p->bg, (float*)(sw + S.out), (float*)(sw + S.final_T),
                                                      (int32_t*)(sw + S.n_contrib), stream, pdl);
            if (rc) return rc;
            fit_move_mask_kernel<<<gfb_div_up(P, 256), 256, 0, st>>>((float*)(sw + S.out), P, p->dyn_mask, s_ctrl, status);
            GFB_CHECK_LAUNCH();
        }
        if (p->use_ssim) {
            ssim_stats_kernel<<<dim3(gx, gy, 3), 256, 0, st>>>(out, p->gt_image, loss_mask, W, H, ssim_maps, loss_acc);
            GFB_CHECK_LAUNCH();
            ssim_grad_kernel<<<dim3(gx, gy, 3), 256, 0, st>>>(out, p->gt_image, loss_mask, W, H, ssim_maps, w_rgb,
                                                               ssim_grad);
            GFB_CHECK_LAUNCH();
        }
        fit_loss_kernel<<<gfb_div_up(P, kThreads), kThreads, 0, st>>>(out, C, p->gt_image, p->gt_depth, loss_mask,
                                                                     p->depth_ab, W, H, w_rgb, w_depth, p->depth_den_min,
                                                                     p->use_ssim ? ssim_grad : nullptr, g_out, loss_acc,
                                                                     reinterpret_cast<uint32_t*>(grad_ws), n_grad_words,
                                                                     reinterpret_cast<uint32_t*>(counts), n_ctrl_words,
                                                                     (size_t)T * R + GFB_CTRL_K);
        GFB_CHECK_LAUNCH();
        rc = gfb_internal_blend_bwd(geom, fstream, capacity, ids, tile_range, C, 0, C, p->bg, W, H, final_T, n_contrib,
                                    g_out, grad_ws, stream, no_rgb, nullptr);
        if (rc) return rc;
        GFB_TRY(gfb_launch_pdl(fit_geometry_bwd_adam_kernel, dim3(nblk), dim3(kThreads), st, pdl, p->xyz, p->scale,
                               reinterpret_cast<float4*>(p->rotate), p->opacity, p->rgb, cam, N, W, H, p->nearest, p->extent,
                               C, reinterpret_cast<const float4*>(grad_ws), mk, rg, loss_acc, adam_m, adam_v,
                               adam_step(p, p->lr, it), d_cam, p->dbg_grads));
        GFB_CHECK_LAUNCH();
        const FitFinish ff{it, use_depth ? 1 : 0, p->use_ssim, N, P, p->freeze_camera, p->lambda_rgb, p->lambda_depth, p->lambda_var,
                           p->lambda_scale, use_still ? p->lambda_still : 0.0f, use_flow ? p->lambda_flow : 0.0f,
                           inv_still, inv_flow2};
        fit_finish_kernel<<<1, 32, 0, st>>>(p->pose, p->depth_ab, p->intr, cam, d_cam, loss_acc, loss_hist, status, ctrl,
                                            adam_m + 14 * (size_t)N, adam_v + 14 * (size_t)N,
                                            adam_step(p, p->lr_camera, it), adam_step(p, p->lr, it), ff);
        GFB_CHECK_LAUNCH();
    }
    return 0;
}

}  // extern "C"
