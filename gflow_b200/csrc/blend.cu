// blend.cu -- msplat.alpha_blending forward / backward for sm_100a.
//
// Reference interface: msplat.alpha_blending(uv, conic, opacity, feature,
// gaussian_ids_sorted, tile_range, bg, W, H) -> (C,H,W), called from
// /root/reference/gflow/utils/render.py:58-64,68-74,84-90,99-105,148-154; its backward is
// triggered by loss.backward() at /root/reference/gflow/trainer.py:533.
//
// Design (B200-first, not the 3DGS renderCUDA structure):
//  * A pack pass gathers the per-Gaussian parameters once per sorted intersection into
//    three tile-contiguous streams of 16-byte records:
//        A = {u, v, hx, hy}   centre + half extents of the alpha >= 1/255 ellipse's bbox
//        B = {a, b, c, o}     conic + opacity
//        F = {f0, f1, f2, f3} up to four feature channels
//    A/B are shared by every blend with the same geometry (rgb / depth / depth-colour in
//    render_multiple) and by the backward pass.
//  * One CTA per 16x16 tile.  Batches of 128 records are staged into shared memory by
//    1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), double buffered, issued by
//    one thread; no thread does scattered global gathers inside the blend loop.
//  * Each warp owns an 8x4 pixel block.  Per 32 records, the lanes test the records' bboxes
//    against the warp's block in parallel (one LDS.128 each), ballot, and the warp then walks
//    only the surviving records.  With sigma ~ 1 px splats this removes ~4/5 of the
//    (pixel, Gaussian) evaluations of a 256-pixel-per-Gaussian tile walk; results are
//    unchanged because a culled pair has alpha < 1/255 and would have been skipped.
//  * Backward: same staging back to front; the per-lane gradients of a (warp, Gaussian) pair
//    are reduced with a transposed butterfly (9 shuffles for 8 values instead of 40) and
//    added with one RED per value into a packed 48-byte-per-Gaussian gradient record, so
//    the (up to) ten atomics of a warp hit one or two L2 sectors.
//  No tensor cores: this is a gather / scatter bounded by issue rate and L2 atomics.
#include <cstdlib>

#include "splat_math.cuh"

namespace {

constexpr int kBatch = 128;          // records per TMA stage
constexpr int kBlendThreads = 256;   // up to 8 warps per CTA, each an 8x4 pixel block
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk TMA copy global -> shared, completion signalled on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// try_wait suspends the thread for a hardware time slice per attempt; a transfer that never
// completes (bad address / size) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    uint32_t spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// -0.5 (a dx^2 + c dy^2) - b dx dy with explicit roundings, so the forward and backward
// kernels take identical skip decisions on every (pixel, Gaussian) pair.
__device__ __forceinline__ float splat_power(float a, float b, float c, float dx, float dy) {
    float q = __fmul_rn(__fmul_rn(a, dx), dx);
    q = __fmaf_rn(__fmul_rn(c, dy), dy, q);
    const float r = __fmul_rn(__fmul_rn(b, dx), dy);
    return __fmaf_rn(-0.5f, q, -r);
}

// exp(power) as one FMUL + MUFU.EX2 (ex2.approx.ftz): power is in [-5.6, 0] wherever alpha can
// reach 1/255, far from the denormal range __expf() guards against with three extra instructions.
// Forward and backward share it, so they agree on every alpha >= 1/255 decision.
__device__ __forceinline__ float splat_exp(float power) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(power * 1.4426950408889634f));
    return r;
}
// 1 / x for x = 1 - alpha in [0.01, 1]: one MUFU.RCP, no range fix-up
__device__ __forceinline__ float splat_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// ------------------------------------------------------------------ pack kernels
__global__ void __launch_bounds__(256)
pack_geometry_kernel(const float2* __restrict__ uv, const float* __restrict__ conic,
                     const float* __restrict__ opacity, const int32_t* __restrict__ ids, long long K,
                     float4* __restrict__ sA, float4* __restrict__ sB) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int id = ids[k];
    const float2 p = uv[id];
    const float a = conic[3 * id], b = conic[3 * id + 1], c = conic[3 * id + 2];
    const float o = opacity[id];
    float hx, hy;
    gfbm::splat_bbox(a, b, c, o, hx, hy);
    sA[k] = make_float4(p.x, p.y, hx, hy);
    sB[k] = make_float4(a, b, c, o);
}

__global__ void __launch_bounds__(256)
pack_feature_kernel(const float* __restrict__ feature, int C, int c0, int Cg, const int32_t* __restrict__ ids,
                    long long K, float4* __restrict__ sF) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const float* f = feature + (size_t)ids[k] * C + c0;
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    r.x = f[0];
    if (Cg > 1) r.y = f[1];
    if (Cg > 2) r.z = f[2];
    if (Cg > 3) r.w = f[3];
    sF[k] = r;
}

// ------------------------------------------------------------------ staging
struct Stage {
    float4 A[kBatch];
    float4 B[kBatch];
    float4 F[kBatch];
};

__device__ __forceinline__ void issue_batch(Stage* st, uint64_t* bar, const float4* __restrict__ gA,
                                            const float4* __restrict__ gB, const float4* __restrict__ gF,
                                            long long first, int cnt) {
    const uint32_t bytes = (uint32_t)cnt * 16u;
    mbar_expect_tx(bar, 3u * bytes);
    bulk_g2s(st->A, gA + first, bytes, bar);
    bulk_g2s(st->B, gB + first, bytes, bar);
    bulk_g2s(st->F, gF + first, bytes, bar);
}

__device__ __forceinline__ float f4_get(const float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// ------------------------------------------------------------------ forward
template <int CG>
__global__ void __launch_bounds__(kBlendThreads)
blend_fwd_kernel(const float4* __restrict__ gA, const float4* __restrict__ gB, const float4* __restrict__ gF,
                 const int2* __restrict__ tile_range, int gx, int c0, float bg, int W, int H,
                 float* __restrict__ out, float* __restrict__ final_T, int32_t* __restrict__ n_contrib) {
    __shared__ __align__(128) Stage s_stage[2];
    __shared__ __align__(8) uint64_t s_bar[2];

    gfb_pdl_wait();  // fused pipeline: tile_sort_pack may still be draining
    // a CTA of `wpc` warps covers wpc of the tile's eight 8x4 pixel blocks (8 / wpc CTAs per tile)
    const int wpc = blockDim.x >> 5, per_tile = 8 / wpc;
    const int tile = blockIdx.x / per_tile;
    const int tx = tile % gx, ty = tile / gx;
    const int2 range = tile_range[tile];
    const int n = range.y - range.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (blockIdx.x % per_tile) * wpc + (tid >> 5);  // block index inside the tile
    const int bx0 = tx * GFB_TILE + (warp & 1) * 8, by0 = ty * GFB_TILE + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = (px < W) && (py < H);
    const float fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 3);
    const float pxf = (float)px, pyf = (float)py;

    float T = 1.0f;
    float acc[CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) acc[c] = 0.0f;
    int last = 0;
    bool done = !inside;

    if (n > 0) {
        const int nb = (n + kBatch - 1) / kBatch;
        if (tid == 0) {
            mbar_init(&s_bar[0], 1);
            mbar_init(&s_bar[1], 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) issue_batch(&s_stage[0], &s_bar[0], gA, gB, gF, range.x, min(kBatch, n));
        for (int b = 0; b < nb; ++b) {
            const int s = b & 1;
            const Stage& st = s_stage[s];
            mbar_wait(&s_bar[s], (uint32_t)(b >> 1) & 1u);
            // stage s^1 was last read in iteration b-1, which ended with a CTA barrier
            if (tid == 0 && b + 1 < nb)
                issue_batch(&s_stage[s ^ 1], &s_bar[s ^ 1], gA, gB, gF, (long long)range.x + (long long)(b + 1) * kBatch,
                            min(kBatch, n - (b + 1) * kBatch));
            const int cnt = min(kBatch, n - b * kBatch);
            if (!__all_sync(kFull, done)) {
                for (int base = 0; base < cnt; base += 32) {
                    const int j = base + lane;
                    bool hit = false;
                    if (j < cnt) {
                        const float4 a4 = st.A[j];
                        hit = (a4.x + a4.z >= fx0) && (a4.x - a4.z <= fx1) && (a4.y + a4.w >= fy0) &&
                              (a4.y - a4.w <= fy1);
                    }
                    unsigned m = __ballot_sync(kFull, hit);
                    while (m) {
                        const int jj = base + __ffs(m) - 1;
                        m &= m - 1;
                        if (!done) {
                            const float4 a4 = st.A[jj];
                            const float4 b4 = st.B[jj];
                            const float dx = a4.x - pxf, dy = a4.y - pyf;
                            const float power = splat_power(b4.x, b4.y, b4.z, dx, dy);
                            if (power <= 0.0f) {
                                const float alpha = fminf(GFB_ALPHA_MAX, b4.w * splat_exp(power));
                                if (alpha >= GFB_ALPHA_MIN) {
                                    const float test_T = T * (1.0f - alpha);
                                    if (test_T < GFB_T_EPS) {
                                        done = true;
                                    } else {
                                        const float4 f4 = st.F[jj];
                                        const float w = alpha * T;
#pragma unroll
                                        for (int c = 0; c < CG; ++c) acc[c] = fmaf(f4_get(f4, c), w, acc[c]);
                                        T = test_T;
                                        last = b * kBatch + jj + 1;
                                    }
                                }
                            }
                        }
                    }
                    if (__all_sync(kFull, done)) break;
                }
            }
            if (__syncthreads_count(done) == (int)blockDim.x) {
                // drain the copy already in flight before the CTA (and its shared memory) retires
                if (tid == 0 && b + 1 < nb) mbar_wait(&s_bar[s ^ 1], (uint32_t)((b + 1) >> 1) & 1u);
                break;
            }
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px;
        const size_t HW = (size_t)W * H;
#pragma unroll
        for (int c = 0; c < CG; ++c) out[(size_t)(c0 + c) * HW + pix] = fmaf(T, bg, acc[c]);
        final_T[pix] = T;
        n_contrib[pix] = last;
    }
}

// ------------------------------------------------------------------ backward
// Transposed butterfly: every lane contributes v[0..7]; afterwards lane 4k holds the warp
// total of slot k (k = 0..7).  4 + 2 + 1 + 1 + 1 = 9 shuffles.
__device__ __forceinline__ float warp_reduce8(const float (&v)[8], int lane) {
    float r4[4], r2[2], r1;
    const bool up16 = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = up16 ? v[i] : v[i + 4];
        const float keep = up16 ? v[i + 4] : v[i];
        r4[i] = keep + __shfl_xor_sync(kFull, send, 16);
    }
    const bool up8 = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = up8 ? r4[i] : r4[i + 2];
        const float keep = up8 ? r4[i + 2] : r4[i];
        r2[i] = keep + __shfl_xor_sync(kFull, send, 8);
    }
    const bool up4 = (lane & 4) != 0;
    {
        const float send = up4 ? r2[0] : r2[1];
        const float keep = up4 ? r2[1] : r2[0];
        r1 = keep + __shfl_xor_sync(kFull, send, 4);
    }
    r1 += __shfl_xor_sync(kFull, r1, 2);
    r1 += __shfl_xor_sync(kFull, r1, 1);
    return r1;  // slot ((lane>>4)&1)*4 + ((lane>>3)&1)*2 + ((lane>>2)&1)
}

// SPARSE (experimental, off by default, GFB_BWD_SPARSE=k): a (warp, record) pair with at most k active lanes skips
// the butterfly and lets every active lane add its own values (k x 10 reductions in L2 instead of ~45
// shuffle / select / add instructions).  With the ~1 px splats GFlow fits, 40 % of the pairs have <= 4 active lanes.
// NO_RGB (CG == 4 only; the native fit loop on frames >= 1, where GFlow zeroes the colour gradient, trainer.py:537-540):
// channels 0..2 still enter dalpha through f . g, but their own gradient is neither reduced nor added -- the depth
// channel's gradient rides in the butterfly's seventh slot, so one reduce8 replaces reduce8 + two warp sums.
template <int CG, bool SPARSE, bool NO_RGB = false>
__global__ void __launch_bounds__(kBlendThreads)
blend_bwd_kernel(const float4* __restrict__ gA, const float4* __restrict__ gB, const float4* __restrict__ gF,
                 const int32_t* __restrict__ ids, const int2* __restrict__ tile_range, int gx, int c0, float bg,
                 int W, int H, const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
                 const float* __restrict__ g_out, float* __restrict__ grad_pack, int sparse_lanes) {
    __shared__ __align__(128) Stage s_stage[2];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_max_last;

    gfb_pdl_launch_dependents();  // fused pipeline: geometry_bwd may queue behind the last wave
    const int wpc = blockDim.x >> 5, per_tile = 8 / wpc;
    const int tile = blockIdx.x / per_tile;
    const int tx = tile % gx, ty = tile / gx;
    const int2 range = tile_range[tile];
    const int n = range.y - range.x;
    if (n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = (blockIdx.x % per_tile) * wpc + (tid >> 5);
    const int bx0 = tx * GFB_TILE + (warp & 1) * 8, by0 = ty * GFB_TILE + (warp >> 1) * 4;
    const int px = bx0 + (lane & 7), py = by0 + (lane >> 3);
    const bool inside = (px < W) && (py < H);
    const float fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 3);
    const float pxf = (float)px, pyf = (float)py;

    float Tf = 1.0f;
    int last = 0;
    float go[CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) go[c] = 0.0f;
    if (inside) {
        const size_t pix = (size_t)py * W + px;
        const size_t HW = (size_t)W * H;
        Tf = final_T[pix];
        last = n_contrib[pix];
#pragma unroll
        for (int c = 0; c < CG; ++c) go[c] = g_out[(size_t)(c0 + c) * HW + pix];
    }
    float bgdot = 0.0f;
#pragma unroll
    for (int c = 0; c < CG; ++c) bgdot = fmaf(bg, go[c], bgdot);

    const int wmax = __reduce_max_sync(kFull, last);
    if (tid == 0) {
        s_max_last = 0;
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (lane == 0 && wmax > 0) atomicMax(&s_max_last, wmax);
    __syncthreads();
    const int max_last = min(s_max_last, n);
    if (max_last <= 0) return;

    const int nb = (max_last + kBatch - 1) / kBatch;  // batches [0, nb) hold every contributor
    // iteration it processes batch b = nb-1-it
    if (tid == 0)
        issue_batch(&s_stage[0], &s_bar[0], gA, gB, gF, (long long)range.x + (long long)(nb - 1) * kBatch,
                    max_last - (nb - 1) * kBatch);

    float T = Tf;
    float S = Tf * bgdot;

    for (int it = 0; it < nb; ++it) {
        const int b = nb - 1 - it;
        const int s = it & 1;
        const Stage& st = s_stage[s];
        mbar_wait(&s_bar[s], (uint32_t)(it >> 1) & 1u);
        if (tid == 0 && it + 1 < nb)
            issue_batch(&s_stage[s ^ 1], &s_bar[s ^ 1], gA, gB, gF, (long long)range.x + (long long)(b - 1) * kBatch,
                        kBatch);
        const int cnt = min(kBatch, max_last - b * kBatch);
        const int pos0 = b * kBatch;
        if (pos0 < wmax) {
            for (int base = ((cnt - 1) >> 5) << 5; base >= 0; base -= 32) {
                const int j = base + lane;
                bool hit = false;
                if (j < cnt && pos0 + j < wmax) {
                    const float4 a4 = st.A[j];
                    hit = (a4.x + a4.z >= fx0) && (a4.x - a4.z <= fx1) && (a4.y + a4.w >= fy0) &&
                          (a4.y - a4.w <= fy1);
                }
                unsigned m = __ballot_sync(kFull, hit);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
                    const int jj = base + bit;
                    const int pos = pos0 + jj;
                    float v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
                    float v8 = 0.0f, v9 = 0.0f;
                    bool act = false;
                    if (pos < last) {
                        const float4 a4 = st.A[jj];
                        const float4 b4 = st.B[jj];
                        const float dx = a4.x - pxf, dy = a4.y - pyf;
                        const float power = splat_power(b4.x, b4.y, b4.z, dx, dy);
                        if (power <= 0.0f) {
                            const float G = splat_exp(power);
                            const float alpha = fminf(GFB_ALPHA_MAX, b4.w * G);
                            if (alpha >= GFB_ALPHA_MIN) {
                                act = true;
                                const float4 f4 = st.F[jj];
                                const float inv1ma = splat_rcp(1.0f - alpha);
                                T = T * inv1ma;  // transmittance in front of this Gaussian
                                const float w = alpha * T;
                                // S = sum_{k behind j} w_k (f_k . g) + T_final (bg . g): one scalar carries what
                                // the 3DGS formulation keeps as accum_rec[C] / last_color[C] / last_alpha
                                float fg = 0.0f;
                                float df[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                                for (int c = 0; c < CG; ++c) {
                                    fg = fmaf(f4_get(f4, c), go[c], fg);
                                    df[c] = w * go[c];
                                }
                                const float dalpha = fmaf(T, fg, -S * inv1ma);
                                S = fmaf(w, fg, S);
                                const float dG = b4.w * dalpha;
                                const float gdx = G * dx, gdy = G * dy;
                                v[0] = dG * (-gdx * b4.x - gdy * b4.y);
                                v[1] = dG * (-gdy * b4.z - gdx * b4.y);
                                v[2] = -0.5f * gdx * dx * dG;
                                v[3] = -gdx * dy * dG;
                                v[4] = -0.5f * gdy * dy * dG;
                                v[5] = G * dalpha;
                                if constexpr (NO_RGB) {
                                    v[6] = df[3];
                                } else {
                                    v[6] = df[0];
                                    if (CG > 1) v[7] = df[1];
                                    if (CG > 2) v8 = df[2];
                                    if (CG > 3) v9 = df[3];
                                }
                            }
                        }
                    }
                    if constexpr (SPARSE) {
                        const unsigned actm = __ballot_sync(kFull, act);
                        if (actm == 0u) continue;
                        if (__popc(actm) <= sparse_lanes) {
                            if (act) {
                                float* gp = grad_pack + (size_t)ids[(long long)range.x + pos] * 12;
#pragma unroll
                                for (int i = 0; i < 6; ++i) atomicAdd(gp + i, v[i]);
                                if constexpr (NO_RGB) {
                                    atomicAdd(gp + 9, v[6]);
                                } else {
                                    atomicAdd(gp + 6, v[6]);
                                    if (CG > 1) atomicAdd(gp + 7, v[7]);
                                    if (CG > 2) atomicAdd(gp + 8, v8);
                                    if (CG > 3) atomicAdd(gp + 9, v9);
                                }
                            }
                            continue;
                        }
                    } else {
                        if (!__any_sync(kFull, act)) continue;
                    }
                    const float r = warp_reduce8(v, lane);
                    if constexpr (NO_RGB) {  // slots 0..5 -> gp[0..5], slot 6 (depth channel) -> gp[9]
                        const int slot = lane >> 2;
                        if ((lane & 3) == 0 && slot < 7)
                            atomicAdd(grad_pack + (size_t)ids[(long long)range.x + pos] * 12 + (slot < 6 ? slot : 9), r);
                        continue;
                    }
                    if (CG > 2) v8 = gfb_warp_sum(v8);
                    if (CG > 3) v9 = gfb_warp_sum(v9);
                    const int id = ids[(long long)range.x + pos];
                    float* gp = grad_pack + (size_t)id * 12;
                    // lanes 0,4,..,28 hold slots 0..7; lanes 1 / 2 carry slots 8 / 9: one RED instruction
                    constexpr int kSlots8 = 6 + (CG > 1 ? 2 : 1);
                    const bool lead = (lane & 3) == 0;
                    const int slot = lead ? (lane >> 2) : (7 + lane);
                    const float val = lead ? r : (lane == 1 ? v8 : v9);
                    if (lead ? (slot < kSlots8) : ((lane == 1 && CG > 2) || (lane == 2 && CG > 3))) atomicAdd(gp + slot, val);
                }
            }
        }
        __syncthreads();  // every warp is done with stage s before it is refilled (it+2)
    }
}

__global__ void __launch_bounds__(256)
unpack_grads_kernel(const float4* __restrict__ grad_pack, int N, int C, int c0, int Cg, float2* __restrict__ d_uv,
                    float* __restrict__ d_conic, float* __restrict__ d_opacity, float* __restrict__ d_feature,
                    int accumulate) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float4 g0 = grad_pack[3 * (size_t)i], g1 = grad_pack[3 * (size_t)i + 1], g2 = grad_pack[3 * (size_t)i + 2];
    if (accumulate) {
        float2 p = d_uv[i];
        p.x += g0.x;
        p.y += g0.y;
        d_uv[i] = p;
        d_conic[3 * i] += g0.z;
        d_conic[3 * i + 1] += g0.w;
        d_conic[3 * i + 2] += g1.x;
        d_opacity[i] += g1.y;
    } else {
        d_uv[i] = make_float2(g0.x, g0.y);
        d_conic[3 * i] = g0.z;
        d_conic[3 * i + 1] = g0.w;
        d_conic[3 * i + 2] = g1.x;
        d_opacity[i] = g1.y;
    }
    float* f = d_feature + (size_t)i * C + c0;
    f[0] = g1.z;
    if (Cg > 1) f[1] = g1.w;
    if (Cg > 2) f[2] = g2.x;
    if (Cg > 3) f[3] = g2.y;
}

}  // namespace

// warps per CTA for the blend kernels (2, 4 or 8): fewer warps per CTA = less waiting at the per-batch
// CTA barrier when the tile's pixel blocks see different numbers of splats, at the price of staging
// the tile's records 8 / wpc times.  GFB_BLEND_WPC overrides the default.
static int blend_warps_per_cta() {
    static int wpc = [] {
        const char* e = getenv("GFB_BLEND_WPC");
        const int v = e ? atoi(e) : 8;
        return (v == 2 || v == 4 || v == 8) ? v : 8;
    }();
    return wpc;
}

// GFB_BWD_SPARSE=k (1..8): experimental direct-reduction path of the backward for pairs with <= k active lanes
static int blend_bwd_sparse_lanes() {
    static int k = [] {
        const char* e = getenv("GFB_BWD_SPARSE");
        const int v = e ? atoi(e) : 0;
        return (v >= 1 && v <= 8) ? v : 0;
    }();
    return k;
}

// ====================================================================== C ABI
extern "C" {

size_t gfb_blend_geometry_stream_bytes(int64_t K) { return K < 0 ? 0 : (size_t)K * 32; }
size_t gfb_blend_feature_stream_bytes(int64_t K) { return K < 0 ? 0 : (size_t)K * 16; }
size_t gfb_blend_grad_pack_bytes(int N) { return N < 0 ? 0 : (size_t)N * 48; }

int gfb_blend_pack_geometry(const float* uv, const float* conic, const float* opacity,
                            const int32_t* gaussian_ids_sorted, int64_t K, void* geom_stream, void* stream) {
    if (K < 0) return GFB_E_BADARG;
    if (K == 0) return 0;
    if (!uv || !conic || !opacity || !gaussian_ids_sorted || !geom_stream) return GFB_E_BADARG;
    float4* sA = reinterpret_cast<float4*>(geom_stream);
    float4* sB = sA + K;
    pack_geometry_kernel<<<gfb_div_up(K, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(uv), conic, opacity, gaussian_ids_sorted, (long long)K, sA, sB);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_blend_pack_feature(const float* feature, int C, int c0, int Cg, const int32_t* gaussian_ids_sorted, int64_t K,
                           void* feat_stream, void* stream) {
    if (K < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (K == 0) return 0;
    if (!feature || !gaussian_ids_sorted || !feat_stream) return GFB_E_BADARG;
    pack_feature_kernel<<<gfb_div_up(K, 256), 256, 0, (cudaStream_t)stream>>>(
        feature, C, c0, Cg, gaussian_ids_sorted, (long long)K, reinterpret_cast<float4*>(feat_stream));
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"

int gfb_internal_blend_fwd(const void* geom_stream, const void* feat_stream, int64_t K, const int32_t* tile_range,
                           int C, int c0, int Cg, float bg, int W, int H, float* out, float* final_T,
                           int32_t* n_contrib, void* stream, bool pdl) {
    if (W <= 0 || H <= 0 || K < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (!tile_range || !out || !final_T || !n_contrib) return GFB_E_BADARG;
    if (K > 0 && (!geom_stream || !feat_stream)) return GFB_E_BADARG;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const float4* gA = reinterpret_cast<const float4*>(geom_stream);
    const float4* gB = gA + K;
    const float4* gF = reinterpret_cast<const float4*>(feat_stream);
    const int2* tr = reinterpret_cast<const int2*>(tile_range);
    cudaStream_t st = (cudaStream_t)stream;
    const int wpc = blend_warps_per_cta();
    const dim3 grid(gx * gy * (8 / wpc)), block(32 * wpc);
    cudaError_t le;
    switch (Cg) {
        case 1: le = gfb_launch_pdl(blend_fwd_kernel<1>, grid, block, st, pdl, gA, gB, gF, tr, gx, c0, bg, W, H, out, final_T, n_contrib); break;
        case 2: le = gfb_launch_pdl(blend_fwd_kernel<2>, grid, block, st, pdl, gA, gB, gF, tr, gx, c0, bg, W, H, out, final_T, n_contrib); break;
        case 3: le = gfb_launch_pdl(blend_fwd_kernel<3>, grid, block, st, pdl, gA, gB, gF, tr, gx, c0, bg, W, H, out, final_T, n_contrib); break;
        default: le = gfb_launch_pdl(blend_fwd_kernel<4>, grid, block, st, pdl, gA, gB, gF, tr, gx, c0, bg, W, H, out, final_T, n_contrib); break;
    }
    if (le != cudaSuccess) return (int)le;
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_internal_blend_bwd(const void*, const void*, int64_t, const int32_t*, const int32_t*, int, int, int, float, int, int,
                           const float*, const int32_t*, const float*, float*, void*, bool no_rgb);

extern "C" {

int gfb_alpha_blending_fwd(const void* geom_stream, const void* feat_stream, int64_t K, const int32_t* tile_range,
                           int C, int c0, int Cg, float bg, int W, int H, float* out, float* final_T,
                           int32_t* n_contrib, void* stream) {
    return gfb_internal_blend_fwd(geom_stream, feat_stream, K, tile_range, C, c0, Cg, bg, W, H, out, final_T, n_contrib,
                                  stream, false);
}

int gfb_alpha_blending_bwd(const void* geom_stream, const void* feat_stream, int64_t K,
                           const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int C, int c0, int Cg,
                           float bg, int W, int H, const float* final_T, const int32_t* n_contrib, const float* g_out,
                           float* grad_pack, void* stream) {
    return gfb_internal_blend_bwd(geom_stream, feat_stream, K, gaussian_ids_sorted, tile_range, C, c0, Cg, bg, W, H, final_T,
                                  n_contrib, g_out, grad_pack, stream, false);
}

}  // extern "C"

// no_rgb: only with Cg == 4 and c0 == 0 (rgb + depth in one blend): skip the colour channels' own gradient
int gfb_internal_blend_bwd(const void* geom_stream, const void* feat_stream, int64_t K,
                           const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int C, int c0, int Cg,
                           float bg, int W, int H, const float* final_T, const int32_t* n_contrib, const float* g_out,
                           float* grad_pack, void* stream, bool no_rgb) {
    if (no_rgb && (Cg != 4 || c0 != 0)) return GFB_E_BADARG;
    if (W <= 0 || H <= 0 || K < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (K == 0) return 0;
    if (!geom_stream || !feat_stream || !gaussian_ids_sorted || !tile_range || !final_T || !n_contrib || !g_out ||
        !grad_pack)
        return GFB_E_BADARG;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    const float4* gA = reinterpret_cast<const float4*>(geom_stream);
    const float4* gB = gA + K;
    const float4* gF = reinterpret_cast<const float4*>(feat_stream);
    const int2* tr = reinterpret_cast<const int2*>(tile_range);
    cudaStream_t st = (cudaStream_t)stream;
    const int wpc = blend_warps_per_cta();
    const int sparse = blend_bwd_sparse_lanes();
    const dim3 grid(gx * gy * (8 / wpc)), block(32 * wpc);
#define GFB_BWD_LAUNCH(CGV, SP)                                                                                          \
    blend_bwd_kernel<CGV, SP><<<grid, block, 0, st>>>(gA, gB, gF, gaussian_ids_sorted, tr, gx, c0, bg, W, H, final_T, \
                                                      n_contrib, g_out, grad_pack, sparse)
    if (no_rgb) {
        if (sparse > 0)
            blend_bwd_kernel<4, true, true><<<grid, block, 0, st>>>(gA, gB, gF, gaussian_ids_sorted, tr, gx, c0, bg, W, H,
                                                                     final_T, n_contrib, g_out, grad_pack, sparse);
        else
            blend_bwd_kernel<4, false, true><<<grid, block, 0, st>>>(gA, gB, gF, gaussian_ids_sorted, tr, gx, c0, bg, W, H,
                                                                      final_T, n_contrib, g_out, grad_pack, sparse);
    } else if (sparse > 0) {
        switch (Cg) {
            case 1: GFB_BWD_LAUNCH(1, true); break;
            case 2: GFB_BWD_LAUNCH(2, true); break;
            case 3: GFB_BWD_LAUNCH(3, true); break;
            default: GFB_BWD_LAUNCH(4, true); break;
        }
    } else {
        switch (Cg) {
            case 1: GFB_BWD_LAUNCH(1, false); break;
            case 2: GFB_BWD_LAUNCH(2, false); break;
            case 3: GFB_BWD_LAUNCH(3, false); break;
            default: GFB_BWD_LAUNCH(4, false); break;
        }
    }
#undef GFB_BWD_LAUNCH
    GFB_CHECK_LAUNCH();
    return 0;
}

extern "C" {

int gfb_blend_unpack_grads(const float* grad_pack, int N, int C, int c0, int Cg, float* d_uv, float* d_conic,
                           float* d_opacity, float* d_feature, int accumulate, void* stream) {
    if (N < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!grad_pack || !d_uv || !d_conic || !d_opacity || !d_feature) return GFB_E_BADARG;
    unpack_grads_kernel<<<gfb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(grad_pack), N, C, c0, Cg, reinterpret_cast<float2*>(d_uv), d_conic, d_opacity,
        d_feature, accumulate);
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
