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
//        A = {u, v, ext, id}  centre, half extents of the alpha >= 1/255 ellipse's bbox (2 x bf16), Gaussian id
//        B = {a, b, c, o}     conic + opacity
//        F = {f0, f1, f2, f3} up to four feature channels
//    A/B are shared by every blend with the same geometry (rgb / depth / depth-colour in
//    render_multiple) and by the backward pass.
//  * One 128-thread CTA per 16x16 tile.  Batches of 128 records are staged into shared memory by
//    1-D bulk TMA copies (cp.async.bulk + mbarrier complete_tx), double buffered, issued by
//    one thread; no thread does scattered global gathers inside the blend loop.
//  * Each warp owns an 8x8 pixel block and each lane the horizontal pixel PAIR (x, x+1) of one row.  The
//    per-pixel arithmetic of the pair runs on packed FP32 (fma/mul/add.f32x2 -> SASS FFMA2 / FMUL2 /
//    FADD2): both kernels are bound by the instruction issue rate, and a packed instruction does two
//    pixels per issue slot.  Per 32 records, the lanes test the records' bboxes against the warp's
//    block in parallel (one LDS.128 each), ballot, and the warp then walks only the surviving records.
//    With sigma ~ 1 px splats this removes ~4/5 of the (pixel, Gaussian) evaluations of a
//    256-pixel-per-Gaussian tile walk; results are unchanged because a culled pair has
//    alpha < 1/255 and would have been skipped.
//  * Backward: same staging back to front; a lane first adds its two pixels, then the per-lane
//    gradients of a (warp, Gaussian) pair are reduced with a transposed butterfly (V values in
//    ~V shuffles instead of 5 V) and added with one RED per value into a packed 48-byte-per-Gaussian
//    gradient record, so the (up to) eleven atomics of a warp hit one or two L2 sectors.
//  No tensor cores: this is a gather / scatter bounded by issue rate and L2 atomics.
#include <cstdlib>

#include "records.cuh"
#include "sort_network.cuh"
#include "splat_math.cuh"

namespace {

constexpr int kBatch = 128;          // records per TMA stage
constexpr int kBlendThreads = 128;   // 4 warps per CTA, each an 8x8 pixel block (2 pixels per lane)
// Build-time knobs (tools/build_variants.py compiles A/B variants of this file, tools/ab_blend.py times them side by
// side; defaults = the measured best, profiles/r2_blend_ab.txt):
//   GFB_BLEND_MIN_CTAS  minimum resident CTAs per SM requested from ptxas (register cap 65536 / (128 n))
//   GFB_BLEND_PAIR      1: surviving records are processed two at a time (two independent instruction streams around
//                          the short T / S recurrence; one butterfly for both records' gradients)
#ifndef GFB_BLEND_MIN_CTAS
#define GFB_BLEND_MIN_CTAS 8
#endif
#ifndef GFB_BLEND_PAIR
#define GFB_BLEND_PAIR 1
#endif
constexpr int kBlendMinCtas = GFB_BLEND_MIN_CTAS;

#ifdef GFB_BLEND_TRACE  // variant builds only: per-CTA (smid, start, end) of the last backward launch
__device__ unsigned long long g_trace[3 * 8192];
__device__ __forceinline__ uint64_t trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t trace_smid() {
    unsigned s;
    asm volatile("mov.u32 %0, %smid;" : "=r"(s));
    return s;
}
struct TraceScope {
    unsigned long long t0;
    __device__ __forceinline__ TraceScope() : t0(trace_now()) {}
    __device__ __forceinline__ ~TraceScope() {
        __syncthreads();
        if (threadIdx.x == 0 && blockIdx.x < 8192) {
            g_trace[3 * blockIdx.x] = trace_smid();
            g_trace[3 * blockIdx.x + 1] = t0;
            g_trace[3 * blockIdx.x + 2] = trace_now();
        }
    }
};
#endif
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// orders this thread's earlier generic-proxy writes (ordinary stores) before later async-proxy accesses (bulk copies)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async;" ::: "memory");
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

// Packed FP32 (Blackwell): one instruction works on the two halves of a 64-bit register pair, each half
// rounded like the scalar .rn instruction.  ptxas broadcasts a scalar operand (make_float2(s, s)) and
// folds negations into the instruction's operand modifiers, so neither costs a move.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)),
          "l"(*reinterpret_cast<uint64_t*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(d)
        : "l"(*reinterpret_cast<uint64_t*>(&a)), "l"(*reinterpret_cast<uint64_t*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 dup2(float s) { return make_float2(s, s); }

// -0.5 (a dx^2 + c dy^2) - b dx dy for the pixel pair (dx.x, dy), (dx.y, dy), every step rounded exactly as
// __fmul_rn / __fmaf_rn round it, so the forward and backward kernels take identical skip decisions on every
// (pixel, Gaussian) pair.
__device__ __forceinline__ float2 splat_power2(float a, float b, float c, float2 dx, float dy) {
    float2 q = mul2(mul2(dup2(a), dx), dx);
    q = fma2(dup2(__fmul_rn(c, dy)), dup2(dy), q);
    const float2 nr = mul2(mul2(dup2(b), dx), dup2(-dy));
    return fma2(dup2(-0.5f), q, nr);
}

// exp(power) as one FMUL + MUFU.EX2 (ex2.approx.ftz): power is in [-5.6, 0] wherever alpha can
// reach 1/255, far from the denormal range __expf() guards against with three extra instructions.
// Forward and backward share it, so they agree on every alpha >= 1/255 decision.
__device__ __forceinline__ float splat_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float2 splat_exp2(float2 power) {
    const float2 e = mul2(power, dup2(1.4426950408889634f));
    return make_float2(splat_ex2(e.x), splat_ex2(e.y));
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
    sA[k] = gfb_pack_record_a(p.x, p.y, hx, hy, id);
    sB[k] = make_float4(a, b, c, o);
}

// both streams in one launch: the first blend of a geometry (the later blends of render_multiple reuse the geometry
// stream and pack only their own features)
__global__ void __launch_bounds__(256)
pack_geometry_feature_kernel(const float2* __restrict__ uv, const float* __restrict__ conic,
                             const float* __restrict__ opacity, const float* __restrict__ feature, int C, int c0, int Cg,
                             const int32_t* __restrict__ ids, long long K, float4* __restrict__ sA,
                             float4* __restrict__ sB, float4* __restrict__ sF) {
    const long long k = (long long)blockIdx.x * 256 + threadIdx.x;
    if (k >= K) return;
    const int id = ids[k];
    const float2 p = uv[id];
    const float a = conic[3 * id], b = conic[3 * id + 1], c = conic[3 * id + 2];
    const float o = opacity[id];
    float hx, hy;
    gfbm::splat_bbox(a, b, c, o, hx, hy);
    sA[k] = gfb_pack_record_a(p.x, p.y, hx, hy, id);
    sB[k] = make_float4(a, b, c, o);
    const float* f = feature + (size_t)id * C + c0;
    float4 r = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    r.x = f[0];
    if (Cg > 1) r.y = f[1];
    if (Cg > 2) r.z = f[2];
    if (Cg > 3) r.w = f[3];
    sF[k] = r;
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

// the lane's record of the 32 starting at `base` overlaps the warp's pixel block [fx0, fx1] x [fy0, fy1]
__device__ __forceinline__ bool record_hits_block(const float4& a4, float fx0, float fx1, float fy0, float fy1) {
    float hx, hy;
    gfb_unpack_extents(a4.z, hx, hy);
    return (a4.x + hx >= fx0) && (a4.x - hx <= fx1) && (a4.y + hy >= fy0) && (a4.y - hy <= fy1);
}

// The warp's 8x8 block inside the tile and the lane's pixel pair inside the block.
struct LanePixels {
    int px, py;           // left pixel of the pair
    bool in0, in1;        // inside the image
    float fx0, fx1, fy0, fy1;
    float2 nx;            // (-px, -(px + 1))
    float pyf;
    __device__ __forceinline__ LanePixels(int tx, int ty, int warp, int lane, int W, int H) {
        const int bx0 = tx * GFB_TILE + (warp & 1) * 8, by0 = ty * GFB_TILE + (warp >> 1) * 8;
        px = bx0 + 2 * (lane & 3);
        py = by0 + (lane >> 2);
        in0 = (px < W) && (py < H);
        in1 = (px + 1 < W) && (py < H);
        fx0 = (float)bx0, fx1 = (float)(bx0 + 7), fy0 = (float)by0, fy1 = (float)(by0 + 7);
        nx = make_float2(-(float)px, -(float)(px + 1));
        pyf = (float)py;
    }
};

// ------------------------------------------------------------------ forward
template <int CG>
struct FwdPixels {  // running state of the lane's two pixels
    float2 T;  // negative once the pixel is finished (|T| = the transmittance it stopped at): no flag registers in the loop
    float2 acc[CG];
    int last0, last1;
    __device__ __forceinline__ void init(bool in0, bool in1) { T = make_float2(in0 ? 1.0f : -1.0f, in1 ? 1.0f : -1.0f); }
    __device__ __forceinline__ bool done() const { return T.x < 0.0f && T.y < 0.0f; }
};

// NR surviving records (indices jj[] of the staged batch, in list order; pos_base + jj + 1 = 1-based position in the
// tile's list) against the pixel pair.  With NR = 2 the two records' exponent / alpha evaluations are independent
// instruction streams the scheduler interleaves; only the short T recurrence is sequential.
template <int CG, int NR>
__device__ __forceinline__ void fwd_records(const Stage& st, const int (&jj)[NR], int pos_base, const LanePixels& lp,
                                            FwdPixels<CG>& s) {
    float2 alpha[NR], power[NR];
    float4 f4[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float4 a4 = st.A[jj[r]];
        const float4 b4 = st.B[jj[r]];
        f4[r] = st.F[jj[r]];
        const float dy = a4.y - lp.pyf;
        const float2 dx = add2(dup2(a4.x), lp.nx);
        power[r] = splat_power2(b4.x, b4.y, b4.z, dx, dy);
        alpha[r] = mul2(dup2(b4.w), splat_exp2(power[r]));
        alpha[r].x = fminf(GFB_ALPHA_MAX, alpha[r].x);
        alpha[r].y = fminf(GFB_ALPHA_MAX, alpha[r].y);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float2 test_T = mul2(s.T, sub2(dup2(1.0f), alpha[r]));
        // a pixel takes the record when alpha >= 1/255; it is finished (record not applied) when its transmittance
        // would fall below 1e-4.  test_T is negative for a finished pixel, so it can neither take a record nor
        // finish a second time.
        const bool ok0 = power[r].x <= 0.0f && alpha[r].x >= GFB_ALPHA_MIN;
        const bool ok1 = power[r].y <= 0.0f && alpha[r].y >= GFB_ALPHA_MIN;
        const bool hit0 = ok0 && !(test_T.x < GFB_T_EPS);
        const bool hit1 = ok1 && !(test_T.y < GFB_T_EPS);
        alpha[r].x = hit0 ? alpha[r].x : 0.0f;
        alpha[r].y = hit1 ? alpha[r].y : 0.0f;
        const float2 w = mul2(alpha[r], s.T);  // effective alpha 0: w = 0 and T (1 - 0) = T exactly
        s.T = mul2(s.T, sub2(dup2(1.0f), alpha[r]));
        s.T.x = (ok0 && !hit0) ? -fabsf(s.T.x) : s.T.x;
        s.T.y = (ok1 && !hit1) ? -fabsf(s.T.y) : s.T.y;
#pragma unroll
        for (int c = 0; c < CG; ++c) s.acc[c] = fma2(dup2(f4_get(f4[r], c)), w, s.acc[c]);
        const int pos1 = pos_base + jj[r] + 1;
        s.last0 = hit0 ? pos1 : s.last0;
        s.last1 = hit1 ? pos1 : s.last1;
    }
}

// The forward blend of one tile by the calling CTA (kBlendThreads threads): records [range.x, range.y) of the streams.
template <int CG>
__device__ __forceinline__ void blend_forward_tile(const float4* __restrict__ gA, const float4* __restrict__ gB,
                                                   const float4* __restrict__ gF, int tile, int2 range, int gx, int c0,
                                                   float bg, int W, int H, float* __restrict__ out,
                                                   float* __restrict__ final_T, int32_t* __restrict__ n_contrib,
                                                   Stage* s_stage, uint64_t* s_bar) {
    const int tx = tile % gx, ty = tile / gx;
    const int n = range.y - range.x;
    const int tid = threadIdx.x, lane = tid & 31;
    const LanePixels lp(tx, ty, tid >> 5, lane, W, H);

    FwdPixels<CG> px;
    px.init(lp.in0, lp.in1);
#pragma unroll
    for (int c = 0; c < CG; ++c) px.acc[c] = make_float2(0.0f, 0.0f);
    px.last0 = px.last1 = 0;

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
            if (!__all_sync(kFull, px.done())) {
                bool live = true;  // warp-uniform: some pixel of the block is still open
                for (int base = 0; base < cnt && live; base += 32) {
                    // lane L tests record base + 31 - L: the top ballot bit is the front-most record (FLO only; __ffs
                    // would add a BREV on the quarter-rate XU pipe the two ex2 per record already load)
                    const int j = base + 31 - lane;
                    bool hit = false;
                    if (j < cnt) hit = record_hits_block(st.A[j], lp.fx0, lp.fx1, lp.fy0, lp.fy1);
                    unsigned m = __ballot_sync(kFull, hit);
                    while (m) {
                        const int j0 = base + __clz(m);
                        m &= ~(0x80000000u >> __clz(m));
#if GFB_BLEND_PAIR
                        if (m) {
                            const int j1 = base + __clz(m);
                            m &= ~(0x80000000u >> __clz(m));
                            const int jj2[2] = {j0, j1};
                            fwd_records<CG, 2>(st, jj2, b * kBatch, lp, px);
                            continue;
                        }
#endif
                        const int jj1[1] = {j0};
                        fwd_records<CG, 1>(st, jj1, b * kBatch, lp, px);
                    }
                    live = !__all_sync(kFull, px.done());
                }
            }
            if (__syncthreads_count(px.done()) == (int)blockDim.x) {
                // drain the copy already in flight before the CTA (and its shared memory) retires
                if (tid == 0 && b + 1 < nb) mbar_wait(&s_bar[s ^ 1], (uint32_t)((b + 1) >> 1) & 1u);
                break;
            }
        }
    }
    const size_t HW = (size_t)W * H;
    if (lp.in0) {
        const size_t pix = (size_t)lp.py * W + lp.px;
#pragma unroll
        for (int c = 0; c < CG; ++c) out[(size_t)(c0 + c) * HW + pix] = fmaf(fabsf(px.T.x), bg, px.acc[c].x);
        final_T[pix] = fabsf(px.T.x);
        n_contrib[pix] = px.last0;
    }
    if (lp.in1) {
        const size_t pix = (size_t)lp.py * W + lp.px + 1;
#pragma unroll
        for (int c = 0; c < CG; ++c) out[(size_t)(c0 + c) * HW + pix] = fmaf(fabsf(px.T.y), bg, px.acc[c].y);
        final_T[pix] = fabsf(px.T.y);
        n_contrib[pix] = px.last1;
    }
}

template <int CG>
__global__ void __launch_bounds__(kBlendThreads, kBlendMinCtas)
blend_fwd_kernel(const float4* __restrict__ gA, const float4* __restrict__ gB, const float4* __restrict__ gF,
                 const int2* __restrict__ tile_range, int gx, int c0, float bg, int W, int H,
                 float* __restrict__ out, float* __restrict__ final_T, int32_t* __restrict__ n_contrib) {
    __shared__ __align__(128) Stage s_stage[2];
    __shared__ __align__(8) uint64_t s_bar[2];
    gfb_pdl_wait();  // fused pipeline: tile_sort_pack may still be draining
    blend_forward_tile<CG>(gA, gB, gF, blockIdx.x, tile_range[blockIdx.x], gx, c0, bg, W, H, out, final_T, n_contrib, s_stage,
                           s_bar);
}

// Fused pipeline: per-tile sort + record packing + forward blend in ONE kernel.  A tile's blend only needs that
// tile's records, so there is no reason to wait for every other tile's sort (a kernel boundary is a grid-wide
// barrier): the CTA sorts its segment, writes ids and the A / B / F records (the backward reads them later), and
// blends from them straight away.  Sorting is a dependent-latency chain with half-empty issue slots, blending is
// issue bound -- with both in one kernel, tiles in different phases share an SM and fill each other's gaps, and one
// launch + ramp + drain disappears.  The records written by this CTA with ordinary stores are read back by its own
// bulk copies: fence.proxy.async orders the two proxies.
static_assert(kSortThreads == kBlendThreads, "the fused kernel sorts and blends with the same CTA");
template <int CG>
__global__ void __launch_bounds__(kBlendThreads, kBlendMinCtas)
tile_sort_blend_fwd_kernel(const int32_t* __restrict__ offsets, int R, unsigned long long* __restrict__ keys,
                           int2* __restrict__ tile_range, long long capacity, GfbPackArgs pa, int gx, float bg, int W, int H,
                           float* __restrict__ out, float* __restrict__ final_T, int32_t* __restrict__ n_contrib) {
    // the sort's exchange buffer and the blend's stages are never live at the same time
    __shared__ __align__(128) unsigned char s_raw[sizeof(Stage) * 2 > sizeof(unsigned long long) * kSortSmemSmall
                                                     ? sizeof(Stage) * 2 : sizeof(unsigned long long) * kSortSmemSmall];
    __shared__ __align__(8) uint64_t s_bar[2];
    gfb_pdl_launch_dependents();
    gfb_pdl_wait();  // keys come from scatter
    sort_tile_cta(offsets, R, keys, capacity, reinterpret_cast<unsigned long long*>(s_raw), tile_range,
                  [pa](long long pos, unsigned long long key) { gfb_write_record(pa, pos, (int)(unsigned int)key); });
    const int tile = blockIdx.x;
    const int start = offsets[tile * R];
    long long end = offsets[(tile + 1) * R];
    if (end > capacity) end = max((long long)start, capacity);
    const int2 range = (end > start) ? make_int2(start, (int)end) : make_int2(0, 0);
    fence_proxy_async();
    __syncthreads();
    blend_forward_tile<CG>(pa.sA, pa.sB, pa.sF, tile, range, gx, 0, bg, W, H, out, final_T, n_contrib,
                           reinterpret_cast<Stage*>(s_raw), s_bar);
}

// ------------------------------------------------------------------ backward
// Transposed butterfly over the warp: every lane contributes v[0..N); at each of the five stages a lane keeps
// one half of its values and sends the other half to its partner, so N values cost about N shuffles instead of
// 5 N.  bfly_slot() tells which value's warp total a lane ends up holding.
template <int N, int MASK>
__device__ __forceinline__ float bfly_reduce(const float (&v)[N], int lane) {
    if constexpr (MASK == 0) {
        static_assert(N == 1, "at most 32 values");
        return v[0];
    } else {
        constexpr int HN = (N + 1) / 2;
        const bool up = (lane & MASK) != 0;
        float keep[HN], recv[HN], r[HN];
#pragma unroll
        for (int k = 0; k < HN; ++k) {
            if (k + HN < N) {  // a pair: lower lanes collect v[k], upper lanes v[k + HN]
                const float send = up ? v[k] : v[k + HN];
                keep[k] = up ? v[k + HN] : v[k];
                recv[k] = __shfl_xor_sync(kFull, send, MASK);
            } else {  // odd one out: both halves end with its total (the upper copy is ignored)
                keep[k] = v[k];
                recv[k] = __shfl_xor_sync(kFull, v[k], MASK);
            }
        }
#pragma unroll
        for (int k = 0; k + 1 < HN; k += 2) {
            const float2 t = add2(make_float2(keep[k], keep[k + 1]), make_float2(recv[k], recv[k + 1]));
            r[k] = t.x;
            r[k + 1] = t.y;
        }
        if (HN & 1) r[HN - 1] = keep[HN - 1] + recv[HN - 1];
        return bfly_reduce<HN, MASK / 2>(r, lane);
    }
}
// index of the value whose total bfly_reduce<V, 16> leaves in this lane, or -1 for a duplicate
template <int V>
__device__ __forceinline__ int bfly_slot(int lane) {
    int n[6];
    n[0] = V;
#pragma unroll
    for (int s = 0; s < 5; ++s) n[s + 1] = (n[s] + 1) / 2;
    int k = 0;
    bool valid = true;
#pragma unroll
    for (int s = 4; s >= 0; --s) {
        if (lane & (16 >> s)) {
            if (k + n[s + 1] < n[s]) k += n[s + 1];
            else valid = false;
        }
    }
    return valid ? k : -1;
}

template <int CG>
struct BwdPixels {  // running state of the lane's two pixels
    float2 T, S;
    float2 go[CG];
    int last0, last1;
};

// NO_RGB (CG == 4 only; the native fit loop on frames >= 1, where GFlow zeroes the colour gradient, trainer.py:537-540):
// channels 0..2 still enter dalpha through f . g, but their own gradient is neither reduced nor added -- only the
// depth channel's gradient joins the six geometry values in the butterfly.
//
// NR surviving records (indices jj[] of the staged batch, jj[0] the one furthest back; pos0 + jj = 0-based position in
// the tile's list), back to front.  With NR = 2 everything but the short T / S recurrence is two independent
// instruction streams, and ONE butterfly reduces the 2 V values of both records.
// slot_rec / slot_off: which record and which grad_pack column this lane's butterfly total belongs to (-1: none).
template <int CG, bool NO_RGB, int NR>
__device__ __forceinline__ void bwd_records(const Stage& st, const int (&jj)[NR], int pos0, const LanePixels& lp, int lane,
                                            int slot_rec, int slot_off, float* __restrict__ grad_pack, BwdPixels<CG>& s) {
    constexpr int V = NO_RGB ? 7 : 6 + CG;
    float4 a4[NR], b4[NR];
    float dy[NR];
    float2 dx[NR], G[NR], alpha[NR];
    bool act0[NR], act1[NR];
    bool any = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        a4[r] = st.A[jj[r]];
        b4[r] = st.B[jj[r]];
        dy[r] = a4[r].y - lp.pyf;
        dx[r] = add2(dup2(a4[r].x), lp.nx);
        const float2 power = splat_power2(b4[r].x, b4[r].y, b4[r].z, dx[r], dy[r]);
        G[r] = splat_exp2(power);
        alpha[r] = mul2(dup2(b4[r].w), G[r]);
        alpha[r].x = fminf(GFB_ALPHA_MAX, alpha[r].x);
        alpha[r].y = fminf(GFB_ALPHA_MAX, alpha[r].y);
        const int pos = pos0 + jj[r];
        act0[r] = pos < s.last0 && power.x <= 0.0f && alpha[r].x >= GFB_ALPHA_MIN;
        act1[r] = pos < s.last1 && power.y <= 0.0f && alpha[r].y >= GFB_ALPHA_MIN;
        any = any || act0[r] || act1[r];
    }
    if (!__any_sync(kFull, any)) return;
    float v[NR * V];
    float2 inv[NR], fg[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        // a pixel the record did not touch rides along with alpha = G = 0: T, S unchanged, zero gradient
        G[r].x = act0[r] ? G[r].x : 0.0f;
        G[r].y = act1[r] ? G[r].y : 0.0f;
        alpha[r].x = act0[r] ? alpha[r].x : 0.0f;
        alpha[r].y = act1[r] ? alpha[r].y : 0.0f;
        const float2 om = sub2(dup2(1.0f), alpha[r]);
        inv[r] = make_float2(splat_rcp(om.x), splat_rcp(om.y));
        const float4 f4 = st.F[jj[r]];
        fg[r] = mul2(dup2(f4.x), s.go[0]);
#pragma unroll
        for (int c = 1; c < CG; ++c) fg[r] = fma2(dup2(f4_get(f4, c)), s.go[c], fg[r]);
    }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        s.T = mul2(s.T, inv[r]);  // transmittance in front of this Gaussian
        const float2 w = mul2(alpha[r], s.T);
        const float2 sinv = mul2(s.S, inv[r]);
        const float2 dalpha = fma2(s.T, fg[r], make_float2(-sinv.x, -sinv.y));
        s.S = fma2(w, fg[r], s.S);
        // moments of nq = -G dG over the lane's two pixels (dG = o dalpha)
        const float2 nq = mul2(G[r], mul2(dup2(-b4[r].w), dalpha));
        const float2 m0 = mul2(G[r], dalpha);
        const float2 nqx = mul2(nq, dx[r]);
        const float2 nqxx = mul2(nqx, dx[r]);
        const float Mq = nq.x + nq.y, Mx = nqx.x + nqx.y, Mxx = nqxx.x + nqxx.y;
        const float My = dy[r] * Mq, Mxy = dy[r] * Mx, Myy = dy[r] * My;
        float* vr = v + r * V;
        vr[0] = fmaf(b4[r].x, Mx, b4[r].y * My);   // dL/du
        vr[1] = fmaf(b4[r].z, My, b4[r].y * Mx);   // dL/dv
        vr[2] = 0.5f * Mxx;                        // dL/da
        vr[3] = Mxy;                               // dL/db
        vr[4] = 0.5f * Myy;                        // dL/dc
        vr[5] = m0.x + m0.y;                       // dL/do
        if constexpr (NO_RGB) {
            const float2 d3 = mul2(w, s.go[3]);
            vr[6] = d3.x + d3.y;
        } else {
#pragma unroll
            for (int c = 0; c < CG; ++c) {
                const float2 dc = mul2(w, s.go[c]);
                vr[6 + c] = dc.x + dc.y;
            }
        }
    }
    const float red = bfly_reduce<NR * V, 16>(v, lane);
    if (slot_off >= 0) {
        const float idf = (NR == 2 && slot_rec == 1) ? a4[NR - 1].w : a4[0].w;
        atomicAdd(grad_pack + (size_t)__float_as_int(idf) * 12 + slot_off, red);
    }
}

template <int CG, bool NO_RGB = false>
__global__ void __launch_bounds__(kBlendThreads, kBlendMinCtas)
blend_bwd_kernel(const float4* __restrict__ gA, const float4* __restrict__ gB, const float4* __restrict__ gF,
                 const int2* __restrict__ tile_range, int gx, int c0, float bg, int W, int H,
                 const float* __restrict__ final_T, const int32_t* __restrict__ n_contrib,
                 const float* __restrict__ g_out, float* __restrict__ grad_pack, float* __restrict__ zero16) {
    __shared__ __align__(128) Stage s_stage[2];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_max_last;

    gfb_pdl_launch_dependents();  // fused pipeline: geometry_bwd may queue behind the last wave
    // the camera-gradient block geometry_bwd accumulates into afterwards (saves the caller a memset launch)
    if (zero16 != nullptr && blockIdx.x == 0 && threadIdx.x < 16) zero16[threadIdx.x] = 0.0f;
#ifdef GFB_BLEND_TRACE
    TraceScope trace_scope;
#endif
    const int tile = blockIdx.x;
    const int tx = tile % gx, ty = tile / gx;
    const int2 range = tile_range[tile];
    const int n = range.y - range.x;
    if (n <= 0) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const LanePixels lp(tx, ty, tid >> 5, lane, W, H);

    // values reduced per (warp, record): du dv da db dc do, then the feature gradients
    constexpr int V = NO_RGB ? 7 : 6 + CG;
    // grad_pack row: {du dv da db dc do df0..3 - -}; the NO_RGB butterfly's seventh value is the depth channel (column 9)
    const int slot1 = bfly_slot<V>(lane);
    const int off1 = (NO_RGB && slot1 == 6) ? 9 : slot1;
#if GFB_BLEND_PAIR
    const int slot2 = bfly_slot<2 * V>(lane);  // butterfly over the 2 V values of a record pair
    const int rec2 = slot2 >= V ? 1 : 0;
    const int k2 = slot2 >= V ? slot2 - V : slot2;
    const int off2 = slot2 < 0 ? -1 : ((NO_RGB && k2 == 6) ? 9 : k2);
#endif

    BwdPixels<CG> px;
    float2 Tf = make_float2(1.0f, 1.0f);
    px.last0 = px.last1 = 0;
#pragma unroll
    for (int c = 0; c < CG; ++c) px.go[c] = make_float2(0.0f, 0.0f);
    const size_t HW = (size_t)W * H;
    if (lp.in0) {
        const size_t pix = (size_t)lp.py * W + lp.px;
        Tf.x = final_T[pix];
        px.last0 = n_contrib[pix];
#pragma unroll
        for (int c = 0; c < CG; ++c) px.go[c].x = g_out[(size_t)(c0 + c) * HW + pix];
    }
    if (lp.in1) {
        const size_t pix = (size_t)lp.py * W + lp.px + 1;
        Tf.y = final_T[pix];
        px.last1 = n_contrib[pix];
#pragma unroll
        for (int c = 0; c < CG; ++c) px.go[c].y = g_out[(size_t)(c0 + c) * HW + pix];
    }
    float2 bgdot = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int c = 0; c < CG; ++c) bgdot = fma2(dup2(bg), px.go[c], bgdot);

    const int wmax = __reduce_max_sync(kFull, max(px.last0, px.last1));
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

    px.T = Tf;
    // S = sum_{k behind j} w_k (f_k . g) + T_final (bg . g): one scalar per pixel carries what the 3DGS
    // formulation keeps as accum_rec[C] / last_color[C] / last_alpha
    px.S = mul2(Tf, bgdot);

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
                if (j < cnt && pos0 + j < wmax) hit = record_hits_block(st.A[j], lp.fx0, lp.fx1, lp.fy0, lp.fy1);
                unsigned m = __ballot_sync(kFull, hit);
                while (m) {
                    const int bit = 31 - __clz(m);
                    m &= ~(1u << bit);
#if GFB_BLEND_PAIR
                    if (m) {
                        const int bit1 = 31 - __clz(m);
                        m &= ~(1u << bit1);
                        const int jj2[2] = {base + bit, base + bit1};
                        bwd_records<CG, NO_RGB, 2>(st, jj2, pos0, lp, lane, rec2, off2, grad_pack, px);
                        continue;
                    }
#endif
                    const int jj1[1] = {base + bit};
                    bwd_records<CG, NO_RGB, 1>(st, jj1, pos0, lp, lane, 0, off1, grad_pack, px);
                }
            }
        }
        __syncthreads();  // every warp is done with stage s before it is refilled (it+2)
    }
}

__global__ void __launch_bounds__(256)
unpack_grads_kernel(float4* __restrict__ grad_pack, int N, int C, int c0, int Cg, float2* __restrict__ d_uv,
                    float* __restrict__ d_conic, float* __restrict__ d_opacity, float* __restrict__ d_feature,
                    int flags) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const float4 g0 = grad_pack[3 * (size_t)i], g1 = grad_pack[3 * (size_t)i + 1], g2 = grad_pack[3 * (size_t)i + 2];
    if (flags & GFB_UNPACK_CLEAR) {  // a pack the caller keeps between calls goes back to zero here: no memset per call
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        grad_pack[3 * (size_t)i] = z, grad_pack[3 * (size_t)i + 1] = z, grad_pack[3 * (size_t)i + 2] = z;
    }
    if (flags & GFB_UNPACK_ACCUMULATE) {
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

int gfb_blend_pack_geometry_feature(const float* uv, const float* conic, const float* opacity, const float* feature,
                                    int C, int c0, int Cg, const int32_t* gaussian_ids_sorted, int64_t K,
                                    void* geom_stream, void* feat_stream, void* stream) {
    if (K < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (K == 0) return 0;
    if (!uv || !conic || !opacity || !feature || !gaussian_ids_sorted || !geom_stream || !feat_stream) return GFB_E_BADARG;
    float4* sA = reinterpret_cast<float4*>(geom_stream);
    pack_geometry_feature_kernel<<<gfb_div_up(K, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float2*>(uv), conic, opacity, feature, C, c0, Cg, gaussian_ids_sorted, (long long)K, sA,
        sA + K, reinterpret_cast<float4*>(feat_stream));
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
    const dim3 grid(gx * gy), block(kBlendThreads);
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
                           const float*, const int32_t*, const float*, float*, void*, bool no_rgb, float* zero16);

// Fused pipeline / native loop: per-tile sort + pack + forward blend of all C <= 4 channels in one kernel
// (tile_sort_blend_fwd_kernel).  tile_offsets / R: the scanned (tile, replica) counters of the control block;
// keys_ws: the keys scatter wrote.  Writes ids, both record streams, tile_range, out, final_T, n_contrib.
int gfb_internal_sort_pack_blend_fwd(const int32_t* tile_offsets, int R, void* keys_ws, int32_t* tile_range, int64_t capacity,
                                     const float* uv, const float* conic, const float* opacity, const float* feature, int C,
                                     int32_t* gaussian_ids_sorted, void* geom_stream, void* feat_stream, float bg, int W,
                                     int H, float* out, float* final_T, int32_t* n_contrib, void* stream, bool pdl) {
    if (W <= 0 || H <= 0 || capacity < 0 || C < 1 || C > 4) return GFB_E_BADARG;
    if (!tile_offsets || !tile_range || !out || !final_T || !n_contrib) return GFB_E_BADARG;
    const int gx = (W + GFB_TILE - 1) / GFB_TILE, gy = (H + GFB_TILE - 1) / GFB_TILE;
    float4* sA = reinterpret_cast<float4*>(geom_stream);
    GfbPackArgs pa{reinterpret_cast<const float2*>(uv), conic, opacity, feature, C, sA, sA + capacity,
                   reinterpret_cast<float4*>(feat_stream), gaussian_ids_sorted};
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(keys_ws);
    int2* tr = reinterpret_cast<int2*>(tile_range);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 grid(gx * gy), block(kBlendThreads);
    const long long cap = (long long)capacity;
    cudaError_t le;
    switch (C) {
        case 1: le = gfb_launch_pdl(tile_sort_blend_fwd_kernel<1>, grid, block, st, pdl, tile_offsets, R, keys, tr, cap, pa, gx, bg, W, H, out, final_T, n_contrib); break;
        case 2: le = gfb_launch_pdl(tile_sort_blend_fwd_kernel<2>, grid, block, st, pdl, tile_offsets, R, keys, tr, cap, pa, gx, bg, W, H, out, final_T, n_contrib); break;
        case 3: le = gfb_launch_pdl(tile_sort_blend_fwd_kernel<3>, grid, block, st, pdl, tile_offsets, R, keys, tr, cap, pa, gx, bg, W, H, out, final_T, n_contrib); break;
        default: le = gfb_launch_pdl(tile_sort_blend_fwd_kernel<4>, grid, block, st, pdl, tile_offsets, R, keys, tr, cap, pa, gx, bg, W, H, out, final_T, n_contrib); break;
    }
    if (le != cudaSuccess) return (int)le;
    GFB_CHECK_LAUNCH();
    return 0;
}

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
                                  n_contrib, g_out, grad_pack, stream, false, nullptr);
}

}  // extern "C"

// no_rgb: only with Cg == 4 and c0 == 0 (rgb + depth in one blend): skip the colour channels' own gradient
// zero16: optional 16 floats cleared by the kernel before anything else (the camera gradients of geometry_bwd)
int gfb_internal_blend_bwd(const void* geom_stream, const void* feat_stream, int64_t K,
                           const int32_t* gaussian_ids_sorted, const int32_t* tile_range, int C, int c0, int Cg,
                           float bg, int W, int H, const float* final_T, const int32_t* n_contrib, const float* g_out,
                           float* grad_pack, void* stream, bool no_rgb, float* zero16) {
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
    const dim3 grid(gx * gy), block(kBlendThreads);
    (void)gaussian_ids_sorted;  // the record stream carries the Gaussian id (gfb_pack_record_a)
#define GFB_BWD_LAUNCH(CGV, NORGB)                                                                                    \
    blend_bwd_kernel<CGV, NORGB><<<grid, block, 0, st>>>(gA, gB, gF, tr, gx, c0, bg, W, H, final_T, n_contrib, g_out, \
                                                         grad_pack, zero16)
    if (no_rgb) {
        GFB_BWD_LAUNCH(4, true);
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

int gfb_blend_unpack_grads(float* grad_pack, int N, int C, int c0, int Cg, float* d_uv, float* d_conic,
                           float* d_opacity, float* d_feature, int flags, void* stream) {
    if (N < 0 || C <= 0 || c0 < 0 || Cg < 1 || Cg > 4 || c0 + Cg > C) return GFB_E_BADARG;
    if (N == 0) return 0;
    if (!grad_pack || !d_uv || !d_conic || !d_opacity || !d_feature) return GFB_E_BADARG;
    unpack_grads_kernel<<<gfb_div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<float4*>(grad_pack), N, C, c0, Cg, reinterpret_cast<float2*>(d_uv), d_conic, d_opacity,
        d_feature, flags);
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"

#ifdef GFB_BLEND_TRACE
extern "C" int gfb_debug_blend_trace(unsigned long long* host_dst) {
    return (int)cudaMemcpyFromSymbol(host_dst, g_trace, sizeof(g_trace));
}
#endif
