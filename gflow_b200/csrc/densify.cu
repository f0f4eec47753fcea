// densify.cu -- error-driven densification on the device (sm_100a).
//
// Reference: SimpleGaussian.densify_by_pixels, /root/reference/gflow/trainer.py:878-939, which copies the
// per-pixel error map to the host, draws pixels with np.random.choice(p = error / sum) and builds the new
// Gaussians back on the device (a full GPU -> CPU -> GPU round trip per densification).  Here the map never
// leaves HBM:
//
//   rgb_error_map     loss_rgb_pixel of trainer.py:457: mean over the 3 channels of (rendered - target)^2
//   densify_minpos    np.nanmin(error_map[error_map > 0])                      (atomicMin on the float bits)
//   densify_weights   w = (error + minpos) * mask, mask = given or (error + minpos) > threshold; per-block
//                     (1024 pixel) sums, number of mask pixels
//   densify_scan      exclusive scan of the block sums (one CTA)
//   densify_sample    one thread per sample: u ~ U[0,1) from a counter-based generator, inverse CDF by
//                     binary search over the block prefix + a walk inside the 1024-pixel block; minimum
//                     sampled depth by atomicMin
//   densify_emit      new raw attributes of trainer.py:908-933: xyz = pix2world(pixel, gt_depth)
//                     (geometry.py:104-116, focal = fx for both axes), scale = depth / (min depth * num_points),
//                     rgb = logit(target colour), rotate = (1,0,0,0), opacity = logit(0.99) / 10
//
// Sampling is WITH replacement and proportional to w, like np.random.choice; the random stream itself is
// not numpy's, so parity of the drawn set is distributional, parity of everything computed from a drawn
// pixel is exact (tests feed the drawn pixels to the oracle).  HBM-bound streaming over P pixels (12 B read
// + 4 B written per pixel for the map, 5 B + 4 B for the weights); sampling is latency-bound and tiny.
#include "splat_math.cuh"

namespace {

constexpr int kDBlock = 1024;  // pixels per sampling block
enum { DS_MINPOS = 0, DS_MASK_COUNT = 1, DS_MIN_DEPTH = 2, DS_TOTAL = 3, DS_WORDS = 8 };
constexpr unsigned int kNone = 0x7f7f7f7fu;  // "no value yet" for the atomicMin words (what a byte-wise memset can write)

__global__ void __launch_bounds__(256)
rgb_error_map_kernel(const float* __restrict__ rendered, const float* __restrict__ gt_image,
                     const uint8_t* __restrict__ mask, int P, float* __restrict__ error_map) {
    const int pix = blockIdx.x * 256 + threadIdx.x;
    if (pix >= P) return;
    const float m = mask ? (mask[pix] ? 1.0f : 0.0f) : 1.0f;
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float r = rendered[(size_t)c * P + pix] * m - gt_image[(size_t)pix * 3 + c] * m;
        s += r * r;
    }
    error_map[pix] = s / 3.0f;
}

__global__ void __launch_bounds__(256)
densify_minpos_kernel(const float* __restrict__ error_map, int P, unsigned int* __restrict__ stats) {
    unsigned int best = kNone;
    for (int pix = blockIdx.x * 256 + threadIdx.x; pix < P; pix += gridDim.x * 256) {
        const float e = error_map[pix];
        if (e > 0.0f) best = min(best, __float_as_uint(e));  // positive floats order like their bit patterns
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best != kNone) atomicMin(stats + DS_MINPOS, best);
}

__global__ void __launch_bounds__(256)
densify_weights_kernel(const float* __restrict__ error_map, const uint8_t* __restrict__ mask, int P, float threshold,
                       unsigned int* __restrict__ stats, float* __restrict__ weights, float* __restrict__ block_sum) {
    __shared__ float s_sum[8];
    __shared__ int s_cnt[8];
    const unsigned int mbits = stats[DS_MINPOS];
    const float minpos = (mbits == kNone) ? 0.0f : __uint_as_float(mbits);  // all-zero map: nothing to add
    float sum = 0.0f;
    int cnt = 0;
    const int base = blockIdx.x * kDBlock;
#pragma unroll
    for (int k = 0; k < kDBlock / 256; ++k) {
        const int pix = base + k * 256 + threadIdx.x;
        if (pix < P) {
            const float e = error_map[pix] + minpos;
            const bool on = mask ? (mask[pix] != 0) : (e > threshold);
            const float w = on ? e : 0.0f;
            weights[pix] = w;
            sum += w;
            cnt += on ? 1 : 0;
        }
    }
    sum = gfb_warp_sum(sum);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) {
        s_sum[threadIdx.x >> 5] = sum;
        s_cnt[threadIdx.x >> 5] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        int c = 0;
        for (int w = 0; w < 8; ++w) {
            t += s_sum[w];
            c += s_cnt[w];
        }
        block_sum[blockIdx.x] = t;
        if (c) atomicAdd(reinterpret_cast<int*>(stats) + DS_MASK_COUNT, c);
    }
}

// exclusive scan of B block sums into prefix[0..B]; double accumulation keeps the tail exact enough for
// the binary search (B is a few hundred to a few thousand)
__global__ void densify_scan_kernel(const float* __restrict__ block_sum, int B, float* __restrict__ prefix,
                                    unsigned int* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double run = 0.0;
    for (int b = 0; b < B; ++b) {
        prefix[b] = (float)run;
        run += (double)block_sum[b];
    }
    prefix[B] = (float)run;
    stats[DS_TOTAL] = __float_as_uint((float)run);
}

// splitmix64 of (seed, counter) -> 24 uniform bits
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long ctr) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(256)
densify_sample_kernel(const float* __restrict__ weights, const float* __restrict__ prefix, int B, int P,
                      const float* __restrict__ gt_depth, int count, unsigned long long seed,
                      int32_t* __restrict__ sampled, unsigned int* __restrict__ stats) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    unsigned int dbits = kNone;
    if (s < count) {
        const float total = prefix[B];
        const float target = uniform01(seed, (unsigned long long)s) * total;
        int lo = 0, hi = B;  // largest b with prefix[b] <= target
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] <= target) lo = mid; else hi = mid;
        }
        const int first = lo * kDBlock, last = min(P, first + kDBlock);
        float run = prefix[lo];
        int pick = -1, last_pos = -1;
        for (int pix = first; pix < last; ++pix) {
            const float w = weights[pix];
            if (w > 0.0f) {
                last_pos = pix;
                run += w;
                if (run > target) {
                    pick = pix;
                    break;
                }
            }
        }
        if (pick < 0) pick = last_pos;  // rounding pushed the target past the block's last positive weight
        if (pick < 0) {                 // (empty block hit through rounding) first positive weight anywhere after it
            for (int pix = last; pix < P && pick < 0; ++pix)
                if (weights[pix] > 0.0f) pick = pix;
            for (int pix = first - 1; pix >= 0 && pick < 0; --pix)
                if (weights[pix] > 0.0f) pick = pix;
        }
        sampled[s] = pick;
        if (pick >= 0) {
            const float d = gt_depth[pick];
            if (d > 0.0f) dbits = __float_as_uint(d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dbits = min(dbits, __shfl_xor_sync(0xffffffffu, dbits, o));
    if ((threadIdx.x & 31) == 0 && dbits != kNone) atomicMin(stats + DS_MIN_DEPTH, dbits);
}

__device__ __forceinline__ float logit1(float x) { return logf(x / (1.0f - x)); }

__global__ void __launch_bounds__(256)
densify_emit_kernel(const int32_t* __restrict__ sampled, int count, const float* __restrict__ gt_image,
                    const float* __restrict__ gt_depth, const float* __restrict__ intr, const float* __restrict__ extr,
                    int W, int num_points, const unsigned int* __restrict__ stats, float* __restrict__ xyz,
                    float* __restrict__ scale, float4* __restrict__ rotate, float* __restrict__ opacity,
                    float* __restrict__ rgb) {
    const int s = blockIdx.x * 256 + threadIdx.x;
    if (s >= count) return;
    const int pix = max(sampled[s], 0);
    const int py = pix / W, px = pix - py * W;
    const float d = gt_depth[pix];
    // geometry.py:115-116: (depth (xy - pp) / focal, depth) with focal = intr[0] for BOTH axes
    const float f = intr[0];
    const float xc = d * ((float)px - intr[2]) / f, yc = d * ((float)py - intr[3]) / f, zc = d;
    // cam2world of the rigid [R|t]: R^T (p_cam - t)
    const float tx = xc - extr[3], ty = yc - extr[7], tz = zc - extr[11];
    xyz[3 * s] = extr[0] * tx + extr[4] * ty + extr[8] * tz;
    xyz[3 * s + 1] = extr[1] * tx + extr[5] * ty + extr[9] * tz;
    xyz[3 * s + 2] = extr[2] * tx + extr[6] * ty + extr[10] * tz;
    const unsigned int mb = stats[DS_MIN_DEPTH];
    const float dmin = (mb == kNone) ? 1.0f : __uint_as_float(mb);
    const float sc = (1.0f / (float)num_points) * (d / dmin);  // trainer.py:912-914
    scale[3 * s] = sc;
    scale[3 * s + 1] = sc;
    scale[3 * s + 2] = sc;
    rotate[s] = make_float4(1.0f, 0.0f, 0.0f, 0.0f);
    opacity[s] = logit1(0.99f) / 10.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // trainer.py:925-928: clamp(rgb, 1e-15, 1 - 1e-15) in float32 (the upper bound rounds to 1.0f), then logit
        const float v = fminf(fmaxf(gt_image[(size_t)pix * 3 + c], 1e-15f), 1.0f);
        rgb[3 * s + c] = logit1(v);
    }
}

struct DLayout {
    size_t stats, weights, block_sum, prefix, total;
    int B;
};

bool d_layout(int W, int H, DLayout& L) {
    if (W <= 0 || H <= 0) return false;
    const size_t P = (size_t)W * H;
    L.B = (int)((P + kDBlock - 1) / kDBlock);
    size_t off = 0;
    auto take = [&off](size_t bytes) {
        const size_t at = off;
        off += (bytes + 255) & ~(size_t)255;
        return at;
    };
    L.stats = take(DS_WORDS * 4);
    L.weights = take(P * 4);
    L.block_sum = take((size_t)L.B * 4);
    L.prefix = take(((size_t)L.B + 1) * 4);
    L.total = off;
    return true;
}

}  // namespace

extern "C" {

int gfb_rgb_error_map(const float* rendered, const float* gt_image, const uint8_t* pixel_mask, int W, int H,
                      float* error_map, void* stream) {
    if (W <= 0 || H <= 0 || !rendered || !gt_image || !error_map) return GFB_E_BADARG;
    const int P = W * H;
    rgb_error_map_kernel<<<gfb_div_up(P, 256), 256, 0, (cudaStream_t)stream>>>(rendered, gt_image, pixel_mask, P, error_map);
    GFB_CHECK_LAUNCH();
    return 0;
}

size_t gfb_densify_workspace_bytes(int W, int H) {
    DLayout L;
    return d_layout(W, H, L) ? L.total : 0;
}

int gfb_densify_prepare(const float* error_map, const uint8_t* mask, int W, int H, float error_threshold, void* workspace,
                        void* stream) {
    DLayout L;
    if (!error_map || !workspace || !d_layout(W, H, L)) return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    const int P = W * H;
    unsigned int* stats = (unsigned int*)(ws + L.stats);
    GFB_TRY(cudaMemsetAsync(stats, 0, DS_WORDS * 4, st));
    GFB_TRY(cudaMemsetAsync(stats + DS_MINPOS, 0x7f, 4, st));      // kNone
    GFB_TRY(cudaMemsetAsync(stats + DS_MIN_DEPTH, 0x7f, 4, st));
    densify_minpos_kernel<<<min(1024, gfb_div_up(P, 256)), 256, 0, st>>>(error_map, P, stats);
    GFB_CHECK_LAUNCH();
    densify_weights_kernel<<<L.B, 256, 0, st>>>(error_map, mask, P, error_threshold, stats, (float*)(ws + L.weights),
                                                (float*)(ws + L.block_sum));
    GFB_CHECK_LAUNCH();
    densify_scan_kernel<<<1, 32, 0, st>>>((const float*)(ws + L.block_sum), L.B, (float*)(ws + L.prefix), stats);
    GFB_CHECK_LAUNCH();
    return 0;
}

int gfb_densify_sample(const void* workspace, const float* gt_image, const float* gt_depth, const float* intr,
                       const float* extr, int W, int H, int count, int num_points, uint64_t seed, float* new_xyz,
                       float* new_scale, float* new_rotate, float* new_opacity, float* new_rgb, int32_t* sampled_pixels,
                       void* stream) {
    DLayout L;
    if (!workspace || !d_layout(W, H, L) || count < 0 || num_points <= 0) return GFB_E_BADARG;
    if (count == 0) return 0;
    if (!gt_image || !gt_depth || !intr || !extr || !new_xyz || !new_scale || !new_rotate || !new_opacity || !new_rgb ||
        !sampled_pixels)
        return GFB_E_BADARG;
    cudaStream_t st = (cudaStream_t)stream;
    const char* ws = (const char*)workspace;
    unsigned int* stats = (unsigned int*)(ws + L.stats);
    densify_sample_kernel<<<gfb_div_up(count, 256), 256, 0, st>>>((const float*)(ws + L.weights), (const float*)(ws + L.prefix),
                                                                 L.B, W * H, gt_depth, count, (unsigned long long)seed,
                                                                 sampled_pixels, stats);
    GFB_CHECK_LAUNCH();
    densify_emit_kernel<<<gfb_div_up(count, 256), 256, 0, st>>>(sampled_pixels, count, gt_image, gt_depth, intr, extr, W,
                                                               num_points, stats, new_xyz, new_scale,
                                                               reinterpret_cast<float4*>(new_rotate), new_opacity, new_rgb);
    GFB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
