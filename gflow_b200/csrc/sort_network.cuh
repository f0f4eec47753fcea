// sort_network.cuh -- direction-free bitonic networks on 64-bit keys (depth bits << 32 | id) used by the
// per-tile segmented sort (binning.cu, pipeline.cu).
#pragma once
#include "common.cuh"

constexpr int kSortThreads = 128;
constexpr int kSortSmemKeys = 4096;  // 32 KB of 64-bit keys per CTA

// Direction-free bitonic network: every compare-exchange puts the smaller key at the lower
// index ("flip" first stage of each merge, then plain half-cleaners).  Because all exchanges are
// ascending, virtual +inf keys at indices >= n never move, so pairs whose upper index is >= n
// are skipped and no padding is stored.  `buf` may be shared or global memory; one CTA.
static __device__ void bitonic_sort_block(unsigned long long* buf, int n, int n_pad) {
    for (int k = 2; k <= n_pad; k <<= 1) {
        const int half = k >> 1;
        for (int q = threadIdx.x; q < (n_pad >> 1); q += blockDim.x) {
            const int blk = q / half, pos = q - blk * half;
            const int lo = blk * k + pos, hi = blk * k + (k - 1 - pos);
            if (hi < n) {
                const unsigned long long a = buf[lo], b = buf[hi];
                if (a > b) {
                    buf[lo] = b;
                    buf[hi] = a;
                }
            }
        }
        __syncthreads();
        for (int j = k >> 2; j > 0; j >>= 1) {
            for (int q = threadIdx.x; q < (n_pad >> 1); q += blockDim.x) {
                const int lo = 2 * q - (q & (j - 1));
                const int hi = lo + j;
                if (hi < n) {
                    const unsigned long long a = buf[lo], b = buf[hi];
                    if (a > b) {
                        buf[lo] = b;
                        buf[hi] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ unsigned long long u64_min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
__device__ __forceinline__ unsigned long long u64_max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// Segments of <= 64 keys: one warp, element e = lane (k0) and lane + 32 (k1); the same
// direction-free network with every compare-exchange done by warp shuffle.
__device__ __forceinline__ void warp_sort64(unsigned long long& k0, unsigned long long& k1, int lane) {
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
        if (k < 64) {
            // flip stage: partner element e ^ (k-1), same register, lane ^ (k-1)
            const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, k0, k - 1);
            const unsigned long long p1 = __shfl_xor_sync(0xffffffffu, k1, k - 1);
            const bool lower = (lane & (k >> 1)) == 0;
            k0 = lower ? u64_min(k0, p0) : u64_max(k0, p0);
            k1 = lower ? u64_min(k1, p1) : u64_max(k1, p1);
        } else {
            // k == 64: partner of e is e ^ 63 -> the other register of lane ^ 31
            const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, k1, 31);
            const unsigned long long p1 = __shfl_xor_sync(0xffffffffu, k0, 31);
            k0 = u64_min(k0, p0);
            k1 = u64_max(k1, p1);
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, k0, j);
            const unsigned long long p1 = __shfl_xor_sync(0xffffffffu, k1, j);
            const bool lower = (lane & j) == 0;
            k0 = lower ? u64_min(k0, p0) : u64_max(k0, p0);
            k1 = lower ? u64_min(k1, p1) : u64_max(k1, p1);
        }
    }
}

