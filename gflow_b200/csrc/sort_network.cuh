// sort_network.cuh -- direction-free bitonic networks on 64-bit keys (depth bits << 32 | id) used by the
// per-tile segmented sort (binning.cu, pipeline.cu).
#pragma once
#include "common.cuh"

constexpr int kSortThreads = 128;

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


// ------------------------------------------------------------------ register-resident warp sort
// 32*R keys per warp, element e = r*32 + lane.  Same direction-free network; exchanges with a
// stride below 32 are warp shuffles, strides of 32 and above are in-thread register exchanges, so
// a tile segment of up to 256 keys is sorted by one warp with no shared memory and no barrier.
template <int R>
__device__ __forceinline__ void warp_sort_shuffle_stage(unsigned long long (&k)[R], int lane, int xor_mask, int low_bit) {
    const bool lower = (lane & low_bit) == 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned long long p = __shfl_xor_sync(0xffffffffu, k[r], xor_mask);
        k[r] = lower ? u64_min(k[r], p) : u64_max(k[r], p);
    }
}

// The shuffle stages run as rolled loops (a fully unrolled network is ~14k SASS instructions and
// thrashes the instruction cache); only the three in-register merge levels are unrolled.
template <int R>
__device__ __forceinline__ void warp_sort_regs(unsigned long long (&k)[R], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 32; size <<= 1) {
        warp_sort_shuffle_stage<R>(k, lane, size - 1, size >> 1);  // flip: partner lane ^ (size-1)
#pragma unroll 1
        for (int j = size >> 2; j > 0; j >>= 1) warp_sort_shuffle_stage<R>(k, lane, j, j);
    }
#pragma unroll
    for (int m = 2; m <= R; m <<= 1) {  // size = 32 m: flip partner is register r ^ (m-1), lane ^ 31
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if ((r & (m >> 1)) == 0) {
                const int r2 = r ^ (m - 1);
                const unsigned long long a = k[r], b = k[r2];
                const unsigned long long pa = __shfl_xor_sync(0xffffffffu, b, 31);
                const unsigned long long pb = __shfl_xor_sync(0xffffffffu, a, 31);
                k[r] = u64_min(a, pa);
                k[r2] = u64_max(b, pb);
            }
        }
#pragma unroll
        for (int jr = m >> 2; jr > 0; jr >>= 1) {  // strides 32 jr: in-thread exchanges
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if ((r & jr) == 0) {
                    const unsigned long long a = k[r], b = k[r | jr];
                    k[r] = u64_min(a, b);
                    k[r | jr] = u64_max(a, b);
                }
            }
        }
#pragma unroll 1
        for (int j = 16; j > 0; j >>= 1) warp_sort_shuffle_stage<R>(k, lane, j, j);
    }
}

// ------------------------------------------------------------------ CTA-wide register sort
// 128 threads, R keys per thread, element e = r*128 + tid (up to 512 keys).  Strides below 32 are
// warp shuffles, strides 32..64 (and the flips that cross warps) go through a double-buffered
// shared-memory exchange (one barrier per layer), strides of 128 and above are in-thread.  Four
// warps share one tile's network, so the dependent chain per warp is ~4x shorter than a one-warp
// sort and four times as many warps are in flight.
template <int R>
struct CtaSort {
    unsigned long long (&k)[R];
    unsigned long long* sx;  // 2 * 128 * R keys
    int tid, buf;
    __device__ __forceinline__ CtaSort(unsigned long long (&k_)[R], unsigned long long* sx_, int tid_)
        : k(k_), sx(sx_), tid(tid_), buf(0) {}
    __device__ __forceinline__ void shfl_stage(int mask, int lowbit) {
        const bool lower = (tid & lowbit) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned long long p = __shfl_xor_sync(0xffffffffu, k[r], mask);
            k[r] = lower ? u64_min(k[r], p) : u64_max(k[r], p);
        }
    }
    __device__ __forceinline__ unsigned long long* publish() {
        unsigned long long* b = sx + buf * (kSortThreads * R);
        buf ^= 1;
#pragma unroll
        for (int r = 0; r < R; ++r) b[r * kSortThreads + tid] = k[r];
        __syncthreads();
        return b;
    }
    __device__ __forceinline__ void smem_stage(int mask, int lowbit) {
        const unsigned long long* b = publish();
        const bool lower = (tid & lowbit) == 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned long long p = b[r * kSortThreads + (tid ^ mask)];
            k[r] = lower ? u64_min(k[r], p) : u64_max(k[r], p);
        }
    }
    __device__ __forceinline__ void low_stages() {  // strides 16..1
#pragma unroll 1
        for (int j = 16; j > 0; j >>= 1) shfl_stage(j, j);
    }
    __device__ __forceinline__ void run() {
#pragma unroll 1
        for (int size = 2; size <= 32; size <<= 1) {
            shfl_stage(size - 1, size >> 1);
#pragma unroll 1
            for (int j = size >> 2; j > 0; j >>= 1) shfl_stage(j, j);
        }
        smem_stage(63, 32);  // size 64
        low_stages();
        smem_stage(127, 64);  // size 128
        smem_stage(32, 32);
        low_stages();
#pragma unroll
        for (int m = 2; m <= R; m <<= 1) {  // size 128 m: flip partner is register r ^ (m-1), thread ^ 127
            const unsigned long long* b = publish();
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const unsigned long long p = b[(r ^ (m - 1)) * kSortThreads + (tid ^ 127)];
                k[r] = ((r & (m >> 1)) == 0) ? u64_min(k[r], p) : u64_max(k[r], p);
            }
#pragma unroll
            for (int jr = m >> 2; jr > 0; jr >>= 1) {
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if ((r & jr) == 0) {
                        const unsigned long long x = k[r], y = k[r | jr];
                        k[r] = u64_min(x, y);
                        k[r | jr] = u64_max(x, y);
                    }
                }
            }
            smem_stage(64, 64);
            smem_stage(32, 32);
            low_stages();
        }
    }
};

template <int R, class Emit>
__device__ __forceinline__ void cta_sort_segment(const unsigned long long* __restrict__ keys, long long start, int n,
                                                 unsigned long long* sx, Emit emit) {
    const int tid = threadIdx.x;
    unsigned long long k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) k[r] = (r * kSortThreads + tid < n) ? keys[start + r * kSortThreads + tid] : ~0ull;
    CtaSort<R>(k, sx, tid).run();
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r * kSortThreads + tid < n) emit(start + r * kSortThreads + tid, k[r]);
}

constexpr int kCtaSortMax = 4 * kSortThreads;  // 512 keys: R = 4
constexpr int kSortSmemSmall = 1024;           // 8 KB: exchange buffers (2 x 512) or the generic smem path

// Sort one tile segment keys[start, start+n) with one warp (n <= 64) and hand every sorted key to
// emit(position, key).  All 32 lanes must call.
template <int R, class Emit>
__device__ __forceinline__ void warp_sort_segment(const unsigned long long* __restrict__ keys, long long start, int n,
                                                  int lane, Emit emit) {
    unsigned long long k[R];
#pragma unroll
    for (int r = 0; r < R; ++r) k[r] = (r * 32 + lane < n) ? keys[start + r * 32 + lane] : ~0ull;
    warp_sort_regs<R>(k, lane);
#pragma unroll
    for (int r = 0; r < R; ++r)
        if (r * 32 + lane < n) emit(start + r * 32 + lane, k[r]);
}

// One CTA of 128 threads per tile: one warp for <= 64 keys, the CTA-wide register network up to 512
// keys, a generic shared-memory bitonic pass up to 1024 keys and an in-place pass on the L2-resident
// global segment beyond.  `offsets` holds the exclusive scan of the (tile, replica) counters: tile t
// owns [offsets[t*R], offsets[(t+1)*R]).
template <class Emit>
__device__ __forceinline__ void sort_tile_cta(const int32_t* __restrict__ offsets, int R,
                                              unsigned long long* __restrict__ keys, long long capacity,
                                              unsigned long long* s_keys /* kSortSmemSmall */,
                                              int2* __restrict__ tile_range, Emit emit) {
    const int t = blockIdx.x;
    const int start = offsets[t * R];
    long long end = offsets[(t + 1) * R];
    if (end > capacity) end = max((long long)start, capacity);  // speculative capacity too small: host retries
    const int n = (int)(end - start);
    if (threadIdx.x == 0) tile_range[t] = (n > 0) ? make_int2(start, (int)end) : make_int2(0, 0);
    if (n <= 0) return;
    if (n <= 64) {
        if (threadIdx.x < 32) warp_sort_segment<2>(keys, start, n, threadIdx.x, emit);
    } else if (n <= kSortThreads) {
        cta_sort_segment<1>(keys, start, n, s_keys, emit);
    } else if (n <= 2 * kSortThreads) {
        cta_sort_segment<2>(keys, start, n, s_keys, emit);
    } else if (n <= kCtaSortMax) {
        cta_sort_segment<4>(keys, start, n, s_keys, emit);
    } else {
        int n_pad = 1024;
        while (n_pad < n) n_pad <<= 1;
        unsigned long long* buf = (n <= kSortSmemSmall) ? s_keys : (keys + start);
        if (n <= kSortSmemSmall) {
            for (int i = threadIdx.x; i < n; i += kSortThreads) s_keys[i] = keys[start + i];
            __syncthreads();
        }
        bitonic_sort_block(buf, n, n_pad);
        for (int i = threadIdx.x; i < n; i += kSortThreads) emit((long long)start + i, buf[i]);
    }
}

// ------------------------------------------------------------------ warp-expanded tile loops
// Each lane owns one Gaussian with a w x h tile rectangle at (x0, y0) (w*h == 0: none).  The warp
// walks the flattened list of (Gaussian, tile) pairs 32 at a time, so every lane issues one atomic
// per round regardless of how uneven the rectangles are.  A round yields, for lane `lane`, the
// owner lane `o` and the tile index (or tile < 0 when the round is ragged).
struct WarpTileWalk {
    int incl, cnt, total, w, x0, y0, gx, lane;
    __device__ __forceinline__ WarpTileWalk(int x0_, int y0_, int w_, int h_, int gx_, int lane_)
        : cnt(w_ * h_), w(w_), x0(x0_), y0(y0_), gx(gx_), lane(lane_) {
        incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        total = __shfl_sync(0xffffffffu, incl, 31);
    }
    // all lanes call; returns the tile for flat index base+lane (or -1) and the owner lane
    __device__ __forceinline__ int item(int base, int& owner) const {
        const int j = base + lane;
        int lo = 0;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const int v = __shfl_sync(0xffffffffu, incl, lo + s - 1);
            if (v <= j) lo += s;
        }
        lo = min(lo, 31);
        const int o_excl = __shfl_sync(0xffffffffu, incl - cnt, lo);
        const int o_w = __shfl_sync(0xffffffffu, w, lo);
        const int o_x0 = __shfl_sync(0xffffffffu, x0, lo);
        const int o_y0 = __shfl_sync(0xffffffffu, y0, lo);
        owner = lo;
        if (j >= total) return -1;
        const int local = j - o_excl;
        const int row = local / o_w;
        return (o_y0 + row) * gx + o_x0 + (local - row * o_w);
    }
};

// Tile counters are replicated R ways (replica = CTA index mod R): atomics on one address serialise
// in the L2 at roughly one per 50 cycles, and a 60k-Gaussian frame puts >100 of them on every tile.
#ifndef GFB_TILE_REPLICAS_MAX
#define GFB_TILE_REPLICAS_MAX 4
#endif
__host__ __device__ __forceinline__ int gfb_tile_replicas(int T) {
    int r = GFB_TILE_REPLICAS_MAX;
    while (r > 1 && (long long)T * r > 16384) r >>= 1;
    return r;
}

__device__ __forceinline__ void red_add_s32(int32_t* p, int v) {
    asm volatile("red.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Exclusive scan of n int32 counters by ONE CTA of kScanThreads threads, in rounds of 4096 counters
// staged through shared memory: coalesced global loads (one 128-byte line per warp instruction;
// per-thread contiguous chunks read straight from global would touch 16 lines per instruction),
// each thread scans 16 contiguous counters out of shared memory, the CTA scans the per-thread sums,
// and the offsets go back through shared memory with coalesced stores.  offsets[n] receives the
// total, which is returned to every thread.
constexpr int kScanThreads = 256;
constexpr int kScanPer = 16;
constexpr int kScanRound = kScanThreads * kScanPer;  // 4096 counters per round
constexpr int kScanPitch = kScanPer + 1;                // per-thread row pitch 17: conflict-free
constexpr int kScanSmemInts = kScanThreads * kScanPitch;

__device__ __forceinline__ int cta_exclusive_scan(const int32_t* __restrict__ counts, int n,
                                                  int32_t* __restrict__ offsets, int* s_buf /* kScanSmemInts */,
                                                  int* s_warp /* >= 34 ints */) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nw = kScanThreads / 32;
    if (tid == 0) s_warp[33] = 0;  // running carry
    __syncthreads();
    for (int round0 = 0; round0 < n; round0 += kScanRound) {
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) {
            const int e = k * kScanThreads + tid, i = round0 + e;
            // plain (coalescing) load: the counters were only ever touched by L2 atomics, so no SM holds
            // a stale L1 copy; ld.cg compiles to LDG.STRONG.GPU here, which does not coalesce (13 us for 13k)
            s_buf[(e >> 4) * kScanPitch + (e & 15)] = (i < n) ? counts[i] : 0;
        }
        __syncthreads();
        int c[kScanPer];
        int sum = 0;
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) {
            c[k] = s_buf[tid * kScanPitch + k];
            sum += c[k];
        }
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < nw) ? s_warp[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            s_warp[lane] = w;  // inclusive scan of the warp totals (entries >= nw repeat the total)
        }
        __syncthreads();
        const int carry = s_warp[33];
        int run = carry + (warp ? s_warp[warp - 1] : 0) + incl - sum;
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) {
            s_buf[tid * kScanPitch + k] = run;
            run += c[k];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kScanPer; ++k) {
            const int e = k * kScanThreads + tid, i = round0 + e;
            if (i < n) offsets[i] = s_buf[(e >> 4) * kScanPitch + (e & 15)];
        }
        if (tid == 0) s_warp[33] = carry + s_warp[31];
        __syncthreads();
    }
    const int total = s_warp[33];
    if (tid == 0) offsets[n] = total;
    return total;
}
