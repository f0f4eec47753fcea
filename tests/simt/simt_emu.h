// simt_emu.h -- TEST-ONLY host shim that lets g++ compile gflow_b200/csrc/*.cu and execute the kernels'
// SIMT logic on a CPU, so `pytest -m "not gpu"` can check kernel indexing / synchronisation / warp
// collectives / TMA staging against the oracle in a container that has no GPU.
//
// This is NOT a CPU fallback and is never built, imported or linked by the gflow_b200 package: only
// tests/simt/build_emu.py compiles it (into tests/simt/_build/), only tests/ load the result.
//
// Execution model
//  * a launch runs its CTAs one after the other on the calling OS thread;
//  * every CUDA thread of a CTA is a fiber (own stack, hand-rolled context switch); a fiber runs until
//    it reaches a CTA barrier, a warp collective or an mbarrier wait it cannot pass, then yields;
//  * warp collectives (__shfl*_sync, __ballot_sync, ...) complete when every live lane named by the
//    mask has arrived; exited lanes count as arrived (as on Volta+ hardware);
//  * `__shared__` variables are function-local statics (CTAs never overlap in time);
//  * cp.async.bulk copies are POISONED at issue and performed lazily when some thread first waits on
//    their mbarrier, so reading a stage without waiting, or refilling one that is still being read,
//    shows up as NaNs / wrong results; size and alignment rules of the real instruction are asserted;
//  * a sweep over all fibers that makes no progress aborts with "deadlock".
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <type_traits>

#define GFB_SIMT_EMU 1

// ------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static
#define __constant__ static const

// ------------------------------------------------------------------ vector types
struct alignas(8) float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) ushort4 { unsigned short x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) {
    return ushort4{x, y, z, w};
}

extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
static const int warpSize = 32;

// ------------------------------------------------------------------ scalar builtins
template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)b < (T)a ? (T)b : (T)a;
}
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)b : (T)a;
}
static inline float __fmul_rn(float a, float b) { return a * b; }  // built with -ffp-contract=off
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
// __expf / __logf: glibc exports functions of exactly these names (float precision), used as they are
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __saturatef(float x) { return x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
[[noreturn]] void gfb_emu_fail(const char* what);
static inline void __trap() { gfb_emu_fail("__trap()"); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline size_t __cvta_generic_to_global(const void* p) { return (size_t)p; }

// dynamic operation counts of everything launched since the last reset (gfb_emu_stats / gfb_emu_stats_reset):
// [0] warp collectives (one per warp-level shuffle / vote / reduce), [1] lane-level atomics, [2] CTA barriers (one per
// CTA-wide barrier), [3] bytes moved by emulated cp.async.bulk, [4] CTAs run, [5] threads run
extern unsigned long long gfb_emu_counts[8];

// atomics: one OS thread, fibers switch only at collectives, so plain read-modify-write is atomic
template <class T>
static inline T atomicAdd(T* p, T v) { ++gfb_emu_counts[1]; T o = *p; *p = o + v; return o; }
static inline float atomicAdd(float* p, float v) { ++gfb_emu_counts[1]; float o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { ++gfb_emu_counts[1]; int o = *p; *p = o + v; return o; }
static inline int atomicSub(int* p, int v) { ++gfb_emu_counts[1]; int o = *p; *p = o - v; return o; }
static inline int atomicMax(int* p, int v) { ++gfb_emu_counts[1]; int o = *p; if (v > o) *p = v; return o; }
static inline int atomicMin(int* p, int v) { ++gfb_emu_counts[1]; int o = *p; if (v < o) *p = v; return o; }
static inline int atomicExch(int* p, int v) { ++gfb_emu_counts[1]; int o = *p; *p = v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { ++gfb_emu_counts[1]; unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicMin(unsigned* p, unsigned v) { ++gfb_emu_counts[1]; unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { ++gfb_emu_counts[1]; unsigned o = *p; if (v > o) *p = v; return o; }

// ------------------------------------------------------------------ scheduler interface
namespace gfb_emu {
void launch(dim3 grid, dim3 block, const std::function<void()>& body);
// deposits `v` for the calling lane, blocks until every live lane in `mask` has arrived; returns the 32
// deposited values and (in *part) the set of lanes that took part
const uint64_t* warp_exchange(uint64_t v, unsigned mask, unsigned* part);
int lane_id();
int cta_barrier(int pred, int* or_out, int* and_out);  // returns the count of non-zero pred
// mbarrier + bulk copy
void mbar_init(void* bar, unsigned count);
void mbar_expect_tx(void* bar, unsigned bytes);
void mbar_arrive(void* bar);
void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar);
void mbar_wait(void* bar, unsigned parity);
}  // namespace gfb_emu

static inline void __syncthreads() { gfb_emu::cta_barrier(0, nullptr, nullptr); }
static inline int __syncthreads_count(int pred) { return gfb_emu::cta_barrier(pred, nullptr, nullptr); }
static inline int __syncthreads_or(int pred) { int o; gfb_emu::cta_barrier(pred, &o, nullptr); return o; }
static inline int __syncthreads_and(int pred) { int a; gfb_emu::cta_barrier(pred, nullptr, &a); return a; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { unsigned p; gfb_emu::warp_exchange(0, mask, &p); }

namespace gfb_emu {
template <class T>
static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
static inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
}  // namespace gfb_emu

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    unsigned part;
    const int lane = gfb_emu::lane_id();
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    const int base = lane & ~(width - 1);
    const int from = base + (src & (width - 1));
    return ((part >> from) & 1u) ? gfb_emu::from_bits<T>(s[from]) : v;
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    unsigned part;
    const int lane = gfb_emu::lane_id();
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    const int from = lane ^ lanemask;
    if ((from & ~(width - 1)) != (lane & ~(width - 1))) return v;
    return ((part >> from) & 1u) ? gfb_emu::from_bits<T>(s[from]) : v;
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned part;
    const int lane = gfb_emu::lane_id();
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    const int from = lane - (int)delta;
    if (from < (lane & ~(width - 1))) return v;
    return ((part >> from) & 1u) ? gfb_emu::from_bits<T>(s[from]) : v;
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    unsigned part;
    const int lane = gfb_emu::lane_id();
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    const int from = lane + (int)delta;
    if (from > (lane | (width - 1))) return v;
    return ((part >> from) & 1u) ? gfb_emu::from_bits<T>(s[from]) : v;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned part, r = 0;
    const uint64_t* s = gfb_emu::warp_exchange(pred ? 1 : 0, mask, &part);
    for (int l = 0; l < 32; ++l)
        if (((part >> l) & 1u) && s[l]) r |= 1u << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __activemask() { unsigned part; gfb_emu::warp_exchange(0, 0xffffffffu, &part); return part; }
static inline int __reduce_max_sync(unsigned mask, int v) {
    unsigned part;
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    int r = v;
    for (int l = 0; l < 32; ++l)
        if ((part >> l) & 1u) r = std::max(r, gfb_emu::from_bits<int>(s[l]));
    return r;
}
static inline int __reduce_min_sync(unsigned mask, int v) {
    unsigned part;
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    int r = v;
    for (int l = 0; l < 32; ++l)
        if ((part >> l) & 1u) r = std::min(r, gfb_emu::from_bits<int>(s[l]));
    return r;
}
static inline int __reduce_add_sync(unsigned mask, int v) {
    unsigned part;
    const uint64_t* s = gfb_emu::warp_exchange(gfb_emu::to_bits(v), mask, &part);
    int r = 0;
    for (int l = 0; l < 32; ++l)
        if ((part >> l) & 1u) r += gfb_emu::from_bits<int>(s[l]);
    return r;
}

// ------------------------------------------------------------------ CUDA runtime subset (host side)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorNotReady = 600 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaHostAllocMapped = 2, cudaEventDisableTiming = 2 };
enum { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { int id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t {
    dim3 gridDim, blockDim;
    size_t dynamicSmemBytes;
    cudaStream_t stream;
    cudaLaunchAttribute* attrs;
    unsigned numAttrs;
};
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = calloc(1, n); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
static inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* s) { *s = cudaStreamCaptureStatusNone; return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
// host pipe (hostpipe.cu): streams and events are tokens, copies happen at once, a "graph" is a host callback
typedef void (*cudaGraphExec_t)(void);
enum { cudaStreamNonBlocking = 1, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t g, cudaStream_t) { g(); return cudaSuccess; }

template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args... args) {
    gfb_emu::launch(cfg->gridDim, cfg->blockDim, [&]() { kernel(static_cast<KArgs>(args)...); });
    return cudaSuccess;
}
// `kernel<<<grid, block, smem, stream>>>(args)` is rewritten by build_emu.py into
// GFB_EMU_LAUNCH(kernel, (grid), (block), args)
#define GFB_EMU_LAUNCH(kernel, grid, block, ...) \
    gfb_emu::launch(dim3 grid, dim3 block, [&]() { kernel(__VA_ARGS__); })
