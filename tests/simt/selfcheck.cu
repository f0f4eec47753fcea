// selfcheck.cu -- tiny kernels that pin the semantics of the SIMT shim itself (tests/test_simt_selfcheck.py).
// Compiled ONLY against tests/simt/simt_emu.h; never part of libgflow_b200.so.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void k_warp_sum(const float* in, float* out) {
    float v = in[blockIdx.x * blockDim.x + threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

__global__ void k_scan_up(const int* in, int* out) {
    const int lane = threadIdx.x & 31;
    int v = in[threadIdx.x];
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    out[threadIdx.x] = v;
}

// lanes >= keep leave before the collectives: exited lanes count as arrived and contribute nothing
__global__ void k_partial_exit(int keep, unsigned* ballots, int* src7) {
    const int lane = threadIdx.x & 31;
    if (lane >= keep) return;
    const unsigned b = __ballot_sync(0xffffffffu, (lane & 1) == 0);
    const int any = __any_sync(0xffffffffu, lane == 3), all = __all_sync(0xffffffffu, lane < keep);
    const int from7 = __shfl_sync(0xffffffffu, lane * 10, 7);
    ballots[threadIdx.x] = b ^ ((unsigned)any << 30) ^ ((unsigned)all << 31);
    src7[threadIdx.x] = from7;
}

__global__ void k_barrier_count(const int* flags, int* out) {
    __shared__ int s_val[256];
    s_val[threadIdx.x] = flags[threadIdx.x];
    const int c = __syncthreads_count(flags[threadIdx.x]);
    // after the barrier every thread sees every other thread's shared write
    out[threadIdx.x] = c * 1000 + s_val[(threadIdx.x + 97) % blockDim.x];
}

__global__ void k_block_exit_before_barrier(int* out) {
    __shared__ int s_flag;
    if (threadIdx.x >= 64) return;  // two warps stay; the barrier must not wait for the exited ones
    if (threadIdx.x == 0) s_flag = 41;
    __syncthreads();
    out[threadIdx.x] = s_flag + 1;
}

// cp.async.bulk emulation: data must NOT be visible before somebody waits on the mbarrier
__global__ void k_bulk(const float* src, float* early, float* late, int wait_first) {
    __shared__ __attribute__((aligned(128))) float s_buf[64];
    __shared__ __attribute__((aligned(8))) uint64_t s_bar;
    if (threadIdx.x == 0) {
        gfb_emu::mbar_init(&s_bar, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        gfb_emu::mbar_expect_tx(&s_bar, 256);
        gfb_emu::bulk_g2s(s_buf, src, 256, &s_bar);
    }
    __syncthreads();
    if (wait_first) gfb_emu::mbar_wait(&s_bar, 0);
    early[threadIdx.x] = s_buf[threadIdx.x];
    gfb_emu::mbar_wait(&s_bar, 0);
    late[threadIdx.x] = s_buf[threadIdx.x];
}

__global__ void k_deadlock(int* out) {
    if (threadIdx.x == 5) return;
    if (threadIdx.x & 1) __syncthreads();  // only the odd threads ever reach the barrier
    else out[0] = __shfl_sync(0xffffffffu, 1, 0);  // ... the even ones wait for the odd lanes here
}

__global__ void k_grid_atomics(int* counter, float* fsum) {
    atomicAdd(counter, 1);
    atomicAdd(fsum, 0.5f);
    if (threadIdx.x == 0) atomicMax(counter + 1, (int)blockIdx.x + (int)blockIdx.y * 100);
}

}  // namespace

extern "C" {
void sc_warp_sum(const float* in, float* out, int blocks) { k_warp_sum<<<blocks, 64>>>(in, out); }
void sc_scan_up(const int* in, int* out) { k_scan_up<<<1, 32>>>(in, out); }
void sc_partial_exit(int keep, unsigned* ballots, int* src7) { k_partial_exit<<<1, 32>>>(keep, ballots, src7); }
void sc_barrier_count(const int* flags, int* out) { k_barrier_count<<<1, 256>>>(flags, out); }
void sc_block_exit(int* out) { k_block_exit_before_barrier<<<1, 256>>>(out); }
void sc_bulk(const float* src, float* early, float* late, int wait_first) { k_bulk<<<1, 64>>>(src, early, late, wait_first); }
void sc_deadlock(int* out) { k_deadlock<<<1, 32>>>(out); }
void sc_grid_atomics(int* counter, float* fsum) { k_grid_atomics<<<dim3(3, 2), 96>>>(counter, fsum); }
}
