// hostpipe.cu -- stream plumbing of a render step whose inputs and results live in HOST memory
// (gflow_b200.hostapi.HostRenderStep; include/gflow_b200.h "host pipe").
//
// One step = H2D copy of the inputs, the compute (a CUDA graph the caller captured over the slot's device
// blocks), D2H copy of the results.  The three run on three streams -- two copy streams owned by the pipe (PCIe is
// full duplex and the GPU has a copy engine per direction) and the caller's compute stream -- and `depth` slots are
// in flight, so the H2D of step i+1 and the D2H of step i-1 overlap the kernels of step i.  All of it is one C call
// per step: driven from Python, the ten stream / event / copy / launch calls of a step cost more host time
// (~130 us) than the step's kernels (~105 us at BASELINE config 2).
#include "common.cuh"

#include <new>
#include <vector>

struct gfb_hostpipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    struct Slot {
        cudaEvent_t in_ready = nullptr, done = nullptr, out_landed = nullptr;
        bool busy = false;
    };
    std::vector<Slot> slots;
};

extern "C" {

int gfb_hostpipe_create(int depth, gfb_hostpipe** out) {
    if (depth < 1 || depth > 64 || !out) return GFB_E_BADARG;
    gfb_hostpipe* p = new (std::nothrow) gfb_hostpipe;
    if (!p) return GFB_E_UNSUPPORTED;
    p->slots.resize((size_t)depth);
    int rc = (int)cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking);
    if (!rc) rc = (int)cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking);
    for (auto& s : p->slots) {
        if (!rc) rc = (int)cudaEventCreateWithFlags(&s.in_ready, cudaEventDisableTiming);
        if (!rc) rc = (int)cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming);
        if (!rc) rc = (int)cudaEventCreateWithFlags(&s.out_landed, cudaEventDisableTiming);
    }
    if (rc) {
        gfb_hostpipe_destroy(p);
        return rc;
    }
    *out = p;
    return 0;
}

int gfb_hostpipe_destroy(gfb_hostpipe* p) {
    if (!p) return 0;
    for (auto& s : p->slots) {
        if (s.in_ready) cudaEventDestroy(s.in_ready);
        if (s.done) cudaEventDestroy(s.done);
        if (s.out_landed) cudaEventDestroy(s.out_landed);
    }
    if (p->h2d) cudaStreamDestroy(p->h2d);
    if (p->d2h) cudaStreamDestroy(p->d2h);
    delete p;
    return 0;
}

int gfb_hostpipe_submit(gfb_hostpipe* p, int slot, void* dev_in, const void* host_in, size_t in_bytes, void* graph_exec,
                        void* compute_stream, void* host_out, const void* dev_out, size_t out_bytes) {
    if (!p || slot < 0 || slot >= (int)p->slots.size() || !graph_exec) return GFB_E_BADARG;
    if ((in_bytes && (!dev_in || !host_in)) || (out_bytes && (!host_out || !dev_out))) return GFB_E_BADARG;
    gfb_hostpipe::Slot& s = p->slots[(size_t)slot];
    cudaStream_t cs = (cudaStream_t)compute_stream;
    // dev_in is still read by the slot's previous compute; dev_out must not be overwritten before its D2H has run
    if (s.busy) GFB_TRY(cudaStreamWaitEvent(p->h2d, s.done, 0));
    if (in_bytes) GFB_TRY(cudaMemcpyAsync(dev_in, host_in, in_bytes, cudaMemcpyHostToDevice, p->h2d));
    GFB_TRY(cudaEventRecord(s.in_ready, p->h2d));
    GFB_TRY(cudaStreamWaitEvent(cs, s.in_ready, 0));
    if (s.busy) GFB_TRY(cudaStreamWaitEvent(cs, s.out_landed, 0));
    GFB_TRY(cudaGraphLaunch((cudaGraphExec_t)graph_exec, cs));
    GFB_TRY(cudaEventRecord(s.done, cs));
    GFB_TRY(cudaStreamWaitEvent(p->d2h, s.done, 0));
    if (out_bytes) GFB_TRY(cudaMemcpyAsync(host_out, dev_out, out_bytes, cudaMemcpyDeviceToHost, p->d2h));
    GFB_TRY(cudaEventRecord(s.out_landed, p->d2h));
    s.busy = true;
    return 0;
}

int gfb_hostpipe_wait(gfb_hostpipe* p) {
    if (!p) return GFB_E_BADARG;
    for (auto& s : p->slots)
        if (s.busy) GFB_TRY(cudaEventSynchronize(s.out_landed));
    return 0;
}

}  // extern "C"
