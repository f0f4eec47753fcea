// simt_emu.cpp -- fiber scheduler behind simt_emu.h (TEST-ONLY; see the header).
#include "simt_emu.h"

#include <sys/mman.h>

#include <unordered_map>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned long long gfb_emu_counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};

extern "C" void gfb_emu_stats(unsigned long long* out) {
    for (int i = 0; i < 8; ++i) out[i] = gfb_emu_counts[i];
}
extern "C" void gfb_emu_stats_reset() {
    for (int i = 0; i < 8; ++i) gfb_emu_counts[i] = 0;
}

[[noreturn]] void gfb_emu_fail(const char* what) {
    fprintf(stderr, "simt_emu: %s (block %u,%u,%u thread %u,%u,%u)\n", what, blockIdx.x, blockIdx.y, blockIdx.z,
            threadIdx.x, threadIdx.y, threadIdx.z);
    fflush(stderr);
    abort();
}

// ------------------------------------------------------------------ context switch
#if defined(__x86_64__)
extern "C" void gfb_emu_switch(void** save_sp, void* new_sp);
asm(R"(
.text
.globl gfb_emu_switch
.type gfb_emu_switch,@function
gfb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size gfb_emu_switch,.-gfb_emu_switch
)");
#else
#error "simt_emu: only x86-64 hosts are supported"
#endif

namespace gfb_emu {
namespace {

constexpr size_t kStackBytes = 256 * 1024;
constexpr int kMaxThreads = 1024;

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    bool done = true;
    uint3 tid{0, 0, 0};
    int warp = 0, lane = 0;
};

struct Warp {
    uint64_t slot[32];
    uint64_t res[2][32];
    unsigned res_part[2];
    unsigned arrived = 0, alive = 0, want = 0, gen = 0;
};

struct Bar {
    unsigned phase = 0;
    int pending = 0, init = 0;
    long long tx = 0;
    struct Copy { void* dst; const void* src; unsigned bytes; };
    std::vector<Copy> copies;
};

struct Cta {
    int nthreads = 0, alive = 0;
    int bar_arrived = 0, bar_count = 0, bar_or = 0, bar_and = 1;
    unsigned bar_gen = 0;
    int bar_res_count[2], bar_res_or[2], bar_res_and[2];
    Warp warps[kMaxThreads / 32];
    std::unordered_map<void*, Bar> mbars;
} g_cta;

Fiber g_fibers[kMaxThreads];
Fiber* g_cur = nullptr;
void* g_sched_sp = nullptr;
const std::function<void()>* g_body = nullptr;
unsigned long long g_progress = 0;
// fiber sweep order: 0 = ascending thread index, 1 = descending, 2 = pseudo-random permutation per sweep
int g_sched_mode = 0;
unsigned g_sched_state = 12345u;

void yield() {
    Fiber* f = g_cur;
    gfb_emu_switch(&f->sp, g_sched_sp);
    // resumed: the scheduler has restored threadIdx and g_cur
}

void try_complete_warp(Warp& w) {
    const unsigned need = w.want & w.alive;
    if (w.arrived == 0 || (w.arrived & need) != need) return;
    const unsigned g = w.gen & 1u;
    memcpy(w.res[g], w.slot, sizeof(w.slot));
    w.res_part[g] = w.arrived;
    w.arrived = 0;
    w.want = 0;
    ++w.gen;
    ++g_progress;
    ++gfb_emu_counts[0];
}

void try_complete_barrier() {
    Cta& c = g_cta;
    if (c.bar_arrived == 0 || c.bar_arrived < c.alive) return;
    const unsigned g = c.bar_gen & 1u;
    c.bar_res_count[g] = c.bar_count;
    c.bar_res_or[g] = c.bar_or;
    c.bar_res_and[g] = c.bar_and;
    c.bar_arrived = 0;
    c.bar_count = 0;
    c.bar_or = 0;
    c.bar_and = 1;
    ++c.bar_gen;
    ++g_progress;
    ++gfb_emu_counts[2];
}

void fiber_exit() {
    Fiber* f = g_cur;
    f->done = true;
    Warp& w = g_cta.warps[f->warp];
    w.alive &= ~(1u << f->lane);
    --g_cta.alive;
    try_complete_warp(w);
    try_complete_barrier();
    ++g_progress;
    void* dummy;
    gfb_emu_switch(&dummy, g_sched_sp);
    gfb_emu_fail("resumed a finished fiber");
}

void fiber_entry() {
    (*g_body)();
    fiber_exit();
}

void prepare(Fiber& f) {
    if (!f.stack) {
        void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) gfb_emu_fail("mmap of a fiber stack failed");
        f.stack = (char*)p;
    }
    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // fake return address of fiber_entry
    *--sp = (void*)&fiber_entry;     // `ret` of the first switch jumps here
    for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = sp;
    f.done = false;
}

void run_cta(dim3 block) {
    Cta& c = g_cta;
    const int n = (int)(block.x * block.y * block.z);
    if (n <= 0 || n > kMaxThreads) gfb_emu_fail("bad block size");
    c.nthreads = c.alive = n;
    ++gfb_emu_counts[4];
    gfb_emu_counts[5] += (unsigned long long)n;
    c.bar_arrived = c.bar_count = c.bar_or = 0;
    c.bar_and = 1;
    c.bar_gen = 0;
    c.mbars.clear();
    const int nwarps = (n + 31) / 32;
    for (int w = 0; w < nwarps; ++w) {
        Warp& W = c.warps[w];
        W.arrived = W.want = 0;
        W.gen = 0;
        const int lanes = std::min(32, n - 32 * w);
        W.alive = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
    }
    for (int t = 0; t < n; ++t) {
        Fiber& f = g_fibers[t];
        prepare(f);
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        f.warp = t >> 5;
        f.lane = t & 31;
    }
    int remaining = n;
    int idle_sweeps = 0;
    while (remaining > 0) {
        const unsigned long long before = g_progress;
        // a correct kernel gives the same results under every interleaving of its threads
        unsigned mul = 1, add = 0;
        if (g_sched_mode == 2) {
            g_sched_state = g_sched_state * 1664525u + 1013904223u;
            add = g_sched_state >> 8;
            mul = 2 * (g_sched_state >> 20) + 1;
        }
        int pow2 = 1;
        while (pow2 < n) pow2 <<= 1;
        for (int i = 0; i < pow2; ++i) {
            int t = g_sched_mode == 1 ? n - 1 - i : (g_sched_mode == 2 ? (int)((i * mul + add) & (unsigned)(pow2 - 1)) : i);
            if (t < 0 || t >= n) continue;
            Fiber& f = g_fibers[t];
            if (f.done) continue;
            g_cur = &f;
            threadIdx = f.tid;
            gfb_emu_switch(&g_sched_sp, f.sp);
            if (f.done) --remaining;
        }
        if (g_progress == before) {
            if (++idle_sweeps > 4) gfb_emu_fail("deadlock: no fiber made progress (barrier / collective / mbarrier never completes)");
        } else {
            idle_sweeps = 0;
        }
    }
    g_cur = nullptr;
}

Bar& bar_of(void* p) {
    auto it = g_cta.mbars.find(p);
    if (it == g_cta.mbars.end()) gfb_emu_fail("mbarrier used before mbarrier.init");
    return it->second;
}

void bar_check(Bar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.init;
        ++g_progress;
    }
}

void bar_flush(Bar& b) {
    for (const Bar::Copy& c : b.copies) {
        memcpy(c.dst, c.src, c.bytes);
        b.tx -= c.bytes;
    }
    if (!b.copies.empty()) {
        b.copies.clear();
        bar_check(b);
    }
}

}  // namespace

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    if (g_cur) gfb_emu_fail("nested launch");
    const std::function<void()>* prev = g_body;
    g_body = &body;
    gridDim = grid;
    blockDim = block;
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                blockIdx = uint3{x, y, z};
                run_cta(block);
            }
    g_body = prev;
}

int lane_id() { return g_cur->lane; }

extern "C" void gfb_emu_set_schedule(int mode, unsigned seed) {
    g_sched_mode = mode;
    g_sched_state = seed * 2654435761u + 1u;
}

const uint64_t* warp_exchange(uint64_t v, unsigned mask, unsigned* part) {
    Fiber* f = g_cur;
    Warp& w = g_cta.warps[f->warp];
    const unsigned bit = 1u << f->lane;
    if (!(mask & bit)) gfb_emu_fail("warp collective: calling lane is not in the mask");
    if (w.arrived & bit) gfb_emu_fail("warp collective: lane arrived twice");
    if (w.arrived && w.want != mask) gfb_emu_fail("warp collective: lanes disagree on the mask");
    w.want = mask;
    w.slot[f->lane] = v;
    w.arrived |= bit;
    const unsigned g = w.gen;
    try_complete_warp(w);
    while (w.gen == g) yield();
    *part = w.res_part[g & 1u];
    return w.res[g & 1u];
}

int cta_barrier(int pred, int* or_out, int* and_out) {
    Cta& c = g_cta;
    ++c.bar_arrived;
    c.bar_count += pred ? 1 : 0;
    c.bar_or |= pred ? 1 : 0;
    c.bar_and &= pred ? 1 : 0;
    const unsigned g = c.bar_gen;
    try_complete_barrier();
    while (c.bar_gen == g) yield();
    if (or_out) *or_out = c.bar_res_or[g & 1u];
    if (and_out) *and_out = c.bar_res_and[g & 1u];
    return c.bar_res_count[g & 1u];
}

void mbar_init(void* bar, unsigned count) {
    Bar& b = g_cta.mbars[bar];
    b = Bar();
    b.pending = b.init = (int)count;
}

void mbar_arrive(void* bar) {
    Bar& b = bar_of(bar);
    if (b.pending <= 0) gfb_emu_fail("mbarrier: more arrivals than the init count");
    --b.pending;
    bar_check(b);
}

void mbar_expect_tx(void* bar, unsigned bytes) {
    Bar& b = bar_of(bar);
    if (b.pending <= 0) gfb_emu_fail("mbarrier: more arrivals than the init count");
    b.tx += bytes;
    --b.pending;
    bar_check(b);
}

void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    if (((uintptr_t)dst & 15u) || ((uintptr_t)src & 15u) || (bytes & 15u) || bytes == 0)
        gfb_emu_fail("cp.async.bulk: dst, src and size must be non-zero multiples of 16 bytes");
    Bar& b = bar_of(bar);
    gfb_emu_counts[3] += bytes;
    memset(dst, 0xff, bytes);  // poison: the data is NOT there until somebody waits on the mbarrier
    b.copies.push_back(Bar::Copy{dst, src, bytes});
}

void mbar_wait(void* bar, unsigned parity) {
    Bar& b = bar_of(bar);
    bar_flush(b);
    while ((b.phase & 1u) == (parity & 1u)) {
        yield();
        bar_flush(bar_of(bar));
    }
}

}  // namespace gfb_emu
