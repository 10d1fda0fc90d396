// hostsim_warp.h — TEST HARNESS ONLY.  Lock-step emulation of one CTA on the host: every thread
// runs as a fiber (ucontext).  A warp collective (__shfl_sync, __shfl_xor_sync, __all_sync,
// __any_sync, __reduce_max_sync, __syncwarp) publishes the lane's value, yields to the scheduler
// and reads its partners' values once all 32 lanes of its warp have arrived; __syncthreads does the
// same for the whole CTA.  The exchange buffers alternate between consecutive collectives of a
// warp, so a lane that is one collective ahead does not overwrite what a slower lane still has to
// read.  Every collective carries a tag (kind and width); the emulator aborts when the lanes of a
// warp do not execute the same sequence of collectives — which checks the kernels' claim that
// their control flow around shuffles is warp-uniform — and when the CTA dead-locks (a barrier not
// reached by every thread).
//
// Between two collectives a thread runs ALONE, the threads one after another in ascending order,
// in descending order, or (order >= 2) in a pseudo-random order drawn anew for every pass of the
// scheduler from the seed `order`.  That is stricter than the hardware about the order of shared-memory
// accesses of different threads: a load that relies on another thread's earlier store (or must
// precede another thread's later store) with no barrier in between goes wrong in one of the two
// orders.  compute-sanitizer's racecheck is not available on the GPU pool of this project; this is
// the substitute for the kernels compiled into the harness.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <functional>
#include <memory>
#include <vector>

namespace hswarp {

enum { LANES = 32, STACK_BYTES = 512 << 10 };

struct Idx { int x; };

struct Cta {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::unique_ptr<char[]> stacks;               // not zero-filled: pages are touched on demand
    std::vector<char> done, wait;                 // wait: 0 runnable, 1 in a warp collective, 2 in __syncthreads
    std::vector<unsigned long long> count;        // warp collectives executed by each thread
    std::vector<double> xd;                       // [warp][2][LANES]
    std::vector<long long> xi;
    std::vector<int> tag;
    std::vector<int> warp_arrived;
    int cta_arrived;
    int nthreads, cur, block;
    int order;                                    // 0 ascending, 1 descending, >= 2: seed of a shuffled order per pass
    unsigned char *smem;
    const std::function<void(int)> *body;
};

inline Cta *&current() { static thread_local Cta *c = nullptr; return c; }

inline void fail(const char *what)
{
    fprintf(stderr, "hostsim_warp: %s\n", what);
    abort();
}

inline Cta *need()
{
    Cta *c = current();
    if (!c) fail("collective outside hswarp::run_cta");
    return c;
}

inline Idx thread_idx() { Idx i; i.x = need()->cur; return i; }
inline Idx block_idx() { Idx i; i.x = need()->block; return i; }
inline Idx block_dim() { Idx i; i.x = need()->nthreads; return i; }
inline unsigned char *dyn_smem() { return need()->smem; }
inline int lane_id() { return need()->cur % LANES; }

// publish in the warp's buffer, wait for the warp, return the index of the buffer to read
inline int arrive(int tag, double d, long long i)
{
    Cta *c = need();
    const int t = c->cur, w = t / LANES, l = t % LANES, ph = (int)(c->count[t] & 1);
    c->count[t] += 1;
    const int slot = (w * 2 + ph) * LANES;
    c->xd[slot + l] = d; c->xi[slot + l] = i; c->tag[slot + l] = tag;
    c->wait[t] = 1;
    if (++c->warp_arrived[w] == LANES) {
        c->warp_arrived[w] = 0;
        for (int o = 0; o < LANES; ++o) c->wait[w * LANES + o] = 0;
    }
    swapcontext(&c->ctx[t], &c->sched);
    for (int o = 0; o < LANES; ++o)
        if (c->tag[slot + o] != tag) fail("the lanes of a warp did not execute the same sequence of collectives");
    return slot;
}

inline void cta_barrier()
{
    Cta *c = need();
    const int t = c->cur;
    c->wait[t] = 2;
    if (++c->cta_arrived == c->nthreads) {
        c->cta_arrived = 0;
        for (int o = 0; o < c->nthreads; ++o) {
            if (c->wait[o] != 2) fail("__syncthreads reached while a thread waits in a warp collective");
            c->wait[o] = 0;
        }
    }
    swapcontext(&c->ctx[t], &c->sched);
}

inline void trampoline()
{
    Cta *c = current();
    const int t = c->cur;
    (*c->body)(t);
    c->done[t] = 1;
}

// run body(tid) for the nthreads (a multiple of 32) threads of one CTA in lock step
inline void run_cta(int nthreads, const std::function<void(int)> &body, int order = 0, int block = 0,
                    unsigned char *smem = nullptr)
{
    if (nthreads % LANES) fail("run_cta: the thread count must be a multiple of 32");
    Cta c;
    c.nthreads = nthreads; c.body = &body; c.order = order; c.block = block; c.smem = smem;
    c.ctx.resize(nthreads);
    c.stacks.reset(new char[(size_t)nthreads * STACK_BYTES]);
    c.done.assign(nthreads, 0); c.wait.assign(nthreads, 0); c.count.assign(nthreads, 0);
    const int nwarps = nthreads / LANES;
    c.xd.assign((size_t)nwarps * 2 * LANES, 0.0); c.xi.assign((size_t)nwarps * 2 * LANES, 0);
    c.tag.assign((size_t)nwarps * 2 * LANES, 0);
    c.warp_arrived.assign(nwarps, 0);
    c.cta_arrived = 0;
    Cta *outer = current();
    current() = &c;
    for (int t = 0; t < nthreads; ++t) {
        getcontext(&c.ctx[t]);
        c.ctx[t].uc_stack.ss_sp = c.stacks.get() + (size_t)t * STACK_BYTES;
        c.ctx[t].uc_stack.ss_size = STACK_BYTES;
        c.ctx[t].uc_link = &c.sched;
        makecontext(&c.ctx[t], (void (*)())trampoline, 0);
    }
    std::vector<int> perm(nthreads);
    for (int i = 0; i < nthreads; ++i) perm[i] = order == 1 ? nthreads - 1 - i : i;
    unsigned long long rng = 0x9e3779b97f4a7c15ull * (unsigned long long)(order + 1) + (unsigned long long)block;
    for (;;) {
        if (order >= 2)                               // Fisher-Yates with a xorshift generator
            for (int i = nthreads - 1; i > 0; --i) {
                rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
                const int j = (int)(rng % (unsigned long long)(i + 1));
                const int tmp = perm[i]; perm[i] = perm[j]; perm[j] = tmp;
            }
        bool ran = false, all_done = true;
        for (int i = 0; i < nthreads; ++i) {
            const int t = perm[i];
            if (c.done[t]) continue;
            all_done = false;
            if (c.wait[t]) continue;
            c.cur = t;
            swapcontext(&c.sched, &c.ctx[t]);
            ran = true;
        }
        if (all_done) break;
        if (!ran) fail("dead-lock: a barrier or warp collective was not reached by all of its threads");
    }
    current() = outer;
}

inline void run_warp(const std::function<void(int)> &body, int order = 0)
{
    run_cta(LANES, body, order);
}

}  // namespace hswarp

// ---- the CUDA built-ins the emulated device code uses, for the host build -----------------------
#define threadIdx (hswarp::thread_idx())
#define blockIdx (hswarp::block_idx())
#define blockDim (hswarp::block_dim())
static inline void __syncthreads() { hswarp::cta_barrier(); }

static inline double __shfl_sync(unsigned, double v, int src, int width = 32)
{
    const int l = hswarp::lane_id();
    const int slot = hswarp::arrive(0x100 + width, v, 0);
    return hswarp::current()->xd[slot + ((l & ~(width - 1)) | (src & (width - 1)))];
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask, int width = 32)
{
    const int l = hswarp::lane_id();
    const int slot = hswarp::arrive(0x200 + width, v, 0);
    const int o = l ^ lane_mask;
    return (o & ~(width - 1)) == (l & ~(width - 1)) ? hswarp::current()->xd[slot + o] : v;
}
static inline int __shfl_xor_sync(unsigned, int v, int lane_mask, int width = 32)
{
    const int l = hswarp::lane_id();
    const int slot = hswarp::arrive(0x280 + width, 0.0, v);
    const int o = l ^ lane_mask;
    return (o & ~(width - 1)) == (l & ~(width - 1)) ? (int)hswarp::current()->xi[slot + o] : v;
}
static inline int __shfl_sync(unsigned, int v, int src, int width = 32)
{
    const int l = hswarp::lane_id();
    const int slot = hswarp::arrive(0x180 + width, 0.0, v);
    return (int)hswarp::current()->xi[slot + ((l & ~(width - 1)) | (src & (width - 1)))];
}
static inline unsigned __shfl_sync(unsigned m, unsigned v, int src, int width = 32)
{
    return (unsigned)__shfl_sync(m, (int)v, src, width);
}
// mma.sync.aligned.m8n8k4.row.col.f64: D (8 x 8) += A (8 x 4) B (4 x 8) with the PTX fragments
// a = A[lane >> 2][lane & 3], b = B[lane & 3][lane >> 2], d0 / d1 = D[lane >> 2][2 (lane & 3) + 0 / 1]
static inline void hs_dmma_m8n8k4(double &d0, double &d1, double a, double b)
{
    long long bits;
    static_assert(sizeof(bits) == sizeof(b), "double is 64 bits");
    __builtin_memcpy(&bits, &b, sizeof(bits));
    const int l = hswarp::lane_id();
    const int slot = hswarp::arrive(0x700, a, bits);
    const int row = l >> 2, c0 = 2 * (l & 3);
    const hswarp::Cta *c = hswarp::current();
    for (int k = 0; k < 4; ++k) {
        const double ak = c->xd[slot + row * 4 + k];
        double b0, b1;
        __builtin_memcpy(&b0, &c->xi[slot + c0 * 4 + k], sizeof(b0));
        __builtin_memcpy(&b1, &c->xi[slot + (c0 + 1) * 4 + k], sizeof(b1));
        d0 = __builtin_fma(ak, b0, d0);
        d1 = __builtin_fma(ak, b1, d1);
    }
}
static inline int __all_sync(unsigned, int pred)
{
    const int slot = hswarp::arrive(0x300, 0.0, pred != 0);
    int r = 1;
    for (int o = 0; o < hswarp::LANES; ++o) r &= (int)hswarp::current()->xi[slot + o];
    return r;
}
static inline int __any_sync(unsigned, int pred)
{
    const int slot = hswarp::arrive(0x400, 0.0, pred != 0);
    int r = 0;
    for (int o = 0; o < hswarp::LANES; ++o) r |= (int)hswarp::current()->xi[slot + o];
    return r;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    const int slot = hswarp::arrive(0x500, 0.0, v);
    long long r = hswarp::current()->xi[slot];
    for (int o = 1; o < hswarp::LANES; ++o) r = hswarp::current()->xi[slot + o] > r ? hswarp::current()->xi[slot + o] : r;
    return (int)r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { hswarp::arrive(0x600, 0.0, 0); }
