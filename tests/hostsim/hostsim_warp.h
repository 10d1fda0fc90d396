// hostsim_warp.h — TEST HARNESS ONLY.  Lock-step emulation of one warp on the host: the 32 lanes
// run as 32 fibers (ucontext); a warp collective (__shfl_sync, __shfl_xor_sync, __all_sync,
// __any_sync, __reduce_max_sync, __syncwarp) publishes the lane's value, yields to the scheduler
// and reads its partners' values once every lane has arrived.  The exchange buffers alternate
// between consecutive collectives, so a lane that is one collective ahead does not overwrite what
// a slower lane still has to read.  Every collective carries a tag (kind and width); the emulator
// aborts when the lanes of a warp do not execute the same sequence of collectives — which checks
// the kernels' claim that their control flow around shuffles is warp-uniform.
//
// Between two collectives a lane runs ALONE, the lanes one after another in ascending or (second
// argument of run_warp) descending order.  That is stricter than the hardware about the order of
// shared-memory accesses of different lanes: a load that relies on another lane's earlier store
// (or must precede another lane's later store) with no collective in between goes wrong in one
// of the two orders.
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <ucontext.h>

#include <functional>
#include <vector>

namespace hswarp {

enum { LANES = 32, STACK_BYTES = 1 << 20 };

struct Warp {
    ucontext_t sched, ctx[LANES];
    bool done[LANES];
    int lane;                       // the lane that is running
    bool descending;                // lanes are resumed 31 .. 0 instead of 0 .. 31
    unsigned long long count[LANES];  // collectives executed by each lane
    double xd[2][LANES];
    long long xi[2][LANES];
    int tag[2][LANES];
    const std::function<void(int)> *body;
    std::vector<char> stacks;
};

inline Warp *&current() { static thread_local Warp *w = nullptr; return w; }

inline void fail(const char *what)
{
    fprintf(stderr, "hostsim_warp: %s\n", what);
    abort();
}

// publish, wait for the whole warp, return the buffer index to read from
inline int arrive(int tag, double d, long long i)
{
    Warp *w = current();
    if (!w) fail("warp collective outside hswarp::run_warp");
    const int l = w->lane, ph = (int)(w->count[l] & 1);
    w->count[l] += 1;
    w->xd[ph][l] = d; w->xi[ph][l] = i; w->tag[ph][l] = tag;
    swapcontext(&w->ctx[l], &w->sched);
    // resumed: every lane has published.  Lanes resumed before l have already run on to their
    // next collective (or to the end); the others still wait in this one.
    for (int o = 0; o < LANES; ++o) {
        const bool ahead = w->descending ? o > l : o < l;
        const bool ok = ahead ? (w->done[o] ? w->count[o] == w->count[l] : w->count[o] == w->count[l] + 1)
                              : (!w->done[o] && w->count[o] == w->count[l]);
        if (!ok || w->tag[ph][o] != tag)
            fail("the lanes of a warp did not execute the same sequence of collectives");
    }
    return ph;
}

inline void trampoline()
{
    Warp *w = current();
    const int l = w->lane;
    (*w->body)(l);
    w->done[l] = true;
}

// run body(lane) for the 32 lanes of one warp in lock step
inline void run_warp(const std::function<void(int)> &body, bool descending = false)
{
    Warp w;
    w.body = &body;
    w.descending = descending;
    w.stacks.resize((size_t)LANES * STACK_BYTES);
    Warp *outer = current();
    current() = &w;
    for (int l = 0; l < LANES; ++l) {
        w.done[l] = false; w.count[l] = 0;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stacks.data() + (size_t)l * STACK_BYTES;
        w.ctx[l].uc_stack.ss_size = STACK_BYTES;
        w.ctx[l].uc_link = &w.sched;
        makecontext(&w.ctx[l], (void (*)())trampoline, 0);
    }
    for (bool any = true; any;) {
        any = false;
        for (int i = 0; i < LANES; ++i) {
            const int l = descending ? LANES - 1 - i : i;
            if (w.done[l]) continue;
            w.lane = l;
            swapcontext(&w.sched, &w.ctx[l]);
            any = true;
        }
        bool first = w.done[0];
        for (int l = 1; l < LANES; ++l)
            if (w.done[l] != first) fail("some lanes of a warp finished while others wait in a collective");
    }
    current() = outer;
}

}  // namespace hswarp

// ---- the CUDA warp intrinsics the K1p device code uses, for the host build ---------------------
static inline double __shfl_sync(unsigned, double v, int src, int width = 32)
{
    const int l = hswarp::current() ? hswarp::current()->lane : 0;
    const int ph = hswarp::arrive(0x100 + width, v, 0);
    return hswarp::current()->xd[ph][(l & ~(width - 1)) | (src & (width - 1))];
}
static inline double __shfl_xor_sync(unsigned, double v, int lane_mask, int width = 32)
{
    const int l = hswarp::current() ? hswarp::current()->lane : 0;
    const int ph = hswarp::arrive(0x200 + width, v, 0);
    const int o = l ^ lane_mask;
    return (o & ~(width - 1)) == (l & ~(width - 1)) ? hswarp::current()->xd[ph][o] : v;
}
static inline int __all_sync(unsigned, int pred)
{
    const int ph = hswarp::arrive(0x300, 0.0, pred != 0);
    int r = 1;
    for (int o = 0; o < hswarp::LANES; ++o) r &= (int)hswarp::current()->xi[ph][o];
    return r;
}
static inline int __any_sync(unsigned, int pred)
{
    const int ph = hswarp::arrive(0x400, 0.0, pred != 0);
    int r = 0;
    for (int o = 0; o < hswarp::LANES; ++o) r |= (int)hswarp::current()->xi[ph][o];
    return r;
}
static inline int __reduce_max_sync(unsigned, int v)
{
    const int ph = hswarp::arrive(0x500, 0.0, v);
    long long r = hswarp::current()->xi[ph][0];
    for (int o = 1; o < hswarp::LANES; ++o) r = hswarp::current()->xi[ph][o] > r ? hswarp::current()->xi[ph][o] : r;
    return (int)r;
}
static inline void __syncwarp(unsigned = 0xffffffffu) { hswarp::arrive(0x600, 0.0, 0); }
