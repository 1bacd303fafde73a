// simt.h -- TEST INFRASTRUCTURE.  A one-warp SIMT emulator for the host: the 32 lanes of a warp are 32 fibers
// (ucontext) on one OS thread; every warp-collective (__ballot_sync, __shfl_sync, __syncwarp, ...) is a rendezvous
// point at which a lane parks until all 32 have arrived.  Divergent code between two collectives simply runs lane by
// lane.  A collective that not every lane reaches (e.g. one hidden behind a short-circuit `&&`) shows up as a
// deadlock and aborts the test -- on the GPU the same bug hangs the kernel.  Used by twin_kernel.cpp to run the real
// qr::k_step on a machine without a GPU.  Nothing in the package loads this.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <algorithm>

struct tw_uint3 { unsigned x, y, z; };
static tw_uint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {32, 1, 1}, gridDim = {1, 1, 1};

namespace simt {

constexpr int LANES = 32;
constexpr size_t STACK = 1 << 20;

struct Warp {
    ucontext_t main_ctx, ctx[LANES];
    char* stacks = nullptr;
    int cur = -1;
    bool done[LANES];
    int alive = 0;
    int arrived = 0;
    unsigned gen = 0;
    uint64_t vals[2][LANES];
    unsigned tid_base = 0;
    void (*body)(void*) = nullptr;
    void* arg = nullptr;
    long idle_switches = 0;
};
static Warp W;

static inline int lane_id() { return W.cur; }

static void switch_to_next()
{
    // round robin over the lanes that have not returned yet
    const int from = W.cur;
    for (int k = 1; k <= LANES; ++k) {
        const int nxt = (from + k) % LANES;
        if (!W.done[nxt]) {
            if (nxt == from) {   // only this lane is left and it is waiting: nobody can release it
                fprintf(stderr, "simt: deadlock -- lane %d waits at a warp collective the other lanes never reach\n", from);
                abort();
            }
            if (++W.idle_switches > 50L * 1000 * 1000) {
                fprintf(stderr, "simt: deadlock -- %d of %d live lanes arrived at a warp collective, the rest never do\n", W.arrived, W.alive);
                abort();
            }
            W.cur = nxt;
            threadIdx.x = W.tid_base + nxt;
            swapcontext(&W.ctx[from], &W.ctx[nxt]);
            return;
        }
    }
}

// rendezvous of all live lanes; returns the generation whose value buffer holds this collective's inputs
static inline unsigned barrier()
{
    const unsigned g0 = W.gen;
    if (++W.arrived == W.alive) { W.arrived = 0; W.gen = g0 + 1; W.idle_switches = 0; }
    else { while (W.gen == g0) switch_to_next(); }
    return g0;
}

static inline void check_mask(unsigned mask)
{
    if (mask != 0xffffffffu) { fprintf(stderr, "simt: only full-warp collectives are emulated (mask %08x)\n", mask); abort(); }
    if (W.alive != LANES) { fprintf(stderr, "simt: collective after %d lanes have exited\n", LANES - W.alive); abort(); }
}

static inline uint64_t exchange_read(unsigned g0, int src) { return W.vals[g0 & 1][src & 31]; }
static inline unsigned exchange(uint64_t v)
{
    W.vals[W.gen & 1][W.cur] = v;
    return barrier();
}

static void trampoline()
{
    W.body(W.arg);
    W.done[W.cur] = true;
    W.alive -= 1;
    if (W.alive > 0 && W.arrived == W.alive && W.arrived > 0) {
        fprintf(stderr, "simt: a lane exited while the others wait at a warp collective\n");
        abort();
    }
    // hand over to another live lane, or back to the caller when the warp is finished
    for (int k = 1; k <= LANES; ++k) {
        const int nxt = (W.cur + k) % LANES;
        if (!W.done[nxt]) { W.cur = nxt; threadIdx.x = W.tid_base + nxt; setcontext(&W.ctx[nxt]); }
    }
    setcontext(&W.main_ctx);
}

// run body(arg) once per lane of one warp whose first thread has threadIdx.x == tid_base
static void run_warp(void (*body)(void*), void* arg, unsigned tid_base)
{
    if (!W.stacks) W.stacks = (char*)malloc(STACK * LANES);
    W.body = body; W.arg = arg; W.tid_base = tid_base;
    W.alive = LANES; W.arrived = 0; W.gen = 0; W.idle_switches = 0;
    for (int l = 0; l < LANES; ++l) {
        W.done[l] = false;
        getcontext(&W.ctx[l]);
        W.ctx[l].uc_stack.ss_sp = W.stacks + STACK * l;
        W.ctx[l].uc_stack.ss_size = STACK;
        W.ctx[l].uc_link = nullptr;
        makecontext(&W.ctx[l], trampoline, 0);
    }
    W.cur = 0; threadIdx.x = tid_base;
    swapcontext(&W.main_ctx, &W.ctx[0]);
}

}  // namespace simt

// ---- the warp-level CUDA builtins k_step uses -----------------------------------------------------------------------
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::check_mask(mask); simt::barrier(); }
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    simt::check_mask(mask);
    const unsigned g = simt::exchange(pred ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) r |= (unsigned)(simt::exchange_read(g, l) & 1) << l;
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xffffffffu; }
template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    (void)width;
    simt::check_mask(mask);
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    const unsigned g = simt::exchange(bits);
    const uint64_t rb = simt::exchange_read(g, src);
    T out; memcpy(&out, &rb, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    (void)width;
    simt::check_mask(mask);
    uint64_t bits = 0; memcpy(&bits, &v, sizeof(T));
    const unsigned g = simt::exchange(bits);
    const uint64_t rb = simt::exchange_read(g, simt::lane_id() ^ lanemask);
    T out; memcpy(&out, &rb, sizeof(T));
    return out;
}
static inline int __reduce_add_sync(unsigned mask, int v)
{
    simt::check_mask(mask);
    const unsigned g = simt::exchange((uint64_t)(uint32_t)v);
    int s = 0;
    for (int l = 0; l < 32; ++l) s += (int)(uint32_t)simt::exchange_read(g, l);
    return s;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
template <typename T> static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <typename T> static inline T __ldcg(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, T v) { *p = v; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __syncthreads() { fprintf(stderr, "simt: __syncthreads is not emulated (one warp at a time)\n"); abort(); }
using std::min;
using std::max;
