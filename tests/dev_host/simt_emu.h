// simt_emu.h -- TEST INFRASTRUCTURE: a small SIMT emulator that lets g++-compiled CUDA kernels run on the CPU, thread for thread.
//
// Every CUDA thread of a block is a fiber (ucontext) of ONE OS thread; fibers run until they reach a warp- or block-level collective
// (__ballot_sync, __shfl_*_sync, __any_sync, __all_sync, __syncthreads), deposit their operand and yield; when every live lane of the
// warp (thread of the block) has arrived, the operands are snapshotted and the lanes resume.  That is exactly the lock-step contract
// the `_sync` intrinsics give on the GPU for full masks, so warp-cooperative code (vote-scheduled traversal, warp-aggregated queue
// appends, striped work claims) executes with its real control flow.  Atomics are plain read-modify-writes (one OS thread, blocks
// run one after the other), `__shared__` is function-static storage.  What the emulator cannot show: races, memory-model bugs and
// anything about performance.  Include AFTER cuda_host_shim.h and BEFORE the kernel headers.
#pragma once
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

#ifdef __CUDACC__
#error "simt_emu.h is for host compilers only"
#endif

// host_defines.h turned these into attributes the host compiler ignores; the emulator needs real meanings
#undef __global__
#undef __shared__
#undef __launch_bounds__
#define __global__
#define __shared__ static
#define __launch_bounds__(...)

namespace simt {

struct Warp {
    unsigned long long vals[32];
    unsigned long long snap[32];
    unsigned snap_active = 0;      // lanes alive when the snapshot was taken
    int arrived = 0, active = 0;
    unsigned alive_mask = 0;
};
struct Thread {
    ucontext_t ctx;
    uint3 tid;
    int lane = 0, warp = 0;
    bool done = false, waiting_warp = false, waiting_block = false;
};
struct State {
    std::vector<Thread> threads;
    std::vector<Warp> warps;
    std::vector<char> stacks;
    ucontext_t sched;
    Thread* cur = nullptr;
    uint3 block_idx{0, 0, 0}, block_dim{1, 1, 1}, grid_dim{1, 1, 1};
    int block_arrived = 0, block_active = 0;
    const std::function<void()>* body = nullptr;
    size_t stack_bytes = 512 * 1024;
};
inline State& S() { static State s; return s; }

inline void yield_to_scheduler() { State& s = S(); swapcontext(&s.cur->ctx, &s.sched); }

inline void complete_warp(Warp& w, State& s, int warp_id) {
    for (int l = 0; l < 32; l++) w.snap[l] = (w.alive_mask >> l) & 1u ? w.vals[l] : 0ull;
    w.snap_active = w.alive_mask;
    w.arrived = 0;
    for (Thread& t : s.threads) if (t.warp == warp_id) t.waiting_warp = false;
}
inline void complete_block(State& s) {
    s.block_arrived = 0;
    for (Thread& t : s.threads) t.waiting_block = false;
}
// deposit `v`, wait for the warp, get everybody's operands (valid until this lane's next collective)
inline const Warp& warp_exchange(unsigned long long v) {
    State& s = S();
    Thread* t = s.cur;
    Warp& w = s.warps[t->warp];
    w.vals[t->lane] = v;
    w.arrived++;
    if (w.arrived == w.active) complete_warp(w, s, t->warp);
    else { t->waiting_warp = true; yield_to_scheduler(); }
    return w;
}
inline void block_barrier() {
    State& s = S();
    Thread* t = s.cur;
    s.block_arrived++;
    if (s.block_arrived == s.block_active) complete_block(s);
    else { t->waiting_block = true; yield_to_scheduler(); }
}
inline void thread_exit() {
    State& s = S();
    Thread* t = s.cur;
    t->done = true;
    Warp& w = s.warps[t->warp];
    w.active--; w.alive_mask &= ~(1u << t->lane);
    if (w.active > 0 && w.arrived == w.active) complete_warp(w, s, t->warp);     // the others were only waiting for this lane
    s.block_active--;
    if (s.block_active > 0 && s.block_arrived == s.block_active) complete_block(s);
}
inline void fiber_main() {
    (*S().body)();
    thread_exit();
    yield_to_scheduler();
    std::abort();                  // a finished fiber is never resumed
}

// kernel<<<grid, block>>>(args...)  ==  simt::launch(grid, block, [&] { kernel(args...); });
inline void launch(int grid, int block, const std::function<void()>& body) {
    State& s = S();
    s.body = &body;
    s.grid_dim = uint3{(unsigned)grid, 1, 1};
    s.block_dim = uint3{(unsigned)block, 1, 1};
    const int n_warps = (block + 31) / 32;
    if (s.threads.size() != (size_t)block) { s.threads.assign((size_t)block, Thread()); s.stacks.assign((size_t)block * s.stack_bytes, 0); }
    for (int b = 0; b < grid; b++) {
        s.block_idx = uint3{(unsigned)b, 1, 1};
        s.warps.assign((size_t)n_warps, Warp());
        s.block_arrived = 0; s.block_active = block;
        for (int i = 0; i < block; i++) {
            Thread& t = s.threads[i];
            t.tid = uint3{(unsigned)i, 0, 0}; t.lane = i & 31; t.warp = i >> 5;
            t.done = t.waiting_warp = t.waiting_block = false;
            s.warps[t.warp].active++; s.warps[t.warp].alive_mask |= 1u << t.lane;
            getcontext(&t.ctx);
            t.ctx.uc_stack.ss_sp = s.stacks.data() + (size_t)i * s.stack_bytes;
            t.ctx.uc_stack.ss_size = s.stack_bytes;
            t.ctx.uc_link = nullptr;
            makecontext(&t.ctx, fiber_main, 0);
        }
        int remaining = block;
        while (remaining > 0) {
            bool progressed = false;
            for (int i = 0; i < block; i++) {
                Thread& t = s.threads[i];
                if (t.done || t.waiting_warp || t.waiting_block) continue;
                s.cur = &t;
                swapcontext(&s.sched, &t.ctx);
                progressed = true;
                if (t.done) remaining--;
            }
            if (!progressed) {
                std::fprintf(stderr, "simt_emu: deadlock in block %d (%d threads left): a collective was not reached by every live lane\n", b, remaining);
                std::abort();
            }
        }
    }
    s.cur = nullptr;
}

template <typename T> inline unsigned long long to_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

#define threadIdx (simt::S().cur->tid)
#define blockIdx (simt::S().block_idx)
#define blockDim (simt::S().block_dim)
#define gridDim (simt::S().grid_dim)

// ---- warp collectives (full masks only, like every call site in the kernels)
inline unsigned __ballot_sync(unsigned, int pred) {
    const simt::Warp& w = simt::warp_exchange(pred ? 1ull : 0ull);
    unsigned r = 0;
    for (int l = 0; l < 32; l++) r |= (unsigned)(w.snap[l] & 1ull) << l;
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0u; }
inline int __all_sync(unsigned, int pred) {
    const simt::Warp& w = simt::warp_exchange(pred ? 1ull : 0ull);
    for (int l = 0; l < 32; l++) if (((w.snap_active >> l) & 1u) && !(w.snap[l] & 1ull)) return 0;
    return 1;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) {
    const simt::Warp& w = simt::warp_exchange(simt::to_bits(v));
    return simt::from_bits<T>(w.snap[src & 31]);
}
template <typename T> inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
    const int lane = simt::S().cur->lane;
    const simt::Warp& w = simt::warp_exchange(simt::to_bits(v));
    return lane >= (int)delta ? simt::from_bits<T>(w.snap[lane - (int)delta]) : v;
}
template <typename T> inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
    const int lane = simt::S().cur->lane;
    const simt::Warp& w = simt::warp_exchange(simt::to_bits(v));
    return lane + (int)delta < 32 ? simt::from_bits<T>(w.snap[lane + (int)delta]) : v;
}
inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_exchange(0ull); }
inline void __syncthreads() { simt::block_barrier(); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
// ---- atomics: one OS thread, so a plain read-modify-write is atomic
template <typename T> inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <typename T> inline T atomicMin(T* p, T v) { T old = *p; if (v < old) *p = v; return old; }
template <typename T> inline T atomicMax(T* p, T v) { T old = *p; if (v > old) *p = v; return old; }
template <typename T> inline T min(T a, T b) { return b < a ? b : a; }
template <typename T> inline T max(T a, T b) { return a < b ? b : a; }
