// TEST INFRASTRUCTURE ONLY — a lock-step emulation of the CUDA device features the index / pooling / RoI kernels use, so
// that the CPU test-suite can execute the KERNEL SOURCE (btcdet_b200/csrc/*.cu, *.cuh) on the host and compare it with the
// oracle and the golden vectors where no GPU exists.  Nothing here is shipped or reachable from the product
// (btcdet_b200/, spconv/): the product has no CPU path, and this emulation is far too slow to be one.
//
// Model: every CUDA thread of a block is its own execution context — a ucontext fiber on the calling thread by default, an
// OS thread with -DEMUL_THREADS (the sanitizer builds); blocks run one after the other, in index order (which also makes
// single-pass look-back scans terminate).  __syncthreads is a barrier over the block, every warp collective
// (__ballot_sync, __shfl_*_sync, __reduce_or_sync) a barrier over the warp's 32 threads — exact for kernels whose
// collectives are reached convergently (a thread that returns early drops out of both barriers, as on the hardware; the
// fiber engine aborts with a message when a barrier can never complete).  `__shared__` variables are function-level
// statics (one block at a time), the dynamic shared memory is a per-launch buffer.  Rounded intrinsics map to the host's
// IEEE operations (compile with -ffp-contract=off); atomics are std::atomic_ref.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>
#include <cstdio>
#include <sys/mman.h>
#include <ucontext.h>

namespace emul {
struct Idx {
    unsigned x = 0, y = 0, z = 0;
};
inline thread_local Idx t_idx, b_idx, b_dim, g_dim;
inline thread_local unsigned lane_id = 0;   // lane = linear thread index % 32
inline unsigned char* dyn_smem = nullptr;   // dynamic shared memory of the running block

#ifdef EMUL_THREADS
// ---- engine 1: one OS thread per CUDA thread (what the sanitizer builds use: ThreadSanitizer then sees every CUDA thread
// as a thread and the barriers / atomics as the only synchronisation) -------------------------------------------------
struct WarpState {
    explicit WarpState(int n) : bar(n) {}
    std::barrier<> bar;
    unsigned bits[32];
};
struct BlockState {
    explicit BlockState(int n) : bar(n) {
        for (int w = 0; w * 32 < n; ++w) warps.emplace_back(new WarpState(std::min(32, n - w * 32)));
    }
    std::barrier<> bar;
    std::vector<std::unique_ptr<WarpState>> warps;
};
inline thread_local WarpState* warp = nullptr;
inline thread_local BlockState* block = nullptr;
inline unsigned* warp_bits() { return warp->bits; }
inline void warp_barrier() { warp->bar.arrive_and_wait(); }
inline void block_barrier() { block->bar.arrive_and_wait(); }

// The same threads run the blocks one after the other (block barriers are per block: a thread that returned early has
// dropped out of them; `next` keeps the blocks from overlapping).
inline void launch(dim3 grid, dim3 blk, size_t smem, const std::function<void()>& body) {
    std::vector<unsigned char> dyn(smem + 64);
    dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    const unsigned block_threads = blk.x * blk.y * blk.z;
    const unsigned nblocks = grid.x * grid.y * grid.z;
    if (block_threads == 0 || nblocks == 0) return;
    std::vector<std::unique_ptr<BlockState>> states;
    states.reserve(nblocks);
    for (unsigned b = 0; b < nblocks; ++b) states.emplace_back(new BlockState((int)block_threads));
    std::barrier<> next((int)block_threads);
    std::vector<std::thread> th;
    th.reserve(block_threads);
    for (unsigned t = 0; t < block_threads; ++t)
        th.emplace_back([&, t] {
            lane_id = t & 31;
            t_idx.x = t % blk.x; t_idx.y = (t / blk.x) % blk.y; t_idx.z = t / (blk.x * blk.y);
            b_dim.x = blk.x; b_dim.y = blk.y; b_dim.z = blk.z;
            g_dim.x = grid.x; g_dim.y = grid.y; g_dim.z = grid.z;
            for (unsigned b = 0; b < nblocks; ++b) {
                BlockState& bs = *states[b];
                block = &bs;
                warp = bs.warps[t >> 5].get();     // warps are formed over the linear thread index
                b_idx.x = b % grid.x; b_idx.y = (b / grid.x) % grid.y; b_idx.z = b / (grid.x * grid.y);
                body();
                warp->bar.arrive_and_drop();      // a finished thread no longer takes part in collectives / barriers
                bs.bar.arrive_and_drop();
                next.arrive_and_wait();
            }
        });
    for (auto& t : th) t.join();
    dyn_smem = nullptr;
}
#else
// ---- engine 2 (default): one FIBER per CUDA thread, all on the calling OS thread.  A barrier is a context switch to the
// next fiber that can run, so a block costs microseconds instead of futex storms; execution is deterministic; a barrier
// that can never complete (divergent collectives) aborts with a message instead of hanging. -----------------------------
struct Barrier {
    int gen = 0, arrived = 0, alive = 0;
};
struct Fiber {
    ucontext_t ctx;
    unsigned tid = 0;
    bool done = false;
    const int* wait_gen = nullptr;
    int wait_val = 0;
};
struct FiberBlock {
    ucontext_t sched;
    std::vector<Fiber> fibers;
    Fiber* cur = nullptr;
    Barrier block_bar;
    Barrier warp_bar[32];
    unsigned bits[32][32];
    const std::function<void()>* body = nullptr;
    dim3 blk;
};
inline FiberBlock* fb = nullptr;

inline void fiber_wait(Barrier& b) {
    const int g = b.gen;
    if (++b.arrived >= b.alive) {       // last one in: release the others, go on
        b.arrived = 0;
        ++b.gen;
        return;
    }
    Fiber* me = fb->cur;
    me->wait_gen = &b.gen;
    me->wait_val = g;
    swapcontext(&me->ctx, &fb->sched);
    me->wait_gen = nullptr;
}
inline void fiber_leave(Barrier& b) {   // a finished thread no longer takes part
    --b.alive;
    if (b.alive > 0 && b.arrived >= b.alive) {
        b.arrived = 0;
        ++b.gen;
    }
}
inline void fiber_entry() {
    (*fb->body)();
    Fiber* me = fb->cur;
    me->done = true;
    fiber_leave(fb->warp_bar[me->tid >> 5]);
    fiber_leave(fb->block_bar);
    swapcontext(&me->ctx, &fb->sched);
}
inline unsigned* warp_bits() { return fb->bits[fb->cur->tid >> 5]; }
inline void warp_barrier() { fiber_wait(fb->warp_bar[fb->cur->tid >> 5]); }
inline void block_barrier() { fiber_wait(fb->block_bar); }

inline void launch(dim3 grid, dim3 blk, size_t smem, const std::function<void()>& body) {
    std::vector<unsigned char> dyn(smem + 64);
    dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
    const unsigned n = blk.x * blk.y * blk.z;
    const unsigned nblocks = grid.x * grid.y * grid.z;
    if (n == 0 || nblocks == 0) return;
    constexpr size_t kStack = 256 * 1024;
    char* stacks = (char*)mmap(nullptr, kStack * n, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (stacks == MAP_FAILED) { std::fprintf(stderr, "cuda_emul: cannot map fiber stacks\n"); std::abort(); }
    FiberBlock state;
    FiberBlock* outer = fb;
    fb = &state;
    state.body = &body;
    state.blk = blk;
    state.fibers.resize(n);
    b_dim.x = blk.x; b_dim.y = blk.y; b_dim.z = blk.z;
    g_dim.x = grid.x; g_dim.y = grid.y; g_dim.z = grid.z;
    for (unsigned b = 0; b < nblocks; ++b) {
        b_idx.x = b % grid.x; b_idx.y = (b / grid.x) % grid.y; b_idx.z = b / (grid.x * grid.y);
        state.block_bar = Barrier{0, 0, (int)n};
        for (unsigned w = 0; w * 32 < n; ++w) state.warp_bar[w] = Barrier{0, 0, (int)std::min(32u, n - w * 32)};
        for (unsigned t = 0; t < n; ++t) {
            Fiber& f = state.fibers[t];
            f.tid = t; f.done = false; f.wait_gen = nullptr;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = stacks + kStack * t;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &state.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        unsigned remaining = n;
        while (remaining) {
            bool progressed = false;
            for (unsigned t = 0; t < n; ++t) {
                Fiber& f = state.fibers[t];
                if (f.done || (f.wait_gen && *f.wait_gen == f.wait_val)) continue;
                progressed = true;
                state.cur = &f;
                lane_id = t & 31;
                t_idx.x = t % blk.x; t_idx.y = (t / blk.x) % blk.y; t_idx.z = t / (blk.x * blk.y);
                swapcontext(&state.sched, &f.ctx);
                if (f.done) --remaining;
            }
            if (!progressed) {
                std::fprintf(stderr, "cuda_emul: deadlock — a barrier / warp collective was not reached by every live thread of "
                                     "block %u (divergent collective?)\n", b);
                std::abort();
            }
        }
    }
    munmap(stacks, kStack * n);
    fb = outer;
    dyn_smem = nullptr;
}
#endif
// the form the RoI tests use: `blocks` blocks of one warp
inline void launch(unsigned blocks, const std::function<void()>& body) { launch(dim3(blocks), dim3(32), 0, body); }
}  // namespace emul

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#undef __shared__
#define __shared__ static
#define threadIdx emul::t_idx
#define blockIdx emul::b_idx
#define blockDim emul::b_dim
#define gridDim emul::g_dim
#define warpSize 32

template <class T>
inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __ffsll(unsigned long long v) { return __builtin_ffsll((long long)v); }
inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
inline void __syncthreads() { emul::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emul::warp_barrier(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

// ---- warp collectives -------------------------------------------------------------------------------------------------
namespace emul {
template <class T>
inline T exchange(T v, int src) {   // every lane publishes v, reads lane `src`'s
    static_assert(sizeof(T) == 4, "32-bit collectives only");
    unsigned* bits = warp_bits();
    std::memcpy(&bits[lane_id], &v, 4);
    warp_barrier();
    T r;
    std::memcpy(&r, &bits[src & 31], 4);
    warp_barrier();
    return r;
}
}  // namespace emul
inline unsigned __ballot_sync(unsigned, bool p) {
    unsigned* bits = emul::warp_bits();
    bits[emul::lane_id] = p ? 1u : 0u;
    emul::warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (bits[i] & 1u) << i;
    emul::warp_barrier();
    return m;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    unsigned* bits = emul::warp_bits();
    bits[emul::lane_id] = v;
    emul::warp_barrier();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= bits[i];
    emul::warp_barrier();
    return m;
}
inline int __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emul::exchange(v, src); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) {
    const int lane = (int)(emul::lane_id);
    return emul::exchange(v, lane >= d ? lane - d : lane);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, int d) {
    const int lane = (int)(emul::lane_id);
    return emul::exchange(v, lane + d < 32 ? lane + d : lane);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
    const int lane = (int)(emul::lane_id);
    return emul::exchange(v, lane ^ m);
}

// ---- atomics ----------------------------------------------------------------------------------------------------------
template <class T>
inline T atomicAdd(T* p, T v) { return std::atomic_ref<T>(*p).fetch_add(v); }
template <class T>
inline T atomicOr(T* p, T v) { return std::atomic_ref<T>(*p).fetch_or(v); }
template <class T>
inline T atomicAnd(T* p, T v) { return std::atomic_ref<T>(*p).fetch_and(v); }
template <class T>
inline T atomicExch(T* p, T v) { return std::atomic_ref<T>(*p).exchange(v); }
template <class T>
inline T atomicCAS(T* p, T expected, T desired) {
    std::atomic_ref<T>(*p).compare_exchange_strong(expected, desired);
    return expected;
}
template <class T>
inline T atomicMin(T* p, T v) {
    std::atomic_ref<T> a(*p);
    T old = a.load();
    while (v < old && !a.compare_exchange_weak(old, v)) {}
    return old;
}
template <class T>
inline T atomicMax(T* p, T v) {
    std::atomic_ref<T> a(*p);
    T old = a.load();
    while (v > old && !a.compare_exchange_weak(old, v)) {}
    return old;
}
using std::max;
using std::min;

// nvcc's templated overload (cuda_runtime.h, __CUDACC__ only): opting a kernel into large dynamic shared memory is a no-op here
template <class R, class... A>
inline cudaError_t cudaFuncSetAttribute(R (*)(A...), cudaFuncAttribute, int) { return cudaSuccess; }

// block-wide OR of a predicate with the barrier semantics of __syncthreads
inline int __syncthreads_or(int p) {
    static std::atomic<int> acc{0};
    emul::block_barrier();
    if (p) acc.store(1);
    emul::block_barrier();
    const int r = acc.load();
    emul::block_barrier();
    acc.store(0);
    emul::block_barrier();
    return r;
}

// conversions with explicit rounding (round to nearest even, like the host's default rounding mode)
inline long long __double2ll_rn(double v) { return std::llrint(v); }
inline long long __float2ll_rn(float v) { return std::llrintf(v); }
inline int __float2int_rn(float v) { return (int)std::lrintf(v); }
inline int __float2int_rd(float v) { return (int)std::floor(v); }
inline int __float2int_rz(float v) { return (int)v; }
inline double __ll2double_rn(long long v) { return (double)v; }
inline float __ll2float_rn(long long v) { return (float)v; }
inline float __int2float_rn(int v) { return (float)v; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
