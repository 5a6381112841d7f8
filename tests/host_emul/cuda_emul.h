// TEST INFRASTRUCTURE ONLY — a lock-step warp emulation of the few CUDA device features the RoI-pooling kernels use, so
// that the CPU test-suite can execute the KERNEL SOURCE (btcdet_b200/csrc/roi_pool_kernels.cuh, unchanged) on the host
// and compare it with the oracle where no GPU exists.  Nothing here is shipped or reachable from the product
// (btcdet_b200/, spconv/): the product has no CPU path.
//
// Model: one block = one warp = 32 OS threads that meet at a barrier inside every warp collective (__ballot_sync,
// __shfl_sync), which is exact for kernels whose collectives are reached convergently (all of these are).  Blocks run
// one after the other.  Rounded intrinsics map to the host's IEEE operations (compile with -ffp-contract=off).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace emul {
struct Idx {
    unsigned x = 0, y = 0, z = 0;
};
struct WarpState {
    std::barrier<> bar{32};
    unsigned bits[32];
};
inline thread_local Idx t_idx, b_idx, b_dim, g_dim;
inline thread_local WarpState* warp = nullptr;

// Runs `body` as `blocks` blocks of one warp each.
inline void launch(unsigned blocks, const std::function<void()>& body) {
    WarpState ws;
    std::vector<std::thread> th;
    for (unsigned lane = 0; lane < 32; ++lane)
        th.emplace_back([&, lane] {
            warp = &ws;
            t_idx.x = lane;
            b_dim.x = 32;
            g_dim.x = blocks;
            for (unsigned b = 0; b < blocks; ++b) {
                b_idx.x = b;
                body();
                ws.bar.arrive_and_wait();
            }
        });
    for (auto& t : th) t.join();
}
}  // namespace emul

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#define threadIdx emul::t_idx
#define blockIdx emul::b_idx
#define blockDim emul::b_dim
#define gridDim emul::g_dim

template <class T>
inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline void __syncthreads() {}

inline unsigned __ballot_sync(unsigned, bool p) {
    auto* w = emul::warp;
    w->bits[emul::t_idx.x] = p ? 1u : 0u;
    w->bar.arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= w->bits[i] << i;
    w->bar.arrive_and_wait();
    return m;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "32-bit shuffles only");
    auto* w = emul::warp;
    std::memcpy(&w->bits[emul::t_idx.x], &v, 4);
    w->bar.arrive_and_wait();
    T r;
    std::memcpy(&r, &w->bits[src & 31], 4);
    w->bar.arrive_and_wait();
    return r;
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, int d) {
    int lane = (int)emul::t_idx.x;
    T r = __shfl_sync(0xffffffffu, v, lane >= d ? lane - d : lane);
    return r;
}
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
using std::min;
using std::max;
