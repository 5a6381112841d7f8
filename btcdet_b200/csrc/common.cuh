// Shared device/host helpers for the sm_100a kernels of the BtcDet hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/btcdet_b200.h"

namespace btc {

// ---- error plumbing ---------------------------------------------------------
extern thread_local char g_last_error[256];
int set_error(int code, const char* what, cudaError_t e);

#define BTC_CHECK_LAUNCH(what)                                   \
    do {                                                         \
        cudaError_t e__ = cudaGetLastError();                    \
        if (e__ != cudaSuccess) return btc::set_error(BTC_E_CUDA, what, e__); \
    } while (0)

#define BTC_CUDA(call, what)                                     \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) return btc::set_error(BTC_E_CUDA, what, e__); \
    } while (0)

inline int badarg(const char* what) { return set_error(BTC_E_BADARG, what, cudaSuccess); }

// B200: 148 SMs.  Grids of streaming kernels are capped at two waves of 8 CTAs per SM and grid-stride, so a launch
// sized by a *capacity* (levels are planned at 2x growth; the live count is a fraction of it) does not pay for thousands
// of CTAs that start only to find nothing to do (measured: ~1 us per wave of dead CTAs, profiles/r2_index_kernels.md).
constexpr int kNumSM = 148;
inline int grid_for(int64_t n, int threads, int max_waves = 2, int ctas_per_sm = 8) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = (int64_t)kNumSM * ctas_per_sm * max_waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

__device__ __forceinline__ int live_count(int n_cap, const int* n_dev) {
    if (n_dev == nullptr) return n_cap;
    int n = __ldg(n_dev);
    return n < n_cap ? n : n_cap;
}

// ---- geometry passed by value to kernels --------------------------------------
struct Shape3 {
    int d, h, w;  // z, y, x extents
};

struct ConvGeom {
    Shape3 in, out;
    int k[3], s[3], p[3], dil[3];
    int batch;
    int transposed;
    int K;
};

__device__ __forceinline__ int64_t flat_key(int b, int z, int y, int x, const Shape3& s) {
    return (((int64_t)b * s.d + z) * s.h + y) * (int64_t)s.w + x;
}

// ---- rank bitmap ------------------------------------------------------------
// entry = {bits (low 32), rank-before-word (high 32)} for cells [32w, 32w+32).
__device__ __forceinline__ int index_lookup(const uint2* __restrict__ index, int64_t key) {
    uint2 e = __ldg(index + (key >> 5));
    unsigned bit = 1u << (unsigned)(key & 31);
    if (!(e.x & bit)) return -1;
    return (int)(e.y + __popc(e.x & (bit - 1u)));
}

// ---- coordinate hash (unsorted site sets: submanifold lookups) ---------------------------------------
// Open addressing, linear probing; keys are flat cell keys (int64, -1 = empty), vals the site row.
__device__ __forceinline__ uint32_t hash_key64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return (uint32_t)k;
}
__device__ __forceinline__ int hash_lookup(const long long* __restrict__ keys, const int* __restrict__ vals,
                                           uint32_t hmask, int64_t key) {
    uint32_t h = hash_key64((unsigned long long)key) & hmask;
    while (true) {
        long long kk = __ldg(keys + h);
        if (kk == key) return __ldg(vals + h);
        if (kk == -1LL) return -1;
        h = (h + 1) & hmask;
    }
}

// ---- warp / block scan helpers ------------------------------------------------
__device__ __forceinline__ int warp_inclusive_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Block-wide exclusive scan of one int per thread (blockDim.x <= 1024, multiple of 32).
// Returns the exclusive prefix; *block_total receives the sum (all threads).
__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp /*[32]*/, int* block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarps = (blockDim.x + 31) >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? smem_warp[lane] : 0;
        int winc = warp_inclusive_scan(w);
        smem_warp[lane] = winc - w;  // exclusive warp base
        if (lane == 31) smem_warp[32] = winc;  // total (smem_warp has 33 slots)
    }
    __syncthreads();
    int base = smem_warp[warp];
    *block_total = smem_warp[32];
    __syncthreads();
    return base + inc - v;
}

// Device-wide exclusive scan of popcounts / flags: three small kernels.
// tile = kScanThreads * kScanItems elements per block, striped per iteration.
constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

inline int scan_num_blocks(int64_t n) { return (int)((n + kScanTile - 1) / kScanTile); }

// workspace layout helpers
inline int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Launchers implemented in coord_index.cu
// Scans the occupancy bits of `index` (n_entries), writing ranks in place and the total to *total (device, may be null).
int launch_index_scan(uint2* index, int64_t n_entries, int* block_sums /*[scan_num_blocks+1]*/, int* total,
                      cudaStream_t stream);
// Exclusive scan of int flags (n_cap elements, live count n_dev) -> out; total -> *total.
int launch_flag_scan(const int* flags, int* out, int n_cap, int* block_sums, int* total, cudaStream_t stream);

}  // namespace btc
