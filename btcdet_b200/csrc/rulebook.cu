// Rulebook ("indice pairs") construction — replaces spconv 1.2.1 src/spconv/indice.cu
// (getSubMIndicePairs / getIndicePairsConv / ...DeConv + torch::_unique) reached from
// SparseConvolution.forward; reference call sites
// btcdet/models/backbones_3d/spconv_backbone.py:12-29 (post_act_block) and every
// SubMConv3d / SparseConv3d / SparseConvTranspose3d / SparseMaxPool3d layer built there.
//
// Design (B200-first): no sort, no atomics on counters, no host sync.
//   * output sites of a strided / transposed conv are marked in a rank bitmap of the output
//     grid (L2-resident); one popcount scan yields both the number of sites and, for each
//     site, its row = rank in ascending flat-key order (the order torch::_unique gave spconv);
//   * neighbour tables are written output-stationary (nbr_out[o][k]) for the conv kernels and
//     input-major (nbr_in[i][k]) for backward / the spconv-format pair list;
//   * the spconv-format pair list is produced in canonical (offset, input row) order with a
//     warp-ballot + prefix-sum stream compaction per offset — deterministic, unlike the
//     atomicAdd slot order of the original.
#include "common.cuh"

namespace btc {

__device__ __forceinline__ void offset_of(int k, const int* ks, int& kz, int& ky, int& kx) {
    kx = k % ks[2];
    int t = k / ks[2];
    ky = t % ks[1];
    kz = t / ks[1];
}

// Output coordinate along one axis for input coordinate `in`, kernel tap `kk`.
// regular:    out*s - p + kk*d == in   ->  out = (in + p - kk*d) / s  when divisible
// transposed: out = in*s - p + kk*d
__device__ __forceinline__ bool out_coord(int in, int kk, int s, int p, int d, int out_dim, int transposed, int& o) {
    if (transposed) {
        o = in * s - p + kk * d;
    } else {
        int t = in + p - kk * d;
        if (t < 0) return false;
        if (s != 1) {
            if (t % s) return false;
            t /= s;
        }
        o = t;
    }
    return o >= 0 && o < out_dim;
}

// ---- submanifold ----------------------------------------------------------------
// Block = 32 tap lanes x 8 sites (the 27 probes of a site hit ~9 index words that sit in L1/L2); the tap ->
// (kz,ky,kx) decomposition comes from a per-block table, so there is no integer division on the hot path.
// HASH = probe a coordinate hash (unsorted rows on huge grids: the KITTI det level-1 grid is 92 M cells per scene
// for ~14 k sites, a 2N-slot hash is ~1 MB) instead of the rank bitmap.
constexpr int kSubmSites = 8;

__device__ __forceinline__ void fill_tap_table(const ConvGeom& g, int* s_tap /*[K]*/) {
    for (int k = threadIdx.y * 32 + threadIdx.x; k < g.K; k += 32 * kSubmSites) {
        int kz, ky, kx;
        offset_of(k, g.k, kz, ky, kx);
        s_tap[k] = ((kz - g.k[0] / 2) * g.dil[0] + 64) | (((ky - g.k[1] / 2) * g.dil[1] + 64) << 8) |
                   (((kx - g.k[2] / 2) * g.dil[2] + 64) << 16);
    }
    __syncthreads();
}

template <bool HASH>
__global__ void __launch_bounds__(32 * kSubmSites)
subm_table_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                  const uint2* __restrict__ index, const int* __restrict__ perm, const long long* __restrict__ keys,
                  const int* __restrict__ vals, uint32_t hmask, int* __restrict__ nbr_out) {
    __shared__ int s_tap[256];
    fill_tap_table(g, s_tap);
    const int n = live_count(n_cap, n_dev);
    const int K = g.K;
    for (int i = blockIdx.x * kSubmSites + threadIdx.y; i < n; i += gridDim.x * kSubmSites) {
        const int4 c = __ldg(coords + i);
        for (int k = threadIdx.x; k < K; k += 32) {
            const int tap = s_tap[k];
            const int z = c.y + (tap & 255) - 64, y = c.z + ((tap >> 8) & 255) - 64, x = c.w + ((tap >> 16) & 255) - 64;
            int j = -1;
            if ((unsigned)z < (unsigned)g.in.d && (unsigned)y < (unsigned)g.in.h && (unsigned)x < (unsigned)g.in.w) {
                const int64_t key = flat_key(c.x, z, y, x, g.in);
                if (HASH) {
                    j = (k == K / 2) ? i : hash_lookup(keys, vals, hmask, key);
                } else {
                    int r = index_lookup(index, key);
                    if (r >= 0 && r < n_cap) j = perm ? __ldg(perm + r) : r;   // rank >= capacity: site not materialised
                }
            }
            nbr_out[(int64_t)i * K + k] = j;
        }
    }
}

// Symmetric variant (K/2 <= 16, i.e. every 3x3x3 layer): the neighbour relation of a sub-manifold rulebook is its own
// mirror image — j = nbr[i][k]  <=>  i = nbr[j][K-1-k] — so only the first K/2 taps are probed (13 instead of 26
// lookups per site) and each hit writes both entries; misses write nothing (the table is pre-filled with -1 by
// fill_table_kernel, a coalesced 16-byte-store pass).  A warp row handles two sites x 16 tap lanes.
template <bool HASH>
__global__ void __launch_bounds__(32 * kSubmSites)
subm_table_sym_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                      const uint2* __restrict__ index, const int* __restrict__ perm, const long long* __restrict__ keys,
                      const int* __restrict__ vals, uint32_t hmask, int* __restrict__ nbr_out) {
    __shared__ int s_tap[256];
    fill_tap_table(g, s_tap);
    const int n = live_count(n_cap, n_dev);
    const int K = g.K, half = K >> 1;
    const int sub = threadIdx.x >> 4, k = threadIdx.x & 15;
    for (int i = (blockIdx.x * kSubmSites + threadIdx.y) * 2 + sub; i < n; i += gridDim.x * kSubmSites * 2) {
        const int4 c = __ldg(coords + i);
        if (k == 0) nbr_out[(int64_t)i * K + half] = i;          // centre tap: the site itself
        if (k < half) {
            const int tap = s_tap[k];
            const int z = c.y + (tap & 255) - 64, y = c.z + ((tap >> 8) & 255) - 64, x = c.w + ((tap >> 16) & 255) - 64;
            if ((unsigned)z < (unsigned)g.in.d && (unsigned)y < (unsigned)g.in.h && (unsigned)x < (unsigned)g.in.w) {
                const int64_t key = flat_key(c.x, z, y, x, g.in);
                int j = -1;
                if (HASH) {
                    j = hash_lookup(keys, vals, hmask, key);
                } else {
                    int r = index_lookup(index, key);
                    if (r >= 0 && r < n_cap) j = perm ? __ldg(perm + r) : r;   // rank >= capacity: site not materialised
                }
                if (j >= 0 && j < n) {
                    nbr_out[(int64_t)i * K + k] = j;
                    nbr_out[(int64_t)j * K + (K - 1 - k)] = i;
                }
            }
        }
    }
}

// ---- strided / transposed -----------------------------------------------------------
// Valid taps of one axis for input coordinate `in`: regular conv -> those kk with (in + p - kk*d) divisible by s
// (k=3, s=2: one or two of the three), transposed -> every kk landing inside the output.  At most 8 per axis.
constexpr int kMaxAxisTaps = 8;
__device__ __forceinline__ int axis_taps(int in, int k, int s, int p, int d, int out_dim, int transposed,
                                         int (&kk_l)[kMaxAxisTaps], int (&o_l)[kMaxAxisTaps]) {
    int cnt = 0;
    for (int kk = 0; kk < k && cnt < kMaxAxisTaps; ++kk) {
        int o;
        if (out_coord(in, kk, s, p, d, out_dim, transposed, o)) {
            kk_l[cnt] = kk;
            o_l[cnt] = o;
            ++cnt;
        }
    }
    return cnt;
}

// One thread per input site; it enumerates only its valid (kz,ky,kx) taps (3.4 of 27 on average for k3 s2).
__global__ void conv_mark_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                                 unsigned* __restrict__ out_index_words) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        int kz_l[kMaxAxisTaps], oz_l[kMaxAxisTaps], ky_l[kMaxAxisTaps], oy_l[kMaxAxisTaps], kx_l[kMaxAxisTaps], ox_l[kMaxAxisTaps];
        const int nz = axis_taps(c.y, g.k[0], g.s[0], g.p[0], g.dil[0], g.out.d, g.transposed, kz_l, oz_l);
        const int ny = axis_taps(c.z, g.k[1], g.s[1], g.p[1], g.dil[1], g.out.h, g.transposed, ky_l, oy_l);
        const int nx = axis_taps(c.w, g.k[2], g.s[2], g.p[2], g.dil[2], g.out.w, g.transposed, kx_l, ox_l);
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b < ny; ++b)
                for (int d = 0; d < nx; ++d) {
                    const int64_t key = flat_key(c.x, oz_l[a], oy_l[b], ox_l[d], g.out);
                    const unsigned bit = 1u << (unsigned)(key & 31);
                    unsigned* wptr = out_index_words + 2 * (key >> 5);
                    // most taps of neighbouring inputs hit an already-set bit: test before the atomic
                    if (!(*(volatile unsigned*)wptr & bit)) atomicOr(wptr, bit);
                }
    }
}

// Decode every occupied cell of the output index into a coordinate row.  (The table rows are initialised by the
// coalesced fill_table_kernel: doing it here, 27 scattered stores per set bit by the few threads that own occupied
// words, made this kernel 4x slower than the bitmap read it is bound by — round-1 launch list, 67 us -> ~15 us.)
__global__ void conv_emit_kernel(const uint2* __restrict__ out_index, int64_t n_entries, Shape3 out, int out_cap,
                                 int4* __restrict__ out_coords) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_entries; w += (int64_t)gridDim.x * blockDim.x) {
        uint2 e = __ldg(out_index + w);
        unsigned bits = e.x;
        int row = (int)e.y;
        while (bits) {
            int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            if (row < out_cap) {
                int64_t key = (w << 5) + bpos;
                int x = (int)(key % out.w);
                int64_t t = key / out.w;
                int y = (int)(t % out.h);
                t /= out.h;
                int z = (int)(t % out.d);
                int b = (int)(t / out.d);
                out_coords[row] = make_int4(b, z, y, x);
            }
            ++row;
        }
    }
}

// All live rows of a neighbour table to -1: 16-byte stores over the flat [live * K] range (cudaMalloc'ed tables are
// 256-byte aligned; the scalar tail covers live * K % 4).
__global__ void fill_table_kernel(int* __restrict__ table, int n_cap, const int* __restrict__ n_dev, int K) {
    const int64_t work = (int64_t)live_count(n_cap, n_dev) * K;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    if (((uintptr_t)table & 15) == 0) {
        const int64_t vec = work >> 2;
        int4* t4 = reinterpret_cast<int4*>(table);
        for (int64_t t = tid; t < vec; t += nthr) t4[t] = make_int4(-1, -1, -1, -1);
        for (int64_t t = (vec << 2) + tid; t < work; t += nthr) table[t] = -1;
    } else {
        for (int64_t t = tid; t < work; t += nthr) table[t] = -1;
    }
}

__global__ void conv_tables_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                                   const uint2* __restrict__ out_index, int out_cap, int* __restrict__ nbr_out,
                                   int* __restrict__ nbr_in) {
    const int n = live_count(n_cap, n_dev);
    const int K = g.K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        if (nbr_in)
            for (int k = 0; k < K; ++k) nbr_in[(int64_t)i * K + k] = -1;
        int kz_l[kMaxAxisTaps], oz_l[kMaxAxisTaps], ky_l[kMaxAxisTaps], oy_l[kMaxAxisTaps], kx_l[kMaxAxisTaps], ox_l[kMaxAxisTaps];
        const int nz = axis_taps(c.y, g.k[0], g.s[0], g.p[0], g.dil[0], g.out.d, g.transposed, kz_l, oz_l);
        const int ny = axis_taps(c.z, g.k[1], g.s[1], g.p[1], g.dil[1], g.out.h, g.transposed, ky_l, oy_l);
        const int nx = axis_taps(c.w, g.k[2], g.s[2], g.p[2], g.dil[2], g.out.w, g.transposed, kx_l, ox_l);
        for (int a = 0; a < nz; ++a)
            for (int b = 0; b < ny; ++b)
                for (int d = 0; d < nx; ++d) {
                    const int k = (kz_l[a] * g.k[1] + ky_l[b]) * g.k[2] + kx_l[d];
                    int o = index_lookup(out_index, flat_key(c.x, oz_l[a], oy_l[b], ox_l[d], g.out));
                    if (o >= out_cap) o = -1;
                    if (nbr_in) nbr_in[(int64_t)i * K + k] = o;
                    if (nbr_out && o >= 0) nbr_out[(int64_t)o * K + k] = i;
                }
    }
}

// ---- strided / transposed, sparse two-level build (round 2) -------------------------------------------------------
// The dense build above streams the whole rank bitmap of the output grid four times (zero, popcount, rank write, decode:
// 47 MB each at batch 16 on the [21,800,704] level, 92 MB on the stress grid) to compact ~2e5 sites, in 8 dependent
// launches.  Here a SUMMARY bitmap (one bit per 32-cell index word, 1/1024 of the grid in bits) records which words were
// touched, so that ONE kernel — a single-pass scan with decoupled look-back over the summary words — ranks the touched
// words, writes their ranks and decodes the output coordinates, touching only occupied words; both bitmaps are cleared
// sparsely from the coordinate list after their last reader (btc_index_clear_sparse) instead of a memset per step.
// Bytes moved per build: ~16 B per input tap + the summary (0.7 MB at batch 16) + 24 B per output site.
// Output coordinate of one axis with the common strides resolved without an integer division.
__device__ __forceinline__ bool out_coord_fast(int in, int kk, int s, int p, int d, int out_dim, int transposed, int& o) {
    if (transposed) {
        o = in * s - p + kk * d;
    } else {
        int t = in + p - kk * d;
        if (t < 0) return false;
        if (s == 2) {
            if (t & 1) return false;
            t >>= 1;
        } else if (s != 1) {
            if (t % s) return false;
            t /= s;
        }
        o = t;
    }
    return o >= 0 && o < out_dim;
}

// ---- fast path for the geometries the backbones use: per axis (k = 3, s = 2, d = 1) or (k = 1, s = 1), not transposed ----
// The generic kernels below walk k^3 taps through runtime-parameter tests; threads of a warp take different paths and
// the integer work (not memory) bounds them (25-30 us per build at every level size, profiles/r2_index_kernels.md).
// For a stride-2 3-tap axis the valid taps follow from a parity: t = in + p - kk must be even and >= 0, so
// kk0 = (in + p) & 1 with o0 = (in + p - kk0) / 2, and, when kk0 = 0, also kk = 2 with o0 - 1: at most two candidates per
// axis, eight per site, in straight-line code.
struct AxisFast {
    int k, s, p, out_dim;
};
__device__ __forceinline__ int axis_cand(int in, const AxisFast& a, int (&kk)[2], int (&o)[2]) {
    if (a.k == 1) {                       // k = 1, s = 1
        kk[0] = 0;
        o[0] = in + a.p;
        kk[1] = 0;
        o[1] = -1;
        return (o[0] >= 0 && o[0] < a.out_dim) ? 1 : 0;
    }
    const int t = in + a.p;               // k = 3, s = 2
    const int k0 = t & 1, o0 = (t - k0) >> 1;
    int cnt = 0;
    if (o0 >= 0 && o0 < a.out_dim) { kk[cnt] = k0; o[cnt] = o0; ++cnt; }
    if (k0 == 0 && o0 - 1 >= 0 && o0 - 1 < a.out_dim) { kk[cnt] = 2; o[cnt] = o0 - 1; ++cnt; }
    if (cnt < 2) { kk[1] = 0; o[1] = -1; }
    if (cnt < 1) { kk[0] = 0; o[0] = -1; }
    return cnt;
}
static bool fast_geometry(const ConvGeom& g) {
    if (g.transposed) return false;
    for (int a = 0; a < 3; ++a) {
        const bool k3s2 = g.k[a] == 3 && g.s[a] == 2 && g.dil[a] == 1;
        const bool k1s1 = g.k[a] == 1 && g.s[a] == 1;
        if (!k3s2 && !k1s1) return false;
    }
    return true;
}

__global__ void conv_mark2_fast_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, Shape3 out,
                                       AxisFast az, AxisFast ay, AxisFast ax, unsigned* __restrict__ out_index_words,
                                       unsigned* __restrict__ summary) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        int kz[2], oz[2], ky[2], oy[2], kx[2], ox[2];
        const int nz = axis_cand(c.y, az, kz, oz), ny = axis_cand(c.z, ay, ky, oy), nx = axis_cand(c.w, ax, kx, ox);
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                if (a < nz && b < ny && nx > 0) {
                    // the (at most two) x candidates are neighbouring cells: one RED carries both bits unless they straddle a
                    // word boundary.  The LSU / L2 atomic rate bounds this kernel (lg_throttle in ncu), so fewer, fatter
                    // reductions are the lever.
                    const int64_t key0 = flat_key(c.x, oz[a], oy[b], ox[0], out);
                    const int64_t w0 = key0 >> 5;
                    unsigned bits0 = 1u << (unsigned)(key0 & 31);
                    if (nx > 1) {
                        const int64_t key1 = key0 + (ox[1] - ox[0]);
                        const int64_t w1 = key1 >> 5;
                        if (w1 == w0) {
                            bits0 |= 1u << (unsigned)(key1 & 31);
                        } else {
                            atomicOr(out_index_words + 2 * w1, 1u << (unsigned)(key1 & 31));
                            atomicOr(summary + (w1 >> 5), 1u << (unsigned)(w1 & 31));
                        }
                    }
                    atomicOr(out_index_words + 2 * w0, bits0);
                    atomicOr(summary + (w0 >> 5), 1u << (unsigned)(w0 & 31));
                }
            }
    }
}

__global__ void conv_tables2_fast_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, Shape3 out,
                                         AxisFast az, AxisFast ay, AxisFast ax, int K, const uint2* __restrict__ out_index,
                                         int out_cap, int* __restrict__ nbr_out, int* __restrict__ nbr_in,
                                         int4* __restrict__ out_coords) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        if (nbr_in)
            for (int k = 0; k < K; ++k) nbr_in[(int64_t)i * K + k] = -1;
        int kz[2], oz[2], ky[2], oy[2], kx[2], ox[2];
        const int nz = axis_cand(c.y, az, kz, oz), ny = axis_cand(c.z, ay, ky, oy), nx = axis_cand(c.w, ax, kx, ox);
        uint2 e[8];
        int64_t key[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {                  // the (at most) eight rank-bitmap loads, issued together
            const int a = t >> 2, b = (t >> 1) & 1, d = t & 1;
            const bool v = a < nz && b < ny && d < nx;
            key[t] = v ? flat_key(c.x, oz[a], oy[b], ox[d], out) : -1;
            e[t] = v ? __ldg(out_index + (key[t] >> 5)) : make_uint2(0u, 0u);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (key[t] < 0) continue;
            const int a = t >> 2, b = (t >> 1) & 1, d = t & 1;
            const int k = (kz[a] * ay.k + ky[b]) * ax.k + kx[d];
            const unsigned bit = 1u << (unsigned)(key[t] & 31);
            int o = (e[t].x & bit) ? (int)(e[t].y + __popc(e[t].x & (bit - 1u))) : -1;
            if (o >= out_cap) o = -1;
            if (nbr_in) nbr_in[(int64_t)i * K + k] = o;
            if (o >= 0) {
                if (nbr_out) nbr_out[(int64_t)o * K + k] = i;
                if (out_coords) out_coords[o] = make_int4(c.x, oz[a], oy[b], ox[d]);
            }
        }
    }
}

// One thread per input site; nested tap loops with the validity tests inline (no per-thread tap arrays: dynamically
// indexed local arrays live in local memory).  Two fire-and-forget reductions per (site, tap) pair (RED.OR, results
// unused: the thread never waits for the L2).  Measured alternatives (profiles/r2_index_kernels.md): a test-before-atomic
// variant costs a dependent L2 round trip per tap (same time), warp-level aggregation of lanes that hit the same word
// (match.any + redux) is 4.7x SLOWER — match.any serialises over the distinct values of a warp.
__global__ void conv_mark2_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                                  unsigned* __restrict__ out_index_words, unsigned* __restrict__ summary) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        for (int kz = 0; kz < g.k[0]; ++kz) {
            int oz;
            if (!out_coord_fast(c.y, kz, g.s[0], g.p[0], g.dil[0], g.out.d, g.transposed, oz)) continue;
            for (int ky = 0; ky < g.k[1]; ++ky) {
                int oy;
                if (!out_coord_fast(c.z, ky, g.s[1], g.p[1], g.dil[1], g.out.h, g.transposed, oy)) continue;
                for (int kx = 0; kx < g.k[2]; ++kx) {
                    int ox;
                    if (!out_coord_fast(c.w, kx, g.s[2], g.p[2], g.dil[2], g.out.w, g.transposed, ox)) continue;
                    const int64_t key = flat_key(c.x, oz, oy, ox, g.out);
                    const int64_t w = key >> 5;
                    atomicOr(out_index_words + 2 * w, 1u << (unsigned)(key & 31));
                    atomicOr(summary + (w >> 5), 1u << (unsigned)(w & 31));
                }
            }
        }
    }
}

// Single-pass ranking scan over the summary words (decoupled look-back).  A tile is kSrWarps x kSrWordsPerWarp = 256
// summary words (8 Ki index words = 256 Ki cells), handed out by an atomic counter so that a tile only ever waits on
// tiles that already run; tiles are chained through `status` (high word: 1 = tile aggregate, 2 = inclusive prefix) with a
// warp-wide look-back (32 predecessors per round trip).  A warp owns 16 consecutive summary words; lane l owns index
// word l of each: the (at most 16) loads of a lane are independent and issued together — only flagged words are read —
// and a word's rank (occupied cells before it) is written in place.  The output coordinates are written by
// conv_tables2 (which knows them without decoding a flat key).
constexpr int kSrWarps = 16;
constexpr int kSrWordsPerWarp = 16;
constexpr int kSeThreads = kSrWarps * 32;
constexpr int kSeTile = kSrWarps * kSrWordsPerWarp;      // summary words per tile
__global__ void __launch_bounds__(kSeThreads)
scan_rank_kernel(uint2* __restrict__ index, const unsigned* __restrict__ summary, int n_sum,
                 unsigned long long* __restrict__ status, int* __restrict__ tile_ctr, int* __restrict__ total) {
    __shared__ int s_wtot[kSrWarps];
    __shared__ int s_tile, s_base;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_ctr, 1);
    __syncthreads();
    const int tile = s_tile;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sw0 = tile * kSeTile + warp * kSrWordsPerWarp;            // first summary word of this warp
    const unsigned mine = (lane < kSrWordsPerWarp && sw0 + lane < n_sum) ? __ldg(summary + sw0 + lane) : 0u;
    int pc[kSrWordsPerWarp];
#pragma unroll
    for (int j = 0; j < kSrWordsPerWarp; ++j) {                         // independent loads, all in flight together
        const unsigned sw = __shfl_sync(0xffffffffu, mine, j);
        pc[j] = ((sw >> lane) & 1u) ? __popc(index[(int64_t)(sw0 + j) * 32 + lane].x) : 0;
    }
    int ex[kSrWordsPerWarp];
    int running = 0;
#pragma unroll
    for (int j = 0; j < kSrWordsPerWarp; ++j) {                         // exclusive ranks in (summary word, lane) order
        const int inc = warp_inclusive_scan(pc[j]);
        ex[j] = running + inc - pc[j];
        running += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_wtot[warp] = running;
    __syncthreads();
    if (warp == 0) {
        int v = lane < kSrWarps ? s_wtot[lane] : 0;
        const int inc = warp_inclusive_scan(v);
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (lane < kSrWarps) s_wtot[lane] = inc - v;                    // exclusive warp bases
        volatile unsigned long long* st = status;
        int base = 0;
        if (tile > 0) {
            if (lane == 0) st[tile] = (1ull << 32) | (unsigned)tot;      // this tile's aggregate
            int p_hi = tile - 1;                                          // nearest predecessor not yet accounted for
            while (true) {
                const int p = p_hi - lane;
                unsigned long long pv = (2ull << 32);                     // before tile 0: inclusive prefix 0
                if (p >= 0) {
                    do { pv = st[p]; } while ((pv >> 32) == 0ull);        // predecessors run already: bounded wait
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (pv >> 32) == 2ull);
                const int first = incl ? __ffs(incl) - 1 : 31;            // nearest predecessor with an inclusive prefix
                int c = lane <= first ? (int)(unsigned)pv : 0;
                for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
                base += c;
                if (incl) break;
                p_hi -= 32;
            }
        }
        if (lane == 0) {
            st[tile] = (2ull << 32) | (unsigned)(base + tot);
            s_base = base;
            if (tile == (int)gridDim.x - 1 && total) *total = base + tot;
        }
    }
    __syncthreads();
    const int row0 = s_base + s_wtot[warp];
#pragma unroll
    for (int j = 0; j < kSrWordsPerWarp; ++j) {
        const unsigned sw = __shfl_sync(0xffffffffu, mine, j);
        if ((sw >> lane) & 1u) index[(int64_t)(sw0 + j) * 32 + lane].y = (unsigned)(row0 + ex[j]);
    }
}

// conv_tables + the output coordinate rows: every (input, tap) pair knows its output cell (b, oz, oy, ox) and, through
// the rank bitmap, its row — so the coordinate list needs no decode pass over the bitmap (pairs that share an output
// store the same 16 bytes).
__global__ void conv_tables2_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, ConvGeom g,
                                    const uint2* __restrict__ out_index, int out_cap, int* __restrict__ nbr_out,
                                    int* __restrict__ nbr_in, int4* __restrict__ out_coords) {
    const int n = live_count(n_cap, n_dev);
    const int K = g.K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        if (nbr_in)
            for (int k = 0; k < K; ++k) nbr_in[(int64_t)i * K + k] = -1;
        // pass 1: issue the rank-bitmap loads of up to 8 valid taps (all of a k3 s2 site's) without using them;
        // pass 2: resolve rows and write.  Sites with more valid taps (transposed / stride-1 kernels) take further rounds.
        int done = 0;                                   // valid taps already handled
        while (true) {
            uint2 e[8];
            long long packed[8];                        // k | oz << 10 | oy << 24 | ox << 44
            int64_t key[8];
            int cnt = 0, seen = 0;
            for (int kz = 0; kz < g.k[0] && cnt < 8; ++kz) {
                int oz;
                if (!out_coord_fast(c.y, kz, g.s[0], g.p[0], g.dil[0], g.out.d, g.transposed, oz)) continue;
                for (int ky = 0; ky < g.k[1] && cnt < 8; ++ky) {
                    int oy;
                    if (!out_coord_fast(c.z, ky, g.s[1], g.p[1], g.dil[1], g.out.h, g.transposed, oy)) continue;
                    for (int kx = 0; kx < g.k[2] && cnt < 8; ++kx) {
                        int ox;
                        if (!out_coord_fast(c.w, kx, g.s[2], g.p[2], g.dil[2], g.out.w, g.transposed, ox)) continue;
                        if (seen++ < done) continue;
                        const int64_t kk = flat_key(c.x, oz, oy, ox, g.out);
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j == cnt) {             // static indices keep the batch in registers
                                key[j] = kk;
                                packed[j] = (long long)((kz * g.k[1] + ky) * g.k[2] + kx) | ((long long)oz << 10) |
                                            ((long long)oy << 24) | ((long long)ox << 44);
                                e[j] = __ldg(out_index + (kk >> 5));
                            }
                        ++cnt;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j >= cnt) break;
                const int k = (int)(packed[j] & 1023);
                const unsigned bit = 1u << (unsigned)(key[j] & 31);
                int o = (e[j].x & bit) ? (int)(e[j].y + __popc(e[j].x & (bit - 1u))) : -1;
                if (o >= out_cap) o = -1;
                if (nbr_in) nbr_in[(int64_t)i * K + k] = o;
                if (o >= 0) {
                    if (nbr_out) nbr_out[(int64_t)o * K + k] = i;
                    if (out_coords)
                        out_coords[o] = make_int4(c.x, (int)((packed[j] >> 10) & 16383), (int)((packed[j] >> 24) & 1048575),
                                                  (int)(packed[j] >> 44));
                }
            }
            done += cnt;
            if (cnt < 8) break;
        }
    }
}

__global__ void index_clear2_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, Shape3 shape,
                                    int batch, uint2* __restrict__ index, unsigned* __restrict__ summary) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);
        if ((unsigned)c.x >= (unsigned)batch || (unsigned)c.y >= (unsigned)shape.d || (unsigned)c.z >= (unsigned)shape.h ||
            (unsigned)c.w >= (unsigned)shape.w)
            continue;
        const int64_t w = flat_key(c.x, c.y, c.z, c.w, shape) >> 5;
        index[w] = make_uint2(0u, 0u);
        summary[w >> 5] = 0u;
    }
}

// ---- spconv-format pair list, canonical order ------------------------------------------------
constexpr int kPairThreads = 256;
constexpr int kPairRowsPerBlock = 2048;

__device__ __forceinline__ int table_at(const int* __restrict__ table, int i, int k, int K, int mirror) {
    return __ldg(table + (int64_t)i * K + (mirror ? K - 1 - k : k));
}

// grid (row blocks, K): count valid entries of column k inside the block's row range.
__global__ void __launch_bounds__(kPairThreads) pairs_count_kernel(const int* __restrict__ table, int n_cap,
                                                                  const int* __restrict__ n_dev, int K, int mirror,
                                                                  int* __restrict__ block_counts /*[K][nblk]*/) {
    __shared__ int s_warp[kPairThreads / 32];
    const int n = live_count(n_cap, n_dev);
    const int k = blockIdx.y;
    const int r0 = blockIdx.x * kPairRowsPerBlock;
    int cnt = 0;
    for (int r = r0 + threadIdx.x; r < r0 + kPairRowsPerBlock && r < n; r += kPairThreads)
        cnt += table_at(table, r, k, K, mirror) >= 0;
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, d);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < kPairThreads / 32; ++w) tot += s_warp[w];
        block_counts[k * gridDim.x + blockIdx.x] = tot;
    }
}

// grid K blocks: exclusive scan over the row blocks of one offset; writes pair_num[k].
__global__ void __launch_bounds__(256) pairs_scan_kernel(int* __restrict__ block_counts, int nblk,
                                                        int* __restrict__ pair_num) {
    __shared__ int s_warp[33];
    int* row = block_counts + (int64_t)blockIdx.x * nblk;
    int carry = 0;
    for (int base = 0; base < nblk; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < nblk ? row[i] : 0;
        int tot;
        int ex = block_exclusive_scan(v, s_warp, &tot);
        if (i < nblk) row[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) pair_num[blockIdx.x] = carry;
}

// grid (row blocks, K): ballot + popc compaction, rows ascending.
__global__ void __launch_bounds__(kPairThreads) pairs_write_kernel(const int* __restrict__ table, int n_cap,
                                                                  const int* __restrict__ n_dev, int K, int mirror,
                                                                  const int* __restrict__ block_offsets,
                                                                  int* __restrict__ pairs) {
    __shared__ int s_warp[33];
    const int n = live_count(n_cap, n_dev);
    const int k = blockIdx.y;
    const int r0 = blockIdx.x * kPairRowsPerBlock;
    int carry = block_offsets[k * gridDim.x + blockIdx.x];
    int* p_in = pairs + (int64_t)k * n_cap;
    int* p_out = pairs + ((int64_t)K + k) * n_cap;
    for (int base = r0; base < r0 + kPairRowsPerBlock; base += kPairThreads) {
        if (base >= n) break;
        int r = base + threadIdx.x;
        int o = r < n ? table_at(table, r, k, K, mirror) : -1;
        int valid = o >= 0;
        int tot;
        int ex = block_exclusive_scan(valid, s_warp, &tot);
        if (valid) {
            p_in[carry + ex] = r;
            p_out[carry + ex] = o;
        }
        carry += tot;
    }
}

// ---- mask-ordered rows for the block-skipping tensor-core tile -------------------------------------------
// One CTA sorts a window of 2048 consecutive rows by their valid-offset bit mask (ties: original order), entirely in
// shared memory (bitonic network on (mask, index) pairs), and writes the permuted table rows + the permutation.
// Windows keep the gather local; measured on LiDAR-like scenes the per-tile union of masks shrinks from 0.56-0.77 of
// the K offsets (spatial order) to 0.32-0.60.
constexpr int kSortWindow = 2048;
constexpr int kSortThreads = 1024;

__global__ void __launch_bounds__(kSortThreads) sort_rows_kernel(const int* __restrict__ nbr, int n_cap,
                                                                 const int* __restrict__ n_dev, int K,
                                                                 int* __restrict__ nbr_sorted, int* __restrict__ out_rows) {
    __shared__ unsigned long long s_mask[kSortWindow];
    __shared__ unsigned short s_idx[kSortWindow];
    const int n = live_count(n_cap, n_dev);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int w0 = blockIdx.x * kSortWindow; w0 < n; w0 += gridDim.x * kSortWindow) {
        // masks: one warp per row, lanes across the offsets (coalesced row reads, ballot = mask)
        for (int i = warp; i < kSortWindow; i += kSortThreads / 32) {
            const int r = w0 + i;
            unsigned long long m = ~0ull;                       // padding rows sort last
            if (r < n) {
                const int* rp = nbr + (int64_t)r * K;
                const int v0 = lane < K ? __ldg(rp + lane) : -1;
                const unsigned lo = __ballot_sync(0xffffffffu, v0 >= 0);
                unsigned hi = 0;
                if (K > 32) {
                    const int v1 = lane + 32 < K ? __ldg(rp + lane + 32) : -1;
                    hi = __ballot_sync(0xffffffffu, v1 >= 0);
                }
                m = (unsigned long long)lo | ((unsigned long long)hi << 32);
            }
            if (lane == 0) { s_mask[i] = m; s_idx[i] = (unsigned short)i; }
        }
        __syncthreads();
        for (int k = 2; k <= kSortWindow; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                const int t = threadIdx.x;
                const int i = 2 * t - (t & (j - 1));            // element with bit j clear
                const int p = i | j;
                const bool up = (i & k) == 0;
                const unsigned long long a = s_mask[i], b = s_mask[p];
                const unsigned short ia = s_idx[i], ib = s_idx[p];
                const bool a_gt_b = a > b || (a == b && ia > ib);
                if (a_gt_b == up) {
                    s_mask[i] = b; s_mask[p] = a;
                    s_idx[i] = ib; s_idx[p] = ia;
                }
                __syncthreads();
            }
        }
        const int live = n - w0 < kSortWindow ? n - w0 : kSortWindow;
        for (int i = warp; i < live; i += kSortThreads / 32) {
            const int src = w0 + (int)s_idx[i];
            const int* rp = nbr + (int64_t)src * K;
            int* wp = nbr_sorted + (int64_t)(w0 + i) * K;
            for (int k = lane; k < K; k += 32) wp[k] = __ldg(rp + k);
            if (lane == 0) out_rows[w0 + i] = src;
        }
        __syncthreads();
    }
}

static int make_geom(ConvGeom& g, int batch, const int* in_shape, const int* out_shape, const int* ksize,
                     const int* stride, const int* padding, const int* dilation, int transposed) {
    g.batch = batch;
    g.in = Shape3{in_shape[0], in_shape[1], in_shape[2]};
    g.out = Shape3{out_shape[0], out_shape[1], out_shape[2]};
    g.transposed = transposed;
    g.K = 1;
    for (int a = 0; a < 3; ++a) {
        g.k[a] = ksize[a];
        g.s[a] = stride ? stride[a] : 1;
        g.p[a] = padding ? padding[a] : 0;
        g.dil[a] = dilation ? dilation[a] : 1;
        if (g.k[a] < 1 || g.s[a] < 1 || g.dil[a] < 1 || g.p[a] < 0) return BTC_E_BADARG;
        g.K *= g.k[a];
    }
    return BTC_OK;
}

}  // namespace btc

using namespace btc;

extern "C" {

int btc_rulebook_subm(const int* coords, int n_cap, const int* n_dev, int batch, const int* shape, const int* ksize,
                      const int* dilation, const uint64_t* index, int64_t n_entries, const int* perm, int* nbr_out,
                      void* stream) {
    if (!shape || !ksize || !index) return badarg("btc_rulebook_subm: null argument");
    if (n_cap > 0 && (!coords || !nbr_out)) return badarg("btc_rulebook_subm: null coordinates");
    if (n_entries != btc_index_entries(batch, shape)) return badarg("btc_rulebook_subm: n_entries mismatch");
    ConvGeom g;
    int one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
    if (make_geom(g, batch, shape, shape, ksize, one, zero, dilation, 0)) return badarg("btc_rulebook_subm: bad geometry");
    if (n_cap <= 0) return BTC_OK;
    if (g.K > 256 || g.k[0] * g.dil[0] > 120 || g.k[1] * g.dil[1] > 120 || g.k[2] * g.dil[2] > 120)
        return badarg("btc_rulebook_subm: kernel too large");
    dim3 blk(32, kSubmSites);
    cudaStream_t st = (cudaStream_t)stream;
    if ((g.K & 1) && g.K / 2 <= 16) {
        fill_table_kernel<<<grid_for((int64_t)n_cap * g.K / 4 + 1, 256), 256, 0, st>>>(nbr_out, n_cap, n_dev, g.K);
        subm_table_sym_kernel<false><<<grid_for((n_cap + 2 * kSubmSites - 1) / (2 * kSubmSites), 1, 8, 8), blk, 0, st>>>(
            (const int4*)coords, n_cap, n_dev, g, (const uint2*)index, perm, nullptr, nullptr, 0u, nbr_out);
    } else {
        subm_table_kernel<false><<<grid_for((n_cap + kSubmSites - 1) / kSubmSites, 1, 8, 8), blk, 0, st>>>(
            (const int4*)coords, n_cap, n_dev, g, (const uint2*)index, perm, nullptr, nullptr, 0u, nbr_out);
    }
    BTC_CHECK_LAUNCH("subm_table");
    return BTC_OK;
}

int btc_rulebook_subm_hash(const int* coords, int n_cap, const int* n_dev, int batch, const int* shape,
                           const int* ksize, const int* dilation, const int64_t* keys, const int* vals,
                           int64_t n_slots, int* nbr_out, void* stream) {
    if (!shape || !ksize || !keys || !vals) return badarg("btc_rulebook_subm_hash: null argument");
    if (n_slots <= 0 || (n_slots & (n_slots - 1))) return badarg("btc_rulebook_subm_hash: n_slots must be a power of two");
    if (n_cap > 0 && (!coords || !nbr_out)) return badarg("btc_rulebook_subm_hash: null coordinates");
    ConvGeom g;
    int one[3] = {1, 1, 1}, zero[3] = {0, 0, 0};
    if (make_geom(g, batch, shape, shape, ksize, one, zero, dilation, 0)) return badarg("btc_rulebook_subm_hash: bad geometry");
    if (n_cap <= 0) return BTC_OK;
    if (g.K > 256 || g.k[0] * g.dil[0] > 120 || g.k[1] * g.dil[1] > 120 || g.k[2] * g.dil[2] > 120)
        return badarg("btc_rulebook_subm_hash: kernel too large");
    dim3 blk(32, kSubmSites);
    cudaStream_t st = (cudaStream_t)stream;
    if ((g.K & 1) && g.K / 2 <= 16) {
        fill_table_kernel<<<grid_for((int64_t)n_cap * g.K / 4 + 1, 256), 256, 0, st>>>(nbr_out, n_cap, n_dev, g.K);
        subm_table_sym_kernel<true><<<grid_for((n_cap + 2 * kSubmSites - 1) / (2 * kSubmSites), 1, 8, 8), blk, 0, st>>>(
            (const int4*)coords, n_cap, n_dev, g, nullptr, nullptr, (const long long*)keys, vals, (uint32_t)(n_slots - 1),
            nbr_out);
    } else {
        subm_table_kernel<true><<<grid_for((n_cap + kSubmSites - 1) / kSubmSites, 1, 8, 8), blk, 0, st>>>(
            (const int4*)coords, n_cap, n_dev, g, nullptr, nullptr, (const long long*)keys, vals, (uint32_t)(n_slots - 1),
            nbr_out);
    }
    BTC_CHECK_LAUNCH("subm_table_hash");
    return BTC_OK;
}

int btc_rulebook_conv(const int* coords_in, int n_in_cap, const int* n_in_dev, int batch, const int* in_shape,
                      const int* out_shape, const int* ksize, const int* stride, const int* padding,
                      const int* dilation, int transposed, uint64_t* out_index, int64_t out_entries, int* out_coords,
                      int out_cap, int* n_out, int* nbr_out, int* nbr_in, void* workspace, int64_t workspace_bytes,
                      void* stream) {
    if (!in_shape || !out_shape || !ksize || !out_index || !n_out || !workspace)
        return badarg("btc_rulebook_conv: null argument");
    if (n_in_cap > 0 && (!coords_in || !out_coords)) return badarg("btc_rulebook_conv: null coordinates");
    if (out_entries != btc_index_entries(batch, out_shape)) return badarg("btc_rulebook_conv: out_entries mismatch");
    if (workspace_bytes < btc_index_workspace_bytes(out_entries)) return badarg("btc_rulebook_conv: workspace too small");
    ConvGeom g;
    if (make_geom(g, batch, in_shape, out_shape, ksize, stride, padding, dilation, transposed))
        return badarg("btc_rulebook_conv: bad geometry");
    cudaStream_t st = (cudaStream_t)stream;
    const int T = 256;
    if (g.k[0] > kMaxAxisTaps || g.k[1] > kMaxAxisTaps || g.k[2] > kMaxAxisTaps)
        return badarg("btc_rulebook_conv: kernel extent above 8 per axis");
    int64_t work = (int64_t)(n_in_cap > 0 ? n_in_cap : 1);   // one thread per input site
    if (n_in_cap > 0) {
        conv_mark_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g,
                                                          (unsigned*)out_index);
        BTC_CHECK_LAUNCH("conv_mark");
    }
    int rc = launch_index_scan((uint2*)out_index, out_entries, (int*)workspace, n_out, st);
    if (rc) return rc;
    if (out_cap > 0 && out_coords)
        conv_emit_kernel<<<grid_for(out_entries, T), T, 0, st>>>((const uint2*)out_index, out_entries, g.out, out_cap,
                                                                (int4*)out_coords);
    if (out_cap > 0 && nbr_out)   // n_out is on the device by now (the scan wrote it)
        fill_table_kernel<<<grid_for((int64_t)out_cap * g.K / 4 + 1, T), T, 0, st>>>(nbr_out, out_cap, n_out, g.K);
    if (n_in_cap > 0 && (nbr_out || nbr_in))
        conv_tables_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g,
                                                            (const uint2*)out_index, out_cap, nbr_out, nbr_in);
    BTC_CHECK_LAUNCH("conv rulebook");
    return BTC_OK;
}

int64_t btc_index_summary_words(int64_t n_entries) { return (n_entries + 31) / 32; }

int64_t btc_rulebook_conv_sparse_workspace_bytes(int64_t n_entries) {
    const int64_t n_sum = (n_entries + 31) / 32;
    const int64_t tiles = (n_sum + kSeTile - 1) / kSeTile;
    return align_up(16 + tiles * 8, 256);
}

int btc_rulebook_conv_sparse(const int* coords_in, int n_in_cap, const int* n_in_dev, int batch, const int* in_shape,
                             const int* out_shape, const int* ksize, const int* stride, const int* padding,
                             const int* dilation, int transposed, uint64_t* out_index, int64_t out_entries,
                             uint32_t* summary, int* out_coords, int out_cap, int* n_out, int* nbr_out, int* nbr_in,
                             void* workspace, int64_t workspace_bytes, void* stream) {
    if (!in_shape || !out_shape || !ksize || !out_index || !summary || !n_out || !workspace)
        return badarg("btc_rulebook_conv_sparse: null argument");
    if (n_in_cap > 0 && (!coords_in || !out_coords)) return badarg("btc_rulebook_conv_sparse: null coordinates");
    if (out_entries != btc_index_entries(batch, out_shape)) return badarg("btc_rulebook_conv_sparse: out_entries mismatch");
    if (workspace_bytes < btc_rulebook_conv_sparse_workspace_bytes(out_entries))
        return badarg("btc_rulebook_conv_sparse: workspace too small");
    ConvGeom g;
    if (make_geom(g, batch, in_shape, out_shape, ksize, stride, padding, dilation, transposed))
        return badarg("btc_rulebook_conv_sparse: bad geometry");
    if (g.k[0] > kMaxAxisTaps || g.k[1] > kMaxAxisTaps || g.k[2] > kMaxAxisTaps)
        return badarg("btc_rulebook_conv_sparse: kernel extent above 8 per axis");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_sum = (out_entries + 31) / 32;
    const int tiles = (int)((n_sum + kSeTile - 1) / kSeTile);
    int* tile_ctr = (int*)workspace;
    unsigned long long* status = (unsigned long long*)((char*)workspace + 16);
    BTC_CUDA(cudaMemsetAsync(workspace, 0, 16 + (size_t)tiles * 8, st), "conv_sparse memset");
    const int64_t work = (int64_t)(n_in_cap > 0 ? n_in_cap : 1);
    const bool fast = fast_geometry(g);
    const AxisFast az{g.k[0], g.s[0], g.p[0], g.out.d}, ay{g.k[1], g.s[1], g.p[1], g.out.h}, ax{g.k[2], g.s[2], g.p[2], g.out.w};
    if (n_in_cap > 0) {
        if (fast)
            conv_mark2_fast_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g.out, az, ay,
                                                                    ax, (unsigned*)out_index, summary);
        else
            conv_mark2_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g,
                                                               (unsigned*)out_index, summary);
        BTC_CHECK_LAUNCH("conv_mark2");
    }
    scan_rank_kernel<<<tiles, kSeThreads, 0, st>>>((uint2*)out_index, summary, (int)n_sum, status, tile_ctr, n_out);
    BTC_CHECK_LAUNCH("scan_rank");
    if (out_cap > 0 && nbr_out)
        fill_table_kernel<<<grid_for((int64_t)out_cap * g.K / 4 + 1, 256), 256, 0, st>>>(nbr_out, out_cap, n_out, g.K);
    if (n_in_cap > 0) {
        if (fast)
            conv_tables2_fast_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g.out, az,
                                                                      ay, ax, g.K, (const uint2*)out_index, out_cap, nbr_out,
                                                                      nbr_in, (int4*)out_coords);
        else
            conv_tables2_kernel<<<grid_for(work, 128), 128, 0, st>>>((const int4*)coords_in, n_in_cap, n_in_dev, g,
                                                                 (const uint2*)out_index, out_cap, nbr_out, nbr_in,
                                                                 (int4*)out_coords);
    }
    BTC_CHECK_LAUNCH("conv rulebook (sparse)");
    return BTC_OK;
}

int btc_index_clear_sparse(const int* coords, int n_cap, const int* n_dev, int batch, const int* shape, uint64_t* index,
                           int64_t n_entries, uint32_t* summary, void* stream) {
    if (!index || !shape || !summary) return badarg("btc_index_clear_sparse: null argument");
    if (n_entries != btc_index_entries(batch, shape)) return badarg("btc_index_clear_sparse: n_entries mismatch");
    if (n_cap <= 0) return BTC_OK;
    if (!coords) return badarg("btc_index_clear_sparse: null coordinates");
    Shape3 s{shape[0], shape[1], shape[2]};
    index_clear2_kernel<<<grid_for(n_cap, 256), 256, 0, (cudaStream_t)stream>>>((const int4*)coords, n_cap, n_dev, s, batch,
                                                                              (uint2*)index, summary);
    BTC_CHECK_LAUNCH("index_clear2");
    return BTC_OK;
}

int64_t btc_rulebook_pairs_workspace_bytes(int n_in_cap, int K) {
    int nblk = (n_in_cap + kPairRowsPerBlock - 1) / kPairRowsPerBlock;
    if (nblk < 1) nblk = 1;
    return align_up((int64_t)nblk * K * 4, 256);
}

int btc_rulebook_pairs(const int* table, int n_in_cap, const int* n_in_dev, int K, int mirror, int* pairs, int* pair_num,
                       void* workspace, int64_t workspace_bytes, void* stream) {
    if (!table || !pairs || !pair_num || !workspace) return badarg("btc_rulebook_pairs: null argument");
    if (K < 1 || n_in_cap < 0) return badarg("btc_rulebook_pairs: bad sizes");
    if (workspace_bytes < btc_rulebook_pairs_workspace_bytes(n_in_cap, K)) return badarg("btc_rulebook_pairs: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_in_cap == 0) {
        BTC_CUDA(cudaMemsetAsync(pair_num, 0, (size_t)K * 4, st), "pairs memset");
        return BTC_OK;
    }
    BTC_CUDA(cudaMemsetAsync(pairs, 0xff, (size_t)2 * K * n_in_cap * 4, st), "pairs memset");
    int nblk = (n_in_cap + kPairRowsPerBlock - 1) / kPairRowsPerBlock;
    dim3 grid(nblk, K);
    int* counts = (int*)workspace;
    pairs_count_kernel<<<grid, kPairThreads, 0, st>>>(table, n_in_cap, n_in_dev, K, mirror, counts);
    pairs_scan_kernel<<<K, 256, 0, st>>>(counts, nblk, pair_num);
    pairs_write_kernel<<<grid, kPairThreads, 0, st>>>(table, n_in_cap, n_in_dev, K, mirror, counts, pairs);
    BTC_CHECK_LAUNCH("rulebook_pairs");
    return BTC_OK;
}

int btc_rulebook_sort_rows(const int* nbr_out, int n_cap, const int* n_dev, int K, int* nbr_sorted, int* out_rows,
                           void* stream) {
    if (K < 1 || K > 64) return badarg("btc_rulebook_sort_rows: K must be in [1, 64]");
    if (n_cap <= 0) return BTC_OK;
    if (!nbr_out || !nbr_sorted || !out_rows) return badarg("btc_rulebook_sort_rows: null argument");
    if (nbr_out == nbr_sorted) return badarg("btc_rulebook_sort_rows: in-place sort is not supported");
    const int windows = (n_cap + kSortWindow - 1) / kSortWindow;
    sort_rows_kernel<<<windows < 4 * kNumSM ? windows : 4 * kNumSM, kSortThreads, 0, (cudaStream_t)stream>>>(
        nbr_out, n_cap, n_dev, K, nbr_sorted, out_rows);
    BTC_CUDA(cudaGetLastError(), "btc_rulebook_sort_rows launch");
    return BTC_OK;
}

}  // extern "C"
