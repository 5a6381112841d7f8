// RoI grid pooling of the second stage: the native ops under `ConvHead.roi_conv_pool`
// (btcdet/models/roi_heads/conv_head.py:247-379).
//
//   ball_query_kernel      stacked ball query of `StackSAModuleMSG` (pointnet2_stack/src/ball_query_gpu.cu:16-60), one
//                          WARP per group of four query points and ALL radii of the module in one pass over the scene's
//                          points (the reference: one thread per query, one launch per radius, O(M*N) serial loads).
//                          Same result bit for bit: the first `nsample` points in index order with d2 < r*r, d2 evaluated
//                          in the reference kernel's own operation order (FMUL, FFMA, FFMA — nvcc contracts its source).
//   group_points(_grad)    stacked grouping (group_points_gpu.cu:16-95).
//   index_volume / tri_flag / tri_emit
//                          `reverse_sparse_trilinear_interpolate_torch` (btcdet/utils/common_utils.py:247-311) followed by
//                          the non-zero-row compaction of `interpolate_from_3d_features` (conv_head.py:505-528), WITHOUT the
//                          dense [B, C, Z, Y, X] volume and without the eight [T, C] corner tensors: a 4-byte row-index
//                          volume, one warp per target, products and sums in the reference's order (separately rounded
//                          multiplies and adds, corner order 000 010 001 011 100 110 101 111) => bit-exact rows.
// Kernels only (no launches): this header is also compiled for the host by tests/host_emul/ (a lock-step warp
// emulation used by the CPU test-suite to check the kernels' logic against the oracle without a GPU; never shipped).
#pragma once
#include "common.cuh"

namespace btc {
namespace roi {

constexpr int kBqQueries = 4;   // queries a warp scans together (they share every loaded point)
constexpr int kBqMaxRadii = 4;  // radii of one StackSAModuleMSG (yaml: 4 for raw points, 3 for occupancy points)
constexpr int kBqFull = 0x3fffffff;

struct BallArgs {
    int n_radii;
    float r2[kBqMaxRadii];
    int nsample[kBqMaxRadii];
    int* idx[kBqMaxRadii];
};

// The reference's walk (ball_query_gpu.cu:21-30): the scene of stacked query q, the first row and the number of that
// scene's points.
__device__ __forceinline__ void scene_of(int q, int B, const int* __restrict__ q_cnt, const int* __restrict__ p_cnt,
                                         int& scene, int& start, int& n) {
    int bs = 0, acc = __ldg(q_cnt);
    for (int k = 1; k < B; ++k) {
        if (q < acc) break;
        acc += __ldg(q_cnt + k);
        bs = k;
    }
    int s = 0;
    for (int k = 0; k < bs; ++k) s += __ldg(p_cnt + k);
    scene = bs;
    start = s;
    n = __ldg(p_cnt + bs);
}

// Scans the n points of one scene for queries [q0, q0 + nq), nq <= QN, all of that scene.
template <int QN>
__device__ __forceinline__ void ball_scan(const BallArgs& a, int q0, int nq, const float* __restrict__ new_xyz,
                                          const float* __restrict__ pts, int n, int lane) {
    float qx[QN], qy[QN], qz[QN];
    int cnt[QN][kBqMaxRadii], first[QN][kBqMaxRadii];
#pragma unroll
    for (int i = 0; i < QN; ++i) {
        const int q = q0 + (i < nq ? i : 0);
        qx[i] = __ldg(new_xyz + 3 * (int64_t)q + 0);
        qy[i] = __ldg(new_xyz + 3 * (int64_t)q + 1);
        qz[i] = __ldg(new_xyz + 3 * (int64_t)q + 2);
#pragma unroll
        for (int r = 0; r < kBqMaxRadii; ++r) {
            cnt[i][r] = i < nq ? 0 : kBqFull;
            first[i][r] = 0;
        }
    }
    const unsigned lt = (1u << lane) - 1u;
    for (int k0 = 0; k0 < n; k0 += 32) {
        const int k = k0 + lane;
        const bool in = k < n;
        float x = 0.f, y = 0.f, z = 0.f;
        if (in) {
            x = __ldg(pts + 3 * (int64_t)k + 0);
            y = __ldg(pts + 3 * (int64_t)k + 1);
            z = __ldg(pts + 3 * (int64_t)k + 2);
        }
        bool all_full = true;
#pragma unroll
        for (int i = 0; i < QN; ++i) {
            const float dx = __fsub_rn(qx[i], x), dy = __fsub_rn(qy[i], y), dz = __fsub_rn(qz[i], z);
            const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
#pragma unroll
            for (int r = 0; r < kBqMaxRadii; ++r) {
                const int ns = a.nsample[r];          // 0 beyond n_radii: never scanned
                if (cnt[i][r] < ns) {                 // warp-uniform: the counters are ballot sums
                    const unsigned m = __ballot_sync(0xffffffffu, in && d2 < a.r2[r]);
                    if (m) {
                        if (cnt[i][r] == 0) first[i][r] = k0 + __ffs(m) - 1;
                        const int pos = cnt[i][r] + __popc(m & lt);
                        if (((m >> lane) & 1u) && pos < ns) a.idx[r][(int64_t)(q0 + i) * ns + pos] = k;
                        cnt[i][r] += __popc(m);
                    }
                    if (cnt[i][r] < ns) all_full = false;
                }
            }
        }
        if (all_full) break;
    }
    // Rows shorter than nsample repeat the first hit; an empty ball is (-1, 0, 0, ...) — what the reference leaves in its
    // zero-initialised idx (pointnet2_utils.py:33, ball_query_gpu.cu:47-58).
#pragma unroll
    for (int i = 0; i < QN; ++i) {
        if (i >= nq) continue;
#pragma unroll
        for (int r = 0; r < kBqMaxRadii; ++r) {
            if (r >= a.n_radii) continue;
            const int ns = a.nsample[r];
            int* row = a.idx[r] + (int64_t)(q0 + i) * ns;
            const int c = cnt[i][r];
            if (c == 0) {
                for (int l = lane; l < ns; l += 32) row[l] = l == 0 ? -1 : 0;
            } else {
                for (int l = c + lane; l < ns; l += 32) row[l] = first[i][r];
            }
        }
    }
}

__global__ void __launch_bounds__(256) ball_query_kernel(int B, int M, BallArgs a, const float* __restrict__ new_xyz,
                                                         const int* __restrict__ q_cnt, const float* __restrict__ xyz,
                                                         const int* __restrict__ p_cnt) {
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * (int64_t)blockDim.x) >> 5);
    const int ngroups = (M + kBqQueries - 1) / kBqQueries;
    for (int g = warp; g < ngroups; g += nwarps) {
        const int q0 = g * kBqQueries;
        const int nq = min(kBqQueries, M - q0);
        int s0, st0, n0, s1, st1, n1;
        scene_of(q0, B, q_cnt, p_cnt, s0, st0, n0);
        scene_of(q0 + nq - 1, B, q_cnt, p_cnt, s1, st1, n1);
        if (s0 == s1) {
            ball_scan<kBqQueries>(a, q0, nq, new_xyz, xyz + 3 * (int64_t)st0, n0, lane);
        } else {                                  // the group straddles a scene boundary: one query at a time
            for (int i = 0; i < nq; ++i) {
                scene_of(q0 + i, B, q_cnt, p_cnt, s1, st1, n1);
                ball_scan<1>(a, q0 + i, 1, new_xyz, xyz + 3 * (int64_t)st1, n1, lane);
            }
        }
    }
}

__global__ void group_points_kernel(int B, int64_t total, int C, int ns, const float* __restrict__ feat,
                                    const int* __restrict__ f_cnt, const int* __restrict__ idx,
                                    const int* __restrict__ i_cnt, float* __restrict__ out) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(e % ns);
        const int c = (int)((e / ns) % C);
        const int m = (int)(e / ns / C);
        int scene, start, n;
        scene_of(m, B, i_cnt, f_cnt, scene, start, n);
        const int j = __ldg(idx + (int64_t)m * ns + s);
        out[e] = (j >= 0 && j < n) ? __ldg(feat + ((int64_t)start + j) * C + c) : 0.f;
    }
}

__global__ void group_points_grad_kernel(int B, int64_t total, int C, int ns, const float* __restrict__ grad_out,
                                         const int* __restrict__ idx, const int* __restrict__ i_cnt,
                                         const int* __restrict__ f_cnt, float* __restrict__ grad_feat) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(e % ns);
        const int c = (int)((e / ns) % C);
        const int m = (int)(e / ns / C);
        int scene, start, n;
        scene_of(m, B, i_cnt, f_cnt, scene, start, n);
        const int j = __ldg(idx + (int64_t)m * ns + s);
        if (j >= 0 && j < n) atomicAdd(grad_feat + ((int64_t)start + j) * C + c, __ldg(grad_out + e));
    }
}

// ---- reverse trilinear gather from a sparse tensor ---------------------------------------------------------------
struct TriGeom {
    int B, Z, Y, X, C;
    int normalize;
    int64_t T, per_scene;
};

__global__ void index_volume_kernel(const int4* __restrict__ coords, int n_cap, const int* __restrict__ n_dev, TriGeom g,
                                    int* __restrict__ vol) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int4 c = __ldg(coords + i);   // (b, z, y, x)
        if ((unsigned)c.x < (unsigned)g.B && (unsigned)c.y < (unsigned)g.Z && (unsigned)c.z < (unsigned)g.Y &&
            (unsigned)c.w < (unsigned)g.X)
            vol[(((int64_t)c.x * g.Z + c.y) * g.Y + c.z) * g.X + c.w] = i;
    }
}

// Lane j < 8 evaluates corner j of target t in the reference's term order (common_utils.py:303-310):
//   j : 0 = 000, 1 = 010, 2 = 001, 3 = 011, 4 = 100, 5 = 110, 6 = 101, 7 = 111   (digits: z y x, 0 = floor, 1 = floor + 1)
// Returns the feature row of the corner (-1: outside the grid or not an active site) and its weight.
__device__ __forceinline__ int tri_corner(int j, const float* __restrict__ zyx, const long long* __restrict__ b_target,
                                          int64_t t, const TriGeom& g, const int* __restrict__ vol, float& w) {
    const float z = __ldg(zyx + 3 * t + 0), y = __ldg(zyx + 3 * t + 1), x = __ldg(zyx + 3 * t + 2);
    const int zb = (j >> 2) & 1, yb = j & 1, xb = (j >> 1) & 1;
    const float z0 = floorf(z), y0 = floorf(y), x0 = floorf(x);
    // weight of corner (zb, yb, xb): |(z_other - z) * (y_other - y) * (x_other - x)| with "other" = the opposite corner,
    // evaluated as ((dz * dy) * dx) in fp32 like the reference's tensor expression
    const float oz = zb ? z0 : __fadd_rn(z0, 1.f), oy = yb ? y0 : __fadd_rn(y0, 1.f), ox = xb ? x0 : __fadd_rn(x0, 1.f);
    w = fabsf(__fmul_rn(__fmul_rn(__fsub_rn(oz, z), __fsub_rn(oy, y)), __fsub_rn(ox, x)));
    float cz = zb ? __fadd_rn(z0, 1.f) : z0, cy = yb ? __fadd_rn(y0, 1.f) : y0, cx = xb ? __fadd_rn(x0, 1.f) : x0;
    if (g.normalize) {   // masks are 1, the clamped corner is read (border replication)
        cz = fminf(fmaxf(cz, 0.f), (float)(g.Z - 1));
        cy = fminf(fmaxf(cy, 0.f), (float)(g.Y - 1));
        cx = fminf(fmaxf(cx, 0.f), (float)(g.X - 1));
    }
    if (!(cz >= 0.f && cz < (float)g.Z && cy >= 0.f && cy < (float)g.Y && cx >= 0.f && cx < (float)g.X)) return -1;
    const int64_t b = b_target ? (int64_t)__ldg(b_target + t) : t / g.per_scene;
    if (b < 0 || b >= g.B) return -1;
    return __ldg(vol + ((b * g.Z + (int)cz) * g.Y + (int)cy) * g.X + (int)cx);
}

// One warp per target: flag = any(|row| > 0) (conv_head.py:526); with `emit` the row goes to out[rank[t]].
template <bool EMIT>
__global__ void __launch_bounds__(256) tri_kernel(const float* __restrict__ feats, const float* __restrict__ zyx,
                                                  const long long* __restrict__ b_target, TriGeom g,
                                                  const int* __restrict__ vol, int* __restrict__ flags,
                                                  const int* __restrict__ rank, int P, int ly, int lx, int out_cap,
                                                  float* __restrict__ out_feats, int4* __restrict__ out_coords,
                                                  long long* __restrict__ out_target) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (gridDim.x * (int64_t)blockDim.x) >> 5;
    for (int64_t t = warp; t < g.T; t += nwarps) {
        int o = 0;
        if (EMIT) {
            if (!__ldg(flags + t)) continue;
            o = __ldg(rank + t);
            if (o >= out_cap) continue;
        }
        float w = 0.f;
        int row = -1;
        if (lane < 8) row = tri_corner(lane, zyx, b_target, t, g, vol, w);
        const unsigned live = __ballot_sync(0xffffffffu, row >= 0) & 0xffu;
        if (!EMIT && live == 0) {
            if (lane == 0) flags[t] = 0;
            continue;
        }
        int rows[8];
        float ws[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            rows[j] = __shfl_sync(0xffffffffu, row, j);
            ws[j] = __shfl_sync(0xffffffffu, w, j);
        }
        bool nz = false;
        for (int c = lane; c < g.C; c += 32) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (rows[j] >= 0) acc = __fadd_rn(acc, __fmul_rn(__ldg(feats + (int64_t)rows[j] * g.C + c), ws[j]));
            if (EMIT) out_feats[(int64_t)o * g.C + c] = acc;
            nz = nz || fabsf(acc) > 0.f;
        }
        if (!EMIT) {
            const unsigned any = __ballot_sync(0xffffffffu, nz);
            if (lane == 0) flags[t] = any ? 1 : 0;
        } else if (lane == 0) {
            const int cell = (int)(t % P);
            out_coords[o] = make_int4((int)(t / P), cell / (ly * lx), (cell / lx) % ly, cell % lx);
            if (out_target) out_target[o] = t;
        }
    }
}

// Backward of the emitted rows with respect to the sparse source features: for output row o (target out_target[o]),
// grad_feats[row_j] += w_j * grad_out[o] over the (up to eight) active corners j — the adjoint of tri_kernel's sum;
// fp32 atomics, like the index_put / gather backward of the reference's autograd graph.
__global__ void __launch_bounds__(256) tri_grad_kernel(const float* __restrict__ grad_out, const long long* __restrict__ out_target,
                                                       int n_out, const int* __restrict__ n_dev, const float* __restrict__ zyx,
                                                       const long long* __restrict__ b_target, TriGeom g,
                                                       const int* __restrict__ vol, float* __restrict__ grad_feats) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (gridDim.x * (int64_t)blockDim.x) >> 5;
    const int n = live_count(n_out, n_dev);
    for (int64_t o = warp; o < n; o += nwarps) {
        const int64_t t = __ldg(out_target + o);
        float w = 0.f;
        int row = -1;
        if (lane < 8 && t >= 0 && t < g.T) row = tri_corner(lane, zyx, b_target, t, g, vol, w);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int rj = __shfl_sync(0xffffffffu, row, j);
            const float wj = __shfl_sync(0xffffffffu, w, j);
            if (rj < 0) continue;
            for (int c = lane; c < g.C; c += 32)
                atomicAdd(grad_feats + (int64_t)rj * g.C + c, __fmul_rn(__ldg(grad_out + o * g.C + c), wj));
        }
    }
}

}  // namespace roi
}  // namespace btc
