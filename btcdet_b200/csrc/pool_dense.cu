// SparseMaxPool3d, SparseConvTensor.dense() and the sorted re-voxelisation of labelled points.
//   maxpool : spconv 1.2.1 src/spconv/maxpool.cu (maxPoolFwd*/Bwd*) — used by
//             btcdet/models/backbones_3d/spconv_backbone.py:29,831-847 (occ_conv2).
//   dense   : spconv SparseConvTensor.dense() — occ_head_3D.py:46,51, height_compression.py:21.
//   revox   : torch.unique(dim=0)+sort+pad of add_occ_template.py:248-268.
// All three are pure HBM-streaming kernels: coalesced along the channel / cell axis,
// one pass, no atomics on the data path.
#include "common.cuh"

namespace btc {

// out[o][c] = max(0, max_k in[nbr[o][k]][c])  (zero-initialised output)
__global__ void maxpool_fwd_kernel(const float* __restrict__ in, const int* __restrict__ nbr, float* __restrict__ out,
                                   int n_cap, const int* __restrict__ n_dev, int K, int c) {
    const int n = live_count(n_cap, n_dev);
    const int64_t work = (int64_t)n * c;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int o = (int)(t / c), ch = (int)(t - (int64_t)o * c);
        float m = 0.f;
        const int* row = nbr + (int64_t)o * K;
        for (int k = 0; k < K; ++k) {
            int i = __ldg(row + k);
            if (i >= 0) m = fmaxf(m, __ldg(in + (int64_t)i * c + ch));
        }
        out[t] = m;
    }
}

// d_in[i][c] += d_out[o][c] for every pair with in[i][c] == out[o][c] (spconv maxPoolBwd).
__global__ void maxpool_bwd_kernel(const float* __restrict__ in, const float* __restrict__ out,
                                   const float* __restrict__ d_out, const int* __restrict__ nbr,
                                   float* __restrict__ d_in, int n_cap, const int* __restrict__ n_dev, int K, int c) {
    const int n = live_count(n_cap, n_dev);
    const int64_t work = (int64_t)n * c;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int o = (int)(t / c), ch = (int)(t - (int64_t)o * c);
        float m = out[t], g = d_out[t];
        const int* row = nbr + (int64_t)o * K;
        for (int k = 0; k < K; ++k) {
            int i = __ldg(row + k);
            if (i >= 0 && __ldg(in + (int64_t)i * c + ch) == m) atomicAdd(d_in + (int64_t)i * c + ch, g);
        }
    }
}

// dense[b][ch][z][y][x] = feat[i][ch].  One warp per site iterates channels so that the
// feature row is read coalesced; the writes are one 4-byte store per (site, channel) into
// channel planes (inherent to the channels-first layout the reference asks for).
__global__ void to_dense_kernel(const float* __restrict__ feat, const int4* __restrict__ coords, int n_cap,
                                const int* __restrict__ n_dev, int c, int batch, Shape3 s, float* __restrict__ out) {
    const int n = live_count(n_cap, n_dev);
    const int64_t plane = (int64_t)s.d * s.h * s.w;
    const int64_t work = (int64_t)n * c;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t / c), ch = (int)(t - (int64_t)i * c);
        int4 q = __ldg(coords + i);
        if ((unsigned)q.x >= (unsigned)batch || (unsigned)q.y >= (unsigned)s.d || (unsigned)q.z >= (unsigned)s.h ||
            (unsigned)q.w >= (unsigned)s.w)
            continue;
        int64_t cell = ((int64_t)q.y * s.h + q.z) * s.w + q.w;
        out[((int64_t)q.x * c + ch) * plane + cell] = __ldg(feat + t);
    }
}

__global__ void from_dense_kernel(const float* __restrict__ d_out, const int4* __restrict__ coords, int n_cap,
                                  const int* __restrict__ n_dev, int c, int batch, Shape3 s,
                                  float* __restrict__ d_feat) {
    const int n = live_count(n_cap, n_dev);
    const int64_t plane = (int64_t)s.d * s.h * s.w;
    const int64_t work = (int64_t)n * c;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t / c), ch = (int)(t - (int64_t)i * c);
        int4 q = __ldg(coords + i);
        float v = 0.f;
        if ((unsigned)q.x < (unsigned)batch && (unsigned)q.y < (unsigned)s.d && (unsigned)q.z < (unsigned)s.h &&
            (unsigned)q.w < (unsigned)s.w) {
            int64_t cell = ((int64_t)q.y * s.h + q.z) * s.w + q.w;
            v = __ldg(d_out + ((int64_t)q.x * c + ch) * plane + cell);
        }
        d_feat[t] = v;
    }
}

// ---- sorted re-voxelisation ------------------------------------------------------------------
__global__ void revox_mark_kernel(const int4* __restrict__ pc, int n_cap, const int* __restrict__ n_dev, Shape3 s,
                                  int batch, unsigned* __restrict__ words) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 q = __ldg(pc + i);
        if ((unsigned)q.x >= (unsigned)batch || (unsigned)q.y >= (unsigned)s.d || (unsigned)q.z >= (unsigned)s.h ||
            (unsigned)q.w >= (unsigned)s.w)
            continue;
        int64_t key = flat_key(q.x, q.y, q.z, q.w, s);
        unsigned bit = 1u << (unsigned)(key & 31);
        unsigned* w = words + 2 * (key >> 5);
        if (!(*(volatile unsigned*)w & bit)) atomicOr(w, bit);
    }
}

__global__ void revox_emit_kernel(const uint2* __restrict__ index, int64_t n_entries, Shape3 s, int vox_cap,
                                  int4* __restrict__ vox_coords) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n_entries; w += (int64_t)gridDim.x * blockDim.x) {
        uint2 e = __ldg(index + w);
        unsigned bits = e.x;
        int row = (int)e.y;
        while (bits) {
            int bpos = __ffs(bits) - 1;
            bits &= bits - 1;
            if (row < vox_cap) {
                int64_t key = (w << 5) + bpos;
                int x = (int)(key % s.w);
                int64_t t = key / s.w;
                int y = (int)(t % s.h);
                t /= s.h;
                vox_coords[row] = make_int4((int)(t / s.d), (int)(t % s.d), y, x);
            }
            ++row;
        }
    }
}

// pt_voxel[i] = rank of the point's cell; count per voxel.
__global__ void revox_assign_kernel(const int4* __restrict__ pc, int n_cap, const int* __restrict__ n_dev, Shape3 s,
                                    int batch, const uint2* __restrict__ index, int vox_cap,
                                    int* __restrict__ pt_voxel, int* __restrict__ vox_count) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int4 q = __ldg(pc + i);
        int v = -1;
        if ((unsigned)q.x < (unsigned)batch && (unsigned)q.y < (unsigned)s.d && (unsigned)q.z < (unsigned)s.h &&
            (unsigned)q.w < (unsigned)s.w) {
            v = index_lookup(index, flat_key(q.x, q.y, q.z, q.w, s));
            if (v >= vox_cap) v = -1;
        }
        pt_voxel[i] = v;
        if (v >= 0) atomicAdd(vox_count + v, 1);
    }
}

// Stable slot of a point inside its voxel = number of earlier points (smaller input index) in
// the same voxel.  The voxel's point list is recovered through a per-voxel cursor walk:
// points of a voxel are few (<= ~20), so each point counts its predecessors among a compact
// per-voxel member list built with atomics and then ranks itself by index.
__global__ void revox_members_kernel(const int* __restrict__ pt_voxel, int n_cap, const int* __restrict__ n_dev,
                                     const int* __restrict__ vox_start, int* __restrict__ vox_cursor,
                                     int* __restrict__ members) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int v = pt_voxel[i];
        if (v < 0) continue;
        int pos = atomicAdd(vox_cursor + v, 1);
        members[vox_start[v] + pos] = i;
    }
}

__global__ void revox_slots_kernel(const int* __restrict__ pt_voxel, int n_cap, const int* __restrict__ n_dev,
                                   const int* __restrict__ vox_start, const int* __restrict__ vox_count,
                                   const int* __restrict__ members, int* __restrict__ slots) {
    const int n = live_count(n_cap, n_dev);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int v = pt_voxel[i];
        if (v < 0) {
            slots[i] = -1;
            continue;
        }
        const int* m = members + vox_start[v];
        int cnt = vox_count[v], s = 0;
        for (int j = 0; j < cnt; ++j) s += (m[j] < i);
        slots[i] = s;
    }
}

__global__ void revox_max_kernel(const int* __restrict__ vox_count, const int* __restrict__ n_voxels, int vox_cap,
                                 int* __restrict__ max_count) {
    int n = *n_voxels;
    if (n > vox_cap) n = vox_cap;
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, vox_count[i]);
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_down_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_count, m);
}

__global__ void revox_fill_kernel(const float* __restrict__ pt_feat, const int* __restrict__ pt_voxel,
                                  const int* __restrict__ slots, int n_cap, const int* __restrict__ n_dev, int c,
                                  int p_max, float* __restrict__ voxels, int vox_rows) {
    const int n = live_count(n_cap, n_dev);
    const int64_t work = (int64_t)n * c;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(t / c), ch = (int)(t - (int64_t)i * c);
        int v = pt_voxel[i], s = slots[i];
        if (v < 0 || v >= vox_rows || s < 0 || s >= p_max) continue;
        voxels[((int64_t)v * p_max + s) * c + ch] = __ldg(pt_feat + t);
    }
}

// dst[0 : n*row_words] = src[...] for the live rows only (n read on the device): stages a capacity-sized result
// buffer into a compact one without a host round-trip.
__global__ void copy_rows_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int n_cap,
                                 const int* __restrict__ n_dev, int row_vec4) {
    const int64_t work = (int64_t)live_count(n_cap, n_dev) * row_vec4;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = __ldg(src + t);
}

}  // namespace btc

using namespace btc;

extern "C" {

int btc_maxpool_fwd(const float* feat_in, const int* nbr_out, float* feat_out, int n_out_cap, const int* n_out_dev,
                    int K, int c, void* stream) {
    if (n_out_cap <= 0) return BTC_OK;   // empty output (null data pointers of 0-row tensors are fine)
    if (!nbr_out || !feat_out) return badarg("btc_maxpool_fwd: null argument");
    if (!feat_in) return badarg("btc_maxpool_fwd: null feat_in");
    maxpool_fwd_kernel<<<grid_for((int64_t)n_out_cap * c, 256), 256, 0, (cudaStream_t)stream>>>(
        feat_in, nbr_out, feat_out, n_out_cap, n_out_dev, K, c);
    BTC_CHECK_LAUNCH("maxpool_fwd");
    return BTC_OK;
}

int btc_maxpool_bwd(const float* feat_in, const float* feat_out, const float* d_out, const int* nbr_out, float* d_in,
                    int n_in, int n_out_cap, const int* n_out_dev, int K, int c, void* stream) {
    if (n_in <= 0) return BTC_OK;
    if (!d_in || (n_out_cap > 0 && !nbr_out)) return badarg("btc_maxpool_bwd: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_in > 0) BTC_CUDA(cudaMemsetAsync(d_in, 0, (size_t)n_in * c * 4, st), "maxpool_bwd memset");
    if (n_out_cap <= 0 || n_in <= 0) return BTC_OK;
    maxpool_bwd_kernel<<<grid_for((int64_t)n_out_cap * c, 256), 256, 0, st>>>(feat_in, feat_out, d_out, nbr_out, d_in,
                                                                             n_out_cap, n_out_dev, K, c);
    BTC_CHECK_LAUNCH("maxpool_bwd");
    return BTC_OK;
}

int btc_copy_rows(const void* src, void* dst, int n_cap, const int* n_dev, int row_bytes, void* stream) {
    if (!src || !dst) return badarg("btc_copy_rows: null argument");
    if (row_bytes <= 0 || row_bytes % 16 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15))
        return badarg("btc_copy_rows: rows must be 16-byte multiples and 16-byte aligned");
    if (n_cap <= 0) return BTC_OK;
    copy_rows_kernel<<<grid_for((int64_t)n_cap * (row_bytes / 16), 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)src, (uint4*)dst, n_cap, n_dev, row_bytes / 16);
    BTC_CHECK_LAUNCH("copy_rows");
    return BTC_OK;
}

int btc_to_dense(const float* feat, const int* coords, int n_cap, const int* n_dev, int c, int batch, const int* shape,
                 float* out, void* stream) {
    if (!shape || !out) return badarg("btc_to_dense: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    Shape3 s{shape[0], shape[1], shape[2]};
    size_t bytes = (size_t)batch * c * s.d * s.h * s.w * sizeof(float);
    BTC_CUDA(cudaMemsetAsync(out, 0, bytes, st), "to_dense memset");
    if (n_cap <= 0) return BTC_OK;
    if (!feat || !coords) return badarg("btc_to_dense: null features");
    to_dense_kernel<<<grid_for((int64_t)n_cap * c, 256), 256, 0, st>>>(feat, (const int4*)coords, n_cap, n_dev, c, batch, s,
                                                                      out);
    BTC_CHECK_LAUNCH("to_dense");
    return BTC_OK;
}

int btc_from_dense(const float* d_out, const int* coords, int n_cap, const int* n_dev, int c, int batch,
                   const int* shape, float* d_feat, void* stream) {
    if (!shape || !d_out || !d_feat || !coords) return badarg("btc_from_dense: null argument");
    if (n_cap <= 0) return BTC_OK;
    Shape3 s{shape[0], shape[1], shape[2]};
    from_dense_kernel<<<grid_for((int64_t)n_cap * c, 256), 256, 0, (cudaStream_t)stream>>>(
        d_out, (const int4*)coords, n_cap, n_dev, c, batch, s, d_feat);
    BTC_CHECK_LAUNCH("from_dense");
    return BTC_OK;
}

// workspace: [scan block sums for the index][scan block sums for counts][vox_start cap+1][cursor cap][members n]
static int64_t revox_ws(int n_points, int64_t n_entries, int vox_cap, int64_t* off_counts, int64_t* off_start,
                        int64_t* off_cursor, int64_t* off_members) {
    int64_t off = 0;
    off += btc_index_workspace_bytes(n_entries);
    *off_counts = off;
    off += align_up((int64_t)(scan_num_blocks(vox_cap > 0 ? vox_cap : 1) + 2) * 4, 256);
    *off_start = off;
    off += align_up((int64_t)(vox_cap + 1) * 4, 256);
    *off_cursor = off;
    off += align_up((int64_t)(vox_cap + 1) * 4, 256);
    *off_members = off;
    off += align_up((int64_t)(n_points > 0 ? n_points : 1) * 4, 256);
    return off;
}

int64_t btc_revoxelize_workspace_bytes(int n_points, int64_t n_entries) {
    int64_t a, b, c, d;
    // vox_cap <= n_points
    return revox_ws(n_points, n_entries, n_points, &a, &b, &c, &d);
}

// exclusive scan of counts (not flags): small dedicated kernels
__global__ void __launch_bounds__(1024) counts_scan_single_kernel(const int* __restrict__ counts, const int* __restrict__ n_dev,
                                                                  int cap, int* __restrict__ start) {
    __shared__ int s_warp[33];
    int n = *n_dev;
    if (n > cap) n = cap;
    int carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n ? counts[i] : 0;
        int tot;
        int ex = block_exclusive_scan(v, s_warp, &tot);
        if (i < n) start[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) start[n] = carry;
}

int btc_revoxelize(const int* pt_coords, int n_cap, const int* n_dev, int batch, const int* shape, uint64_t* index,
                   int64_t n_entries, int* vox_coords, int vox_cap, int* vox_count, int* slots, int* pt_voxel,
                   int* n_voxels, int* max_count, void* workspace, int64_t workspace_bytes, void* stream) {
    if (!pt_coords || !shape || !index || !vox_coords || !vox_count || !slots || !pt_voxel || !n_voxels || !max_count ||
        !workspace)
        return badarg("btc_revoxelize: null argument");
    if (n_entries != btc_index_entries(batch, shape)) return badarg("btc_revoxelize: n_entries mismatch");
    if (vox_cap > n_cap) vox_cap = n_cap;
    int64_t o_counts, o_start, o_cursor, o_members;
    int64_t need = revox_ws(n_cap, n_entries, vox_cap, &o_counts, &o_start, &o_cursor, &o_members);
    if (workspace_bytes < need) return badarg("btc_revoxelize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int* vox_start = (int*)(ws + o_start);
    int* vox_cursor = (int*)(ws + o_cursor);
    int* members = (int*)(ws + o_members);
    Shape3 s{shape[0], shape[1], shape[2]};
    const int T = 256;
    BTC_CUDA(cudaMemsetAsync(max_count, 0, 4, st), "revox memset");
    if (vox_cap > 0) {
        BTC_CUDA(cudaMemsetAsync(vox_count, 0, (size_t)vox_cap * 4, st), "revox memset");
        BTC_CUDA(cudaMemsetAsync(vox_cursor, 0, (size_t)vox_cap * 4, st), "revox memset");
    }
    if (n_cap > 0)
        revox_mark_kernel<<<grid_for(n_cap, T), T, 0, st>>>((const int4*)pt_coords, n_cap, n_dev, s, batch, (unsigned*)index);
    int rc = launch_index_scan((uint2*)index, n_entries, (int*)ws, n_voxels, st);
    if (rc) return rc;
    if (n_cap <= 0 || vox_cap <= 0) return BTC_OK;
    revox_emit_kernel<<<grid_for(n_entries, T), T, 0, st>>>((const uint2*)index, n_entries, s, vox_cap, (int4*)vox_coords);
    revox_assign_kernel<<<grid_for(n_cap, T), T, 0, st>>>((const int4*)pt_coords, n_cap, n_dev, s, batch,
                                                         (const uint2*)index, vox_cap, pt_voxel, vox_count);
    counts_scan_single_kernel<<<1, 1024, 0, st>>>(vox_count, n_voxels, vox_cap, vox_start);
    revox_members_kernel<<<grid_for(n_cap, T), T, 0, st>>>(pt_voxel, n_cap, n_dev, vox_start, vox_cursor, members);
    revox_slots_kernel<<<grid_for(n_cap, T), T, 0, st>>>(pt_voxel, n_cap, n_dev, vox_start, vox_count, members, slots);
    revox_max_kernel<<<grid_for(vox_cap, T, 1, 2), T, 0, st>>>(vox_count, n_voxels, vox_cap, max_count);
    BTC_CHECK_LAUNCH("revoxelize");
    return BTC_OK;
}

int btc_revoxelize_fill(const float* pt_feat, const int* pt_voxel, const int* slots, int n_cap, const int* n_dev, int c,
                        int p_max, float* voxels, int vox_rows, void* stream) {
    if (!voxels) return badarg("btc_revoxelize_fill: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (vox_rows > 0 && p_max > 0)
        BTC_CUDA(cudaMemsetAsync(voxels, 0, (size_t)vox_rows * p_max * c * 4, st), "revox_fill memset");
    if (n_cap <= 0 || vox_rows <= 0 || p_max <= 0) return BTC_OK;
    if (!pt_feat || !pt_voxel || !slots) return badarg("btc_revoxelize_fill: null argument");
    revox_fill_kernel<<<grid_for((int64_t)n_cap * c, 256), 256, 0, st>>>(pt_feat, pt_voxel, slots, n_cap, n_dev, c,
                                                                           p_max, voxels, vox_rows);
    BTC_CHECK_LAUNCH("revox_fill");
    return BTC_OK;
}

}  // extern "C"
