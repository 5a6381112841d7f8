// Deterministic GPU point -> voxel grouping, bit-exact with the sequential first-come
// algorithm of spconv 1.2.1 `points_to_voxel_3d_np` (include/spconv/point2voxel.h), which the
// reference calls through spconv.utils.VoxelGeneratorV2
// (btcdet/datasets/processor/data_processor.py:68-73,85 / :112-117,136 / :165-170,177).
//
// Parallel restatement of the sequential loop:
//   1. hash-insert each in-range point's cell key; atomicMin keeps the smallest point index
//      per cell ("who saw the voxel first");
//   2. flag points that are the first of their cell, exclusive-scan the flags in point order:
//      the rank of a first-point *is* the sequential voxel id (first-come order);
//   3. per scene, ids >= max_voxels are dropped (the sequential loop skips points that would
//      open a new voxel once voxel_num == max_voxels, but still appends to existing voxels);
//   4. slot inside a voxel = rank of the point index among the voxel's points: a cascade of
//      atomicMin over a per-voxel sorted list of max_points entries leaves the k-th smallest
//      index in slot k regardless of execution order;
//   5. gather rows, zero padding, counts, optional MeanVFE.
// All integer outputs are order-independent, hence reproducible run to run.
#include "common.cuh"

namespace btc {

constexpr long long kEmptyKey = -1LL;
constexpr int kBigIdx = 0x7f7f7f7f;  // memset(0x7f) pattern, larger than any point index

struct VoxGeom {
    float vs[3];
    float lo[3];
    int grid[3];  // x, y, z
};

__device__ __forceinline__ int scene_of(int i, const int* __restrict__ offs, int n_scenes) {
    int b = 0;
    while (b + 1 < n_scenes && i >= __ldg(offs + b + 1)) ++b;
    return b;
}

// c = floor((p - lo) / vs) in IEEE fp32, exactly as the C++ reference evaluates it.
__device__ __forceinline__ bool quantize(const float* __restrict__ p, const VoxGeom& g, int& cx, int& cy, int& cz) {
    float fx = floorf(__fdiv_rn(__fsub_rn(p[0], g.lo[0]), g.vs[0]));
    float fy = floorf(__fdiv_rn(__fsub_rn(p[1], g.lo[1]), g.vs[1]));
    float fz = floorf(__fdiv_rn(__fsub_rn(p[2], g.lo[2]), g.vs[2]));
    // comparisons in float first: NaN and +-huge fail them and are dropped
    if (!(fx >= 0.f && fx < (float)g.grid[0])) return false;
    if (!(fy >= 0.f && fy < (float)g.grid[1])) return false;
    if (!(fz >= 0.f && fz < (float)g.grid[2])) return false;
    cx = (int)fx;
    cy = (int)fy;
    cz = (int)fz;
    return true;
}

__global__ void vox_hash_insert_kernel(const float* __restrict__ points, int n, int n_feat,
                                       const int* __restrict__ scene_offsets, int n_scenes, VoxGeom g,
                                       long long* __restrict__ keys, int* __restrict__ min_idx, unsigned hmask,
                                       int* __restrict__ pt_slot) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int cx, cy, cz;
        const float* p = points + (size_t)i * n_feat;
        float xyz[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
        int slot = -1;
        int n_total = __ldg(scene_offsets + n_scenes);
        if (i < n_total && quantize(xyz, g, cx, cy, cz)) {
            int b = scene_of(i, scene_offsets, n_scenes);
            long long key = (((long long)b * g.grid[2] + cz) * g.grid[1] + cy) * (long long)g.grid[0] + cx;
            unsigned h = hash_key64((unsigned long long)key) & hmask;
            while (true) {
                long long old = (long long)atomicCAS((unsigned long long*)(keys + h), (unsigned long long)kEmptyKey,
                                                     (unsigned long long)key);
                if (old == kEmptyKey || old == key) break;
                h = (h + 1) & hmask;
            }
            atomicMin(min_idx + h, i);
            slot = (int)h;
        }
        pt_slot[i] = slot;
    }
}

__global__ void vox_flag_first_kernel(const int* __restrict__ pt_slot, const int* __restrict__ min_idx, int n,
                                      int* __restrict__ flags) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = pt_slot[i];
        flags[i] = (s >= 0 && min_idx[s] == i) ? 1 : 0;
    }
}

// One thread: per-scene rank starts, kept counts and output bases (n_scenes is small).
__global__ void vox_scene_bases_kernel(const int* __restrict__ rank, const int* __restrict__ total, int n,
                                       const int* __restrict__ scene_offsets, int n_scenes, int max_voxels,
                                       int* __restrict__ scene_rank0, int* __restrict__ scene_base,
                                       int* __restrict__ n_voxels) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int tot = *total;
    int base = 0;
    for (int b = 0; b < n_scenes; ++b) {
        int o0 = scene_offsets[b], o1 = scene_offsets[b + 1];
        int r0 = o0 < n ? rank[o0] : tot;
        int r1 = o1 < n ? rank[o1] : tot;
        int kept = r1 - r0;
        if (kept > max_voxels) kept = max_voxels;
        scene_rank0[b] = r0;
        scene_base[b] = base;
        n_voxels[b] = kept;
        base += kept;
    }
    n_voxels[n_scenes] = base;
}

__global__ void vox_assign_kernel(const float* __restrict__ points, int n, int n_feat,
                                  const int* __restrict__ scene_offsets, int n_scenes, VoxGeom g,
                                  const int* __restrict__ flags, const int* __restrict__ rank,
                                  const int* __restrict__ pt_slot, const int* __restrict__ scene_rank0,
                                  const int* __restrict__ scene_base, int max_voxels, int* __restrict__ slot_vid,
                                  int4* __restrict__ coords) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (!flags[i]) continue;
        int b = scene_of(i, scene_offsets, n_scenes);
        int r = rank[i] - scene_rank0[b];
        int vid = r < max_voxels ? scene_base[b] + r : -1;
        slot_vid[pt_slot[i]] = vid;
        if (vid >= 0) {
            const float* p = points + (size_t)i * n_feat;
            float xyz[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
            int cx, cy, cz;
            quantize(xyz, g, cx, cy, cz);
            coords[vid] = make_int4(b, cz, cy, cx);
        }
    }
}

// Sorted insertion by an atomicMin cascade: slot s ends up with the (s+1)-th smallest index.
__global__ void vox_insert_points_kernel(const int* __restrict__ pt_slot, const int* __restrict__ slot_vid, int n,
                                         int max_points, int* __restrict__ lists) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = pt_slot[i];
        if (s < 0) continue;
        int vid = slot_vid[s];
        if (vid < 0) continue;
        int* l = lists + (size_t)vid * max_points;
        int cur = i;
        for (int k = 0; k < max_points; ++k) {
            int old = atomicMin(l + k, cur);
            if (old == kBigIdx) break;     // took an empty slot
            cur = old > cur ? old : cur;    // the larger one moves on
        }
    }
}

__global__ void vox_gather_kernel(const float* __restrict__ points, int n_feat, const int* __restrict__ lists,
                                  const int* __restrict__ n_voxels_total, int max_points,
                                  float* __restrict__ voxels, int* __restrict__ num_points,
                                  float* __restrict__ voxel_mean) {
    const int nv = *n_voxels_total;
    const int64_t work = (int64_t)nv * max_points;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        int idx = lists[t];
        float* dst = voxels + t * n_feat;
        if (idx != kBigIdx) {
            const float* src = points + (size_t)idx * n_feat;
            for (int c = 0; c < n_feat; ++c) dst[c] = __ldg(src + c);
        } else {
            for (int c = 0; c < n_feat; ++c) dst[c] = 0.f;
        }
    }
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nv; v += gridDim.x * blockDim.x) {
        const int* l = lists + (size_t)v * max_points;
        int cnt = 0;
        for (int k = 0; k < max_points; ++k) cnt += (l[k] != kBigIdx);
        num_points[v] = cnt;
        if (voxel_mean) {
            float denom = (float)(cnt > 1 ? cnt : 1);
            for (int c = 0; c < n_feat; ++c) {
                float s = 0.f;
                for (int k = 0; k < cnt; ++k) s += __ldg(points + (size_t)l[k] * n_feat + c);
                voxel_mean[(size_t)v * n_feat + c] = s / denom;
            }
        }
    }
}

struct VoxWorkspace {
    long long* keys;
    int* min_idx;
    int* slot_vid;
    int* pt_slot;
    int* flags;
    int* rank;
    int* lists;
    int* block_sums;
    int* total;
    int* scene_rank0;
    int* scene_base;
    unsigned hsize;
    int64_t bytes;
};

static VoxWorkspace carve(void* ws, int64_t n_points, int n_scenes, int max_voxels, int max_points) {
    VoxWorkspace w;
    unsigned h = 1024;
    while ((int64_t)h < 2 * n_points) h <<= 1;
    w.hsize = h;
    int64_t off = 0;
    char* base = (char*)ws;
    auto take = [&](int64_t bytes) {
        char* p = base ? base + off : nullptr;
        off += align_up(bytes, 256);
        return (void*)p;
    };
    int64_t npad = n_points > 0 ? n_points : 1;
    w.keys = (long long*)take((int64_t)h * 8);
    w.min_idx = (int*)take((int64_t)h * 4);
    w.lists = (int*)take((int64_t)n_scenes * max_voxels * max_points * 4);
    // --- everything above is initialised with one memset(0xff / 0x7f) each; below needs no init
    w.slot_vid = (int*)take((int64_t)h * 4);
    w.pt_slot = (int*)take(npad * 4);
    w.flags = (int*)take(npad * 4);
    w.rank = (int*)take(npad * 4);
    w.block_sums = (int*)take((int64_t)(scan_num_blocks(npad) + 2) * 4);
    w.total = (int*)take(4);
    w.scene_rank0 = (int*)take((int64_t)n_scenes * 4);
    w.scene_base = (int*)take((int64_t)n_scenes * 4);
    w.bytes = off;
    return w;
}

}  // namespace btc

using namespace btc;

extern "C" {

int64_t btc_voxelize_workspace_bytes(int64_t n_points, int n_scenes, int max_voxels, int max_points) {
    if (n_points < 0 || n_scenes < 1 || max_voxels < 1 || max_points < 1) return BTC_E_BADARG;
    return carve(nullptr, n_points, n_scenes, max_voxels, max_points).bytes;
}

// phases: 1 = grouping (hash insert, first-point ranks, per-scene bases, voxel ids + coordinates + counts[n_scenes + 1]),
//         2 = contents (sorted point lists per voxel, gathered voxels / num_points / MeanVFE features), 3 = both.
// Everything that only needs coordinates (the coordinate hash and the rulebooks of level 1) can start after phase 1.
static int voxelize_phases(int phases, const float* points, int n_points, int n_feat, const int* scene_offsets, int n_scenes,
                           const float* voxel_size, const float* range, const int* grid, int max_points, int max_voxels,
                           float* voxels, int* coords, int* num_points, float* voxel_mean, int* n_voxels, void* workspace,
                           int64_t workspace_bytes, void* stream) {
    if (!scene_offsets || !voxel_size || !range || !grid || !voxels || !coords || !num_points || !n_voxels || !workspace)
        return badarg("btc_voxelize: null argument");
    if (n_points < 0 || n_feat < 3 || n_scenes < 1 || max_points < 1 || max_voxels < 1)
        return badarg("btc_voxelize: bad sizes");
    if (n_points > 0 && !points) return badarg("btc_voxelize: null points");
    if ((int64_t)n_scenes * grid[0] * grid[1] * grid[2] <= 0) return badarg("btc_voxelize: bad grid");
    VoxWorkspace w = carve(workspace, n_points, n_scenes, max_voxels, max_points);
    if (workspace_bytes < w.bytes) return badarg("btc_voxelize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    VoxGeom g;
    for (int j = 0; j < 3; ++j) {
        g.vs[j] = voxel_size[j];
        g.lo[j] = range[j];
        g.grid[j] = grid[j];
    }
    const int T = 256;
    const int nthreads_n = n_points > 0 ? n_points : 1;
    if (phases & 1) {
    BTC_CUDA(cudaMemsetAsync(w.keys, 0xff, (size_t)w.hsize * 8, st), "voxelize memset keys");
    BTC_CUDA(cudaMemsetAsync(w.min_idx, 0x7f, (size_t)w.hsize * 4, st), "voxelize memset min_idx");
    if (n_points > 0) {
        vox_hash_insert_kernel<<<grid_for(n_points, T), T, 0, st>>>(points, n_points, n_feat, scene_offsets, n_scenes, g,
                                                                    w.keys, w.min_idx, w.hsize - 1, w.pt_slot);
        vox_flag_first_kernel<<<grid_for(n_points, T), T, 0, st>>>(w.pt_slot, w.min_idx, n_points, w.flags);
        BTC_CHECK_LAUNCH("voxelize insert/flag");
    } else {
        BTC_CUDA(cudaMemsetAsync(w.flags, 0, 4, st), "voxelize memset flags");
    }
    int rc = launch_flag_scan(w.flags, w.rank, nthreads_n, w.block_sums, w.total, st);
    if (rc) return rc;
    vox_scene_bases_kernel<<<1, 32, 0, st>>>(w.rank, w.total, n_points, scene_offsets, n_scenes, max_voxels,
                                             w.scene_rank0, w.scene_base, n_voxels);
    if (n_points > 0)
        vox_assign_kernel<<<grid_for(n_points, T), T, 0, st>>>(points, n_points, n_feat, scene_offsets, n_scenes, g,
                                                               w.flags, w.rank, w.pt_slot, w.scene_rank0, w.scene_base,
                                                               max_voxels, w.slot_vid, (int4*)coords);
    BTC_CHECK_LAUNCH("voxelize grouping");
    }   // phase 1
    if (!(phases & 2)) return BTC_OK;
    BTC_CUDA(cudaMemsetAsync(w.lists, 0x7f, (size_t)n_scenes * max_voxels * max_points * 4, st), "voxelize memset lists");
    if (n_points > 0)
        vox_insert_points_kernel<<<grid_for(n_points, T), T, 0, st>>>(w.pt_slot, w.slot_vid, n_points, max_points,
                                                                      w.lists);
    int64_t cap_work = (int64_t)n_scenes * max_voxels * max_points;
    int64_t est = (int64_t)n_points * max_points;  // live work is bounded by the number of points
    if (est < cap_work) cap_work = est;
    vox_gather_kernel<<<grid_for(cap_work > 0 ? cap_work : 1, T), T, 0, st>>>(points, n_feat, w.lists, n_voxels + n_scenes,
                                                                              max_points, voxels, num_points, voxel_mean);
    BTC_CHECK_LAUNCH("voxelize assign/gather");
    return BTC_OK;
}

// The grouping phase leaves a complete coordinate -> row hash in the workspace: keys[slot] = ((b * grid_z + z) * grid_y + y)
// * grid_x + x (-1 = empty), vals[slot] = voxel row (-1: a voxel beyond its scene's max_voxels).  Same hash function and
// probing as btc_hash_build, i.e. btc_rulebook_subm_hash can probe it directly with shape (grid_z, grid_y, grid_x) — the
// level-1 sub-manifold rulebook then needs no hash build of its own.  Byte offsets into the workspace + slot count.
int btc_voxelize_hash_view(int64_t n_points, int n_scenes, int max_voxels, int max_points, int64_t* keys_offset,
                           int64_t* vals_offset, int64_t* n_slots) {
    if (n_points < 0 || n_scenes < 1 || max_voxels < 1 || max_points < 1 || !keys_offset || !vals_offset || !n_slots)
        return badarg("btc_voxelize_hash_view: bad arguments");
    char* fake = (char*)(uintptr_t)4096;
    VoxWorkspace w = carve(fake, n_points, n_scenes, max_voxels, max_points);
    *keys_offset = (int64_t)((char*)w.keys - fake);
    *vals_offset = (int64_t)((char*)w.slot_vid - fake);
    *n_slots = (int64_t)w.hsize;
    return BTC_OK;
}

int btc_voxelize(const float* points, int n_points, int n_feat, const int* scene_offsets, int n_scenes,
                 const float* voxel_size, const float* range, const int* grid, int max_points, int max_voxels,
                 float* voxels, int* coords, int* num_points, float* voxel_mean, int* n_voxels, void* workspace,
                 int64_t workspace_bytes, void* stream) {
    return voxelize_phases(3, points, n_points, n_feat, scene_offsets, n_scenes, voxel_size, range, grid, max_points, max_voxels,
                           voxels, coords, num_points, voxel_mean, n_voxels, workspace, workspace_bytes, stream);
}

int btc_voxelize_group(const float* points, int n_points, int n_feat, const int* scene_offsets, int n_scenes,
                       const float* voxel_size, const float* range, const int* grid, int max_points, int max_voxels,
                       float* voxels, int* coords, int* num_points, float* voxel_mean, int* n_voxels, void* workspace,
                       int64_t workspace_bytes, void* stream) {
    return voxelize_phases(1, points, n_points, n_feat, scene_offsets, n_scenes, voxel_size, range, grid, max_points, max_voxels,
                           voxels, coords, num_points, voxel_mean, n_voxels, workspace, workspace_bytes, stream);
}

int btc_voxelize_fill(const float* points, int n_points, int n_feat, const int* scene_offsets, int n_scenes,
                      const float* voxel_size, const float* range, const int* grid, int max_points, int max_voxels,
                      float* voxels, int* coords, int* num_points, float* voxel_mean, int* n_voxels, void* workspace,
                      int64_t workspace_bytes, void* stream) {
    return voxelize_phases(2, points, n_points, n_feat, scene_offsets, n_scenes, voxel_size, range, grid, max_points, max_voxels,
                           voxels, coords, num_points, voxel_mean, n_voxels, workspace, workspace_bytes, stream);
}

}  // extern "C"
