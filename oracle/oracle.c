/*
 * oracle.c — CPU restatement of the reference algorithms of the BtcDet hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under btcdet_b200/ or spconv/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, as the checker or the timed CPU baseline.
 *
 * PARITY UNPINNED at the spconv boundary: the algorithms restated here live in the third-party
 * dependency spconv v1.2.1 (traveller59/spconv, commit fad3000, pinned by the reference's
 * README.md:42 / setup.py:42), which is neither vendored under /root/reference nor installable
 * offline, and the reference ships no tests or golden vectors.  The
 * restatement follows spconv 1.2.1's published behaviour as recorded in SURVEY.md App. A and is
 * cross-checked in tests/ against an independent dense formulation with stock torch ops
 * (F.conv3d / conv_transpose3d / max_pool3d on zero-filled volumes) and hand-derived
 * known-answer tests.
 *
 * Functions and the reference code they follow:
 *   orc_points_to_voxel   spconv 1.2.1 include/spconv/point2voxel.h points_to_voxel_3d_np, called
 *                         through spconv.utils.VoxelGeneratorV2.generate
 *                         (btcdet/datasets/processor/data_processor.py:85,136,177)   [App. A.1]
 *   orc_rulebook          spconv 1.2.1 src/spconv/indice.cu + spconv/ops.py get_indice_pairs as
 *                         called by SparseConvolution.forward / SparseMaxPool.forward
 *                         (btcdet/models/backbones_3d/spconv_backbone.py:12-29)      [App. A.2-A.4]
 *   orc_indice_conv       spconv 1.2.1 src/spconv/spconv_ops.cc indiceConv (Native)   [App. A.5]
 *   orc_indice_maxpool    spconv 1.2.1 src/spconv/maxpool.cu                          [App. A.6]
 *   orc_ball_query_stack / orc_group_points_stack(_grad)
 *                         the reference's OWN in-tree kernels (not spconv: parity PINNED, see below):
 *                         btcdet/ops/pointnet2/pointnet2_stack/src/ball_query_gpu.cu:16-60,
 *                         group_points_gpu.cu:16-95, restated thread for thread.  d2 follows the SASS nvcc
 *                         emits for the reference's source line :42 (FMUL, FFMA, FFMA: nvcc's default
 *                         -fmad=true contracts it), written with fmaf().  Pinned on the GPU box against the
 *                         reference kernels themselves, compiled from the checkout into
 *                         oracle/_ref/libpointnet2_ref.so (tests/test_roi_pool_gpu.py).          [SURVEY 8f N1]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* points_to_voxel_3d_np — sequential, first come first served.                                 */
/* points [n, c] f32; voxel_size[3], range[6] in (x,y,z) order; grid[3] (x,y,z).                */
/* lookup: caller-provided int32 volume of grid z*y*x cells, all -1 on entry, restored on exit. */
/* outputs pre-sized for max_voxels; returns voxel_num.                                         */
/* ------------------------------------------------------------------------------------------ */
int orc_points_to_voxel(const float* points, int n, int c, const float* voxel_size, const float* range,
                        const int* grid, int max_points, int max_voxels, float* voxels /*[max_voxels,max_points,c]*/,
                        int* coors /*[max_voxels,3] zyx*/, int* num_points_per_voxel /*[max_voxels]*/,
                        int* lookup /*[gz*gy*gx]*/) {
    int voxel_num = 0;
    memset(voxels, 0, sizeof(float) * (size_t)max_voxels * max_points * c);
    memset(num_points_per_voxel, 0, sizeof(int) * (size_t)max_voxels);
    for (int i = 0; i < n; ++i) {
        int coor[3]; /* z, y, x */
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            /* fp32 arithmetic, as the float template instantiation of the reference */
            float q = (points[(size_t)i * c + j] - range[j]) / voxel_size[j];
            int cj = (int)floorf(q);
            if (!(floorf(q) >= 0.f) || cj < 0 || cj >= grid[j]) {
                failed = 1;
                break;
            }
            coor[2 - j] = cj;
        }
        if (failed) continue;
        size_t cell = ((size_t)coor[0] * grid[1] + coor[1]) * grid[0] + coor[2];
        int voxelidx = lookup[cell];
        if (voxelidx == -1) {
            voxelidx = voxel_num;
            if (voxel_num >= max_voxels) continue;
            voxel_num += 1;
            lookup[cell] = voxelidx;
            for (int k = 0; k < 3; ++k) coors[voxelidx * 3 + k] = coor[k];
        }
        int num = num_points_per_voxel[voxelidx];
        if (num < max_points) {
            memcpy(voxels + ((size_t)voxelidx * max_points + num) * c, points + (size_t)i * c, sizeof(float) * c);
            num_points_per_voxel[voxelidx] += 1;
        }
    }
    for (int v = 0; v < voxel_num; ++v) {
        size_t cell = ((size_t)coors[v * 3] * grid[1] + coors[v * 3 + 1]) * grid[0] + coors[v * 3 + 2];
        lookup[cell] = -1;
    }
    return voxel_num;
}

/* ------------------------------------------------------------------------------------------ */
/* tiny int64 -> int32 open-addressing map (coordinate -> row)                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int64_t* keys;
    int* vals;
    uint64_t mask;
} orc_map;

static uint64_t mix64(uint64_t k) {
    k ^= k >> 31;
    k *= 0x9e3779b97f4a7c15ULL;
    k ^= k >> 29;
    return k;
}

static int map_init(orc_map* m, int64_t n) {
    uint64_t cap = 16;
    while (cap < (uint64_t)(2 * n + 1)) cap <<= 1;
    m->keys = (int64_t*)malloc(sizeof(int64_t) * cap);
    m->vals = (int*)malloc(sizeof(int) * cap);
    if (!m->keys || !m->vals) return -1;
    for (uint64_t i = 0; i < cap; ++i) m->keys[i] = -1;
    m->mask = cap - 1;
    return 0;
}
static void map_free(orc_map* m) {
    free(m->keys);
    free(m->vals);
}
static void map_put(orc_map* m, int64_t key, int val) {
    uint64_t h = mix64((uint64_t)key) & m->mask;
    while (m->keys[h] != -1 && m->keys[h] != key) h = (h + 1) & m->mask;
    m->keys[h] = key;
    m->vals[h] = val;
}
static int map_get(const orc_map* m, int64_t key) {
    uint64_t h = mix64((uint64_t)key) & m->mask;
    while (m->keys[h] != -1) {
        if (m->keys[h] == key) return m->vals[h];
        h = (h + 1) & m->mask;
    }
    return -1;
}

static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------ */
/* Rulebook (get_indice_pairs).                                                                 */
/* indices [n,4] (b,z,y,x).  Kernel offsets row-major over (kz,ky,kx).                          */
/*   subm:       out sites = in sites (same rows); pair (in=j,out=i) at offset k iff             */
/*               coord[j] == coord[i] + (k - k/2)*dil.                                           */
/*   regular:    pair iff out*stride - pad + k*dil == in.                                        */
/*   transposed: pair iff out == in*stride - pad + k*dil.                                        */
/* Output sites of non-subm rulebooks are the unique flat keys b*prod(out)+rowmajor(z,y,x) in    */
/* ascending order (GPU path of spconv: torch::_unique).  Pairs are emitted in the canonical     */
/* order (offset, input row).  pairs [2,K,n] padded with -1, pair_num [K].                       */
/* out_indices must have room for out_cap rows; returns the number of output sites, or           */
/* -(needed) if out_cap is too small, or INT32_MIN on allocation failure.                        */
/* ------------------------------------------------------------------------------------------ */
int orc_rulebook(const int* indices, int n, int batch, const int* in_shape, const int* out_shape, const int* ksize,
                 const int* stride, const int* padding, const int* dilation, int subm, int transposed,
                 int* out_indices, int out_cap, int* pairs, int* pair_num) {
    const int K = ksize[0] * ksize[1] * ksize[2];
    (void)batch;
    for (int64_t t = 0; t < (int64_t)2 * K * n; ++t) pairs[t] = -1;
    for (int k = 0; k < K; ++k) pair_num[k] = 0;
    if (n == 0) return 0;

    if (subm) {
        orc_map m;
        if (map_init(&m, n)) return INT32_MIN;
        for (int i = 0; i < n; ++i) {
            const int* c = indices + (size_t)i * 4;
            int64_t key = (((int64_t)c[0] * in_shape[0] + c[1]) * in_shape[1] + c[2]) * in_shape[2] + c[3];
            map_put(&m, key, i);
        }
        for (int k = 0; k < K; ++k) {
            int kx = k % ksize[2], ky = (k / ksize[2]) % ksize[1], kz = k / (ksize[2] * ksize[1]);
            int cnt = 0;
            /* canonical order: ascending input row j; out row i sits at coord[j] - offset */
            for (int j = 0; j < n; ++j) {
                const int* c = indices + (size_t)j * 4;
                int z = c[1] - (kz - ksize[0] / 2) * dilation[0];
                int y = c[2] - (ky - ksize[1] / 2) * dilation[1];
                int x = c[3] - (kx - ksize[2] / 2) * dilation[2];
                if (z < 0 || z >= in_shape[0] || y < 0 || y >= in_shape[1] || x < 0 || x >= in_shape[2]) continue;
                int64_t key = (((int64_t)c[0] * in_shape[0] + z) * in_shape[1] + y) * in_shape[2] + x;
                int i = map_get(&m, key);
                if (i < 0) continue;
                pairs[(size_t)k * n + cnt] = j;
                pairs[((size_t)K + k) * n + cnt] = i;
                ++cnt;
            }
            pair_num[k] = cnt;
        }
        map_free(&m);
        if (out_cap < n) return -n;
        memcpy(out_indices, indices, sizeof(int) * (size_t)n * 4);
        return n;
    }

    /* pass 1: candidate output keys */
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * K);
    if (!cand) return INT32_MIN;
    int64_t ncand = 0;
    for (int i = 0; i < n; ++i) {
        const int* c = indices + (size_t)i * 4;
        for (int k = 0; k < K; ++k) {
            int kk[3] = {k / (ksize[2] * ksize[1]), (k / ksize[2]) % ksize[1], k % ksize[2]};
            int o[3], ok = 1;
            for (int a = 0; a < 3 && ok; ++a) {
                int in = c[1 + a];
                if (transposed) {
                    o[a] = in * stride[a] - padding[a] + kk[a] * dilation[a];
                } else {
                    int t = in + padding[a] - kk[a] * dilation[a];
                    if (t < 0 || t % stride[a] != 0) ok = 0;
                    o[a] = t / stride[a];
                }
                if (ok && (o[a] < 0 || o[a] >= out_shape[a])) ok = 0;
            }
            if (!ok) continue;
            cand[ncand++] = (((int64_t)c[0] * out_shape[0] + o[0]) * out_shape[1] + o[1]) * out_shape[2] + o[2];
        }
    }
    /* sorted unique == torch::_unique(flat) */
    qsort(cand, (size_t)ncand, sizeof(int64_t), cmp_i64);
    int64_t nout = 0;
    for (int64_t t = 0; t < ncand; ++t)
        if (t == 0 || cand[t] != cand[t - 1]) cand[nout++] = cand[t];
    if (nout > out_cap) {
        free(cand);
        return -(int)nout;
    }
    orc_map m;
    if (map_init(&m, nout)) {
        free(cand);
        return INT32_MIN;
    }
    for (int64_t r = 0; r < nout; ++r) {
        int64_t key = cand[r];
        map_put(&m, key, (int)r);
        int* oi = out_indices + (size_t)r * 4;
        oi[3] = (int)(key % out_shape[2]);
        key /= out_shape[2];
        oi[2] = (int)(key % out_shape[1]);
        key /= out_shape[1];
        oi[1] = (int)(key % out_shape[0]);
        oi[0] = (int)(key / out_shape[0]);
    }
    /* pass 2: pairs in (offset, input row) order */
    for (int k = 0; k < K; ++k) {
        int kk[3] = {k / (ksize[2] * ksize[1]), (k / ksize[2]) % ksize[1], k % ksize[2]};
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            const int* c = indices + (size_t)i * 4;
            int o[3], ok = 1;
            for (int a = 0; a < 3 && ok; ++a) {
                int in = c[1 + a];
                if (transposed) {
                    o[a] = in * stride[a] - padding[a] + kk[a] * dilation[a];
                } else {
                    int t = in + padding[a] - kk[a] * dilation[a];
                    if (t < 0 || t % stride[a] != 0) ok = 0;
                    o[a] = t / stride[a];
                }
                if (ok && (o[a] < 0 || o[a] >= out_shape[a])) ok = 0;
            }
            if (!ok) continue;
            int64_t key = (((int64_t)c[0] * out_shape[0] + o[0]) * out_shape[1] + o[1]) * out_shape[2] + o[2];
            pairs[(size_t)k * n + cnt] = i;
            pairs[((size_t)K + k) * n + cnt] = map_get(&m, key);
            ++cnt;
        }
        pair_num[k] = cnt;
    }
    map_free(&m);
    free(cand);
    return (int)nout;
}

/* ------------------------------------------------------------------------------------------ */
/* indice_conv, Native algorithm: out = 0; (subm: out = feat @ W[center] first, centre skipped);  */
/* for each offset in ascending order with pairs: gather -> matmul -> scatter-add.  fp32.         */
/* features [n_in, cin], weight [K, cin, cout], out [n_out, cout].                               */
/* ------------------------------------------------------------------------------------------ */
void orc_indice_conv(const float* features, int n_in, const float* weight, int K, int cin, int cout,
                     const int* pairs /*[2,K,n_in]*/, const int* pair_num, int n_out, int subm, float* out) {
    memset(out, 0, sizeof(float) * (size_t)n_out * cout);
    int center = -1;
    if (subm) {
        center = K / 2;
        const float* w = weight + (size_t)center * cin * cout;
        for (int r = 0; r < n_in && r < n_out; ++r)
            for (int ci = 0; ci < cin; ++ci) {
                float a = features[(size_t)r * cin + ci];
                const float* wr = w + (size_t)ci * cout;
                float* o = out + (size_t)r * cout;
                for (int co = 0; co < cout; ++co) o[co] += a * wr[co];
            }
    }
    float* buf = (float*)malloc(sizeof(float) * (size_t)cout);
    for (int k = 0; k < K; ++k) {
        if (k == center) continue;
        const float* w = weight + (size_t)k * cin * cout;
        for (int s = 0; s < pair_num[k]; ++s) {
            int i = pairs[(size_t)k * n_in + s], o = pairs[((size_t)K + k) * n_in + s];
            for (int co = 0; co < cout; ++co) buf[co] = 0.f;
            for (int ci = 0; ci < cin; ++ci) {
                float a = features[(size_t)i * cin + ci];
                const float* wr = w + (size_t)ci * cout;
                for (int co = 0; co < cout; ++co) buf[co] += a * wr[co];
            }
            float* orow = out + (size_t)o * cout;
            for (int co = 0; co < cout; ++co) orow[co] += buf[co];
        }
    }
    free(buf);
}

/* indice_maxpool: out zero-initialised, out[o] = max(out[o], in[i]) over all pairs. */
void orc_indice_maxpool(const float* features, int n_in, int c, int K, const int* pairs, const int* pair_num,
                        int n_out, float* out) {
    memset(out, 0, sizeof(float) * (size_t)n_out * c);
    for (int k = 0; k < K; ++k)
        for (int s = 0; s < pair_num[k]; ++s) {
            int i = pairs[(size_t)k * n_in + s], o = pairs[((size_t)K + k) * n_in + s];
            for (int ch = 0; ch < c; ++ch) {
                float v = features[(size_t)i * c + ch];
                if (v > out[(size_t)o * c + ch]) out[(size_t)o * c + ch] = v;
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* Stacked ball query / grouping of the RoI head.                                */
/* One loop iteration = one thread of the reference kernel.  idx must be zero-initialised by    */
/* the caller exactly like the reference's Python wrapper does (pointnet2_utils.py:33).          */
/* ------------------------------------------------------------------------------------------ */
static void orc_scene_of(int pt, int B, const int* q_cnt, const int* p_cnt, int* scene, int* start) {
    int bs = 0, acc = q_cnt[0], k, s = 0;
    for (k = 1; k < B; k++) {
        if (pt < acc) break;
        acc += q_cnt[k];
        bs = k;
    }
    for (k = 0; k < bs; k++) s += p_cnt[k];
    *scene = bs;
    *start = s;
}

void orc_ball_query_stack(int B, int M, float radius, int nsample, const float* new_xyz, const int* new_xyz_batch_cnt,
                          const float* xyz, const int* xyz_batch_cnt, int* idx) {
    int pt;
    const float radius2 = radius * radius;
    for (pt = 0; pt < M; pt++) {
        int scene, start, k, l, cnt = 0, n;
        const float* q = new_xyz + (size_t)pt * 3;
        const float* p;
        int* row = idx + (size_t)pt * nsample;
        orc_scene_of(pt, B, new_xyz_batch_cnt, xyz_batch_cnt, &scene, &start);
        p = xyz + (size_t)start * 3;
        n = xyz_batch_cnt[scene];
        for (k = 0; k < n; ++k) {
            const float dx = q[0] - p[k * 3 + 0], dy = q[1] - p[k * 3 + 1], dz = q[2] - p[k * 3 + 2];
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (d2 < radius2) {
                if (cnt == 0)
                    for (l = 0; l < nsample; ++l) row[l] = k;
                row[cnt] = k;
                ++cnt;
                if (cnt >= nsample) break;
            }
        }
        if (cnt == 0) row[0] = -1;
    }
}

void orc_group_points_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                            const int* idx, const int* idx_batch_cnt, float* out) {
    int pt, c, s;
    for (pt = 0; pt < M; pt++) {
        int scene, start;
        orc_scene_of(pt, B, idx_batch_cnt, features_batch_cnt, &scene, &start);
        for (c = 0; c < C; c++)
            for (s = 0; s < nsample; s++)
                out[((size_t)pt * C + c) * nsample + s] = features[((size_t)start + idx[(size_t)pt * nsample + s]) * C + c];
    }
}

/* grad_features [N, C] is accumulated in double and rounded once (the reference's float atomics are order dependent) */
void orc_group_points_grad_stack(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                 const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features) {
    int pt, c, s;
    size_t i;
    double* acc = (double*)calloc((size_t)N * C + 1, sizeof(double));
    for (pt = 0; pt < M; pt++) {
        int scene, start;
        orc_scene_of(pt, B, idx_batch_cnt, features_batch_cnt, &scene, &start);
        for (c = 0; c < C; c++)
            for (s = 0; s < nsample; s++)
                acc[((size_t)start + idx[(size_t)pt * nsample + s]) * C + c] += grad_out[((size_t)pt * C + c) * nsample + s];
    }
    for (i = 0; i < (size_t)N * C; i++) grad_features[i] = (float)acc[i];
    free(acc);
}
