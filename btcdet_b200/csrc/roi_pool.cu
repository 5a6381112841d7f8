// RoI grid pooling of the second stage: C entry points; the kernels are in roi_pool_kernels.cuh.
#include "roi_pool_kernels.cuh"

namespace btc {
namespace {
using namespace roi;

struct TriWs {
    int* vol;
    int* flags;
    int* rank;
    int* block_sums;
    int* total;
    int64_t bytes;
};

// Layout of the workspace; with workspace == nullptr only `bytes` is meaningful (size query).
TriWs tri_ws(void* workspace, int64_t cells, int64_t T) {
    const int64_t o_flags = align_up(cells * 4, 256);
    const int64_t o_rank = o_flags + align_up(T * 4, 256);
    const int64_t o_sums = o_rank + align_up(T * 4, 256);
    const int64_t o_total = o_sums + align_up((int64_t)(scan_num_blocks(T) + 2) * 4, 256);
    TriWs w = {nullptr, nullptr, nullptr, nullptr, nullptr, o_total + 256};
    if (workspace) {
        char* p = (char*)workspace;
        w.vol = (int*)p;
        w.flags = (int*)(p + o_flags);
        w.rank = (int*)(p + o_rank);
        w.block_sums = (int*)(p + o_sums);
        w.total = (int*)(p + o_total);
    }
    return w;
}

int tri_geom(TriGeom& g, int C, int batch, const int* shape, int64_t n_targets, int64_t per_scene, int normalize,
             const long long* b_target) {
    if (!shape || batch < 1 || C < 1 || shape[0] < 1 || shape[1] < 1 || shape[2] < 1) return badarg("btc_trilinear_sparse: bad geometry");
    if (n_targets < 0 || n_targets > 0x7fffffff) return badarg("btc_trilinear_sparse: bad number of targets");
    if (!b_target && per_scene < 1) return badarg("btc_trilinear_sparse: per_scene must be >= 1 without b_target");
    if ((int64_t)batch * shape[0] * shape[1] * shape[2] > 0x7fffffff) return badarg("btc_trilinear_sparse: grid too large");
    g.B = batch; g.Z = shape[0]; g.Y = shape[1]; g.X = shape[2]; g.C = C; g.normalize = normalize ? 1 : 0;
    g.T = n_targets; g.per_scene = per_scene > 0 ? per_scene : 1;
    return BTC_OK;
}

}  // namespace
}  // namespace btc

using namespace btc;

extern "C" {

int btc_ball_query_stack(int B, int M, int n_radii, const float* radii, const int* nsamples, const float* new_xyz,
                         const int* new_xyz_batch_cnt, const float* xyz, const int* xyz_batch_cnt, int* const* idx,
                         void* stream) {
    if (B < 1 || M < 0 || n_radii < 1 || n_radii > kBqMaxRadii) return badarg("btc_ball_query_stack: bad sizes (1..4 radii)");
    if (M == 0) return BTC_OK;
    if (!radii || !nsamples || !new_xyz || !new_xyz_batch_cnt || !xyz_batch_cnt || !idx) return badarg("btc_ball_query_stack: null argument");
    BallArgs a;
    a.n_radii = n_radii;
    for (int r = 0; r < kBqMaxRadii; ++r) {
        a.r2[r] = 0.f; a.nsample[r] = 0; a.idx[r] = nullptr;
        if (r < n_radii) {
            if (nsamples[r] < 1 || !idx[r]) return badarg("btc_ball_query_stack: nsample < 1 or null idx");
            const float rad = radii[r];
            a.r2[r] = rad * rad;   // fp32 product, as in the reference kernel (ball_query_gpu.cu:36)
            a.nsample[r] = nsamples[r];
            a.idx[r] = idx[r];
        }
    }
    const int64_t groups = ((int64_t)M + kBqQueries - 1) / kBqQueries;
    ball_query_kernel<<<grid_for(groups * 32, 256, 4), 256, 0, (cudaStream_t)stream>>>(B, M, a, new_xyz, new_xyz_batch_cnt, xyz,
                                                                                      xyz_batch_cnt);
    BTC_CHECK_LAUNCH("ball_query");
    return BTC_OK;
}

int btc_group_points_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                           const int* idx, const int* idx_batch_cnt, float* out, void* stream) {
    if (B < 1 || M < 0 || C < 1 || nsample < 1) return badarg("btc_group_points_stack: bad sizes");
    if (M == 0) return BTC_OK;
    if (!features_batch_cnt || !idx || !idx_batch_cnt || !out) return badarg("btc_group_points_stack: null argument");
    const int64_t total = (int64_t)M * C * nsample;
    group_points_kernel<<<grid_for(total, 256, 4), 256, 0, (cudaStream_t)stream>>>(B, total, C, nsample, features, features_batch_cnt,
                                                                                  idx, idx_batch_cnt, out);
    BTC_CHECK_LAUNCH("group_points");
    return BTC_OK;
}

int btc_group_points_stack_grad(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features,
                                void* stream) {
    if (B < 1 || M < 0 || C < 1 || N < 0 || nsample < 1) return badarg("btc_group_points_stack_grad: bad sizes");
    if (M == 0 || N == 0) return BTC_OK;
    if (!grad_out || !idx || !idx_batch_cnt || !features_batch_cnt || !grad_features) return badarg("btc_group_points_stack_grad: null argument");
    const int64_t total = (int64_t)M * C * nsample;
    group_points_grad_kernel<<<grid_for(total, 256, 4), 256, 0, (cudaStream_t)stream>>>(B, total, C, nsample, grad_out, idx,
                                                                                       idx_batch_cnt, features_batch_cnt,
                                                                                       grad_features);
    BTC_CHECK_LAUNCH("group_points_grad");
    return BTC_OK;
}

int64_t btc_trilinear_sparse_workspace_bytes(int64_t n_targets, int batch, const int* shape) {
    if (!shape || batch < 1 || n_targets < 0 || n_targets > 0x7fffffff) return BTC_E_BADARG;
    const int64_t cells = (int64_t)batch * shape[0] * shape[1] * shape[2];
    if (cells < 1 || cells > 0x7fffffff) return BTC_E_BADARG;
    return tri_ws(nullptr, cells, n_targets > 0 ? n_targets : 1).bytes;
}

int btc_trilinear_sparse_flag(const float* feats, const int* coords, int n_cap, const int* n_dev, int C, int batch,
                              const int* shape, const float* zyx, const long long* b_target, int64_t n_targets,
                              int64_t per_scene, int normalize, int* count, void* workspace, int64_t workspace_bytes,
                              void* stream) {
    TriGeom g;
    int rc = tri_geom(g, C, batch, shape, n_targets, per_scene, normalize, b_target);
    if (rc) return rc;
    if (!count || !workspace) return badarg("btc_trilinear_sparse_flag: null argument");
    if (n_cap < 0 || (n_cap > 0 && (!feats || !coords))) return badarg("btc_trilinear_sparse_flag: null inputs");
    if (n_targets > 0 && !zyx) return badarg("btc_trilinear_sparse_flag: null targets");
    const int64_t cells = (int64_t)g.B * g.Z * g.Y * g.X;
    const int64_t T = n_targets > 0 ? n_targets : 1;
    if (workspace_bytes < tri_ws(nullptr, cells, T).bytes) return badarg("btc_trilinear_sparse_flag: workspace too small");
    TriWs w = tri_ws(workspace, cells, T);
    cudaStream_t st = (cudaStream_t)stream;
    BTC_CUDA(cudaMemsetAsync(w.vol, 0xff, cells * 4, st), "trilinear: clear index volume");
    if (n_targets == 0) {
        BTC_CUDA(cudaMemsetAsync(count, 0, 4, st), "trilinear: clear count");
        return BTC_OK;
    }
    if (n_cap > 0) index_volume_kernel<<<grid_for(n_cap, 256), 256, 0, st>>>((const int4*)coords, n_cap, n_dev, g, w.vol);
    tri_kernel<false><<<grid_for(n_targets * 32, 256, 4), 256, 0, st>>>(feats, zyx, b_target, g, w.vol, w.flags, nullptr, 1, 1, 1, 0,
                                                                       nullptr, nullptr, nullptr);
    BTC_CHECK_LAUNCH("trilinear flag");
    rc = launch_flag_scan(w.flags, w.rank, (int)n_targets, w.block_sums, w.total, st);
    if (rc) return rc;
    BTC_CUDA(cudaMemcpyAsync(count, w.total, 4, cudaMemcpyDeviceToDevice, st), "trilinear: count");
    return BTC_OK;
}

int btc_trilinear_sparse_emit(const float* feats, int C, int batch, const int* shape, const float* zyx,
                              const long long* b_target, int64_t n_targets, int64_t per_scene, int normalize, int P,
                              const int* local_shape, int out_cap, float* out_feats, int* out_coords,
                              long long* out_target, void* workspace, int64_t workspace_bytes, void* stream) {
    TriGeom g;
    int rc = tri_geom(g, C, batch, shape, n_targets, per_scene, normalize, b_target);
    if (rc) return rc;
    if (!local_shape || P < 1 || (int64_t)local_shape[0] * local_shape[1] * local_shape[2] != P)
        return badarg("btc_trilinear_sparse_emit: P must be the number of cells of local_shape");
    if (n_targets == 0 || out_cap <= 0) return BTC_OK;
    if (!feats || !zyx || !out_feats || !out_coords || !workspace) return badarg("btc_trilinear_sparse_emit: null argument");
    const int64_t cells = (int64_t)g.B * g.Z * g.Y * g.X;
    if (workspace_bytes < tri_ws(nullptr, cells, n_targets).bytes) return badarg("btc_trilinear_sparse_emit: workspace too small");
    TriWs w = tri_ws(workspace, cells, n_targets);
    tri_kernel<true><<<grid_for(n_targets * 32, 256, 4), 256, 0, (cudaStream_t)stream>>>(feats, zyx, b_target, g, w.vol, w.flags, w.rank,
                                                                                        P, local_shape[1], local_shape[2], out_cap,
                                                                                        out_feats, (int4*)out_coords, out_target);
    BTC_CHECK_LAUNCH("trilinear emit");
    return BTC_OK;
}

int btc_trilinear_sparse_grad(const float* grad_out, const long long* out_target, int n_out, const int* n_out_dev, int C,
                              int batch, const int* shape, const float* zyx, const long long* b_target, int64_t n_targets,
                              int64_t per_scene, int normalize, float* grad_feats, void* workspace, int64_t workspace_bytes,
                              void* stream) {
    TriGeom g;
    int rc = tri_geom(g, C, batch, shape, n_targets, per_scene, normalize, b_target);
    if (rc) return rc;
    if (n_out <= 0 || n_targets == 0) return BTC_OK;
    if (!grad_out || !out_target || !zyx || !grad_feats || !workspace) return badarg("btc_trilinear_sparse_grad: null argument");
    const int64_t cells = (int64_t)g.B * g.Z * g.Y * g.X;
    if (workspace_bytes < tri_ws(nullptr, cells, n_targets).bytes) return badarg("btc_trilinear_sparse_grad: workspace too small");
    TriWs w = tri_ws(workspace, cells, n_targets);
    tri_grad_kernel<<<grid_for((int64_t)n_out * 32, 256, 4), 256, 0, (cudaStream_t)stream>>>(grad_out, out_target, n_out, n_out_dev, zyx,
                                                                                           b_target, g, w.vol, grad_feats);
    BTC_CHECK_LAUNCH("trilinear grad");
    return BTC_OK;
}

}  // extern "C"
