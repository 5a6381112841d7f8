// TEST INFRASTRUCTURE ONLY (see cuda_emul.h): the RoI-pooling kernels of btcdet_b200/csrc/roi_pool_kernels.cuh executed
// on the host under the lock-step warp emulation, behind entry points shaped like the C ABI's
// (include/btcdet_b200.h: btc_ball_query_stack, btc_group_points_stack(_grad), btc_trilinear_sparse_flag/_emit).
// Built by tests/test_roi_pool_cpu.py with g++; the buffers are host memory.
#include "cuda_emul.h"
#include "../../btcdet_b200/csrc/roi_pool_kernels.cuh"

namespace btc {
thread_local char g_last_error[256];
int set_error(int code, const char*, cudaError_t) { return code; }
}  // namespace btc

using namespace btc;
using namespace btc::roi;

extern "C" {

int emul_ball_query_stack(int B, int M, int n_radii, const float* radii, const int* nsamples, const float* new_xyz,
                          const int* new_cnt, const float* xyz, const int* cnt, int* const* idx, int blocks) {
    BallArgs a;
    a.n_radii = n_radii;
    for (int r = 0; r < kBqMaxRadii; ++r) {
        a.r2[r] = 0.f; a.nsample[r] = 0; a.idx[r] = nullptr;
        if (r < n_radii) { const float rad = radii[r]; a.r2[r] = rad * rad; a.nsample[r] = nsamples[r]; a.idx[r] = idx[r]; }
    }
    emul::launch(blocks, [&] { ball_query_kernel(B, M, a, new_xyz, new_cnt, xyz, cnt); });
    return 0;
}

int emul_group_points_stack(int B, int M, int C, int ns, const float* feat, const int* f_cnt, const int* idx,
                            const int* i_cnt, float* out, int blocks) {
    emul::launch(blocks, [&] { group_points_kernel(B, (int64_t)M * C * ns, C, ns, feat, f_cnt, idx, i_cnt, out); });
    return 0;
}

int emul_group_points_stack_grad(int B, int M, int C, int ns, const float* grad_out, const int* idx, const int* i_cnt,
                                 const int* f_cnt, float* grad_feat, int blocks) {
    emul::launch(blocks, [&] { group_points_grad_kernel(B, (int64_t)M * C * ns, C, ns, grad_out, idx, i_cnt, f_cnt, grad_feat); });
    return 0;
}

// flag + host scan + emit; returns the number of non-zero rows (rows beyond out_cap are dropped)
int emul_trilinear_sparse(const float* feats, const int* coords, int n, int C, int batch, const int* shape, const float* zyx,
                          const long long* b_target, long long T, long long per_scene, int normalize, int P,
                          const int* local_shape, int out_cap, float* out_feats, int* out_coords, long long* out_target,
                          int blocks) {
    TriGeom g;
    g.B = batch; g.Z = shape[0]; g.Y = shape[1]; g.X = shape[2]; g.C = C; g.normalize = normalize; g.T = T;
    g.per_scene = per_scene > 0 ? per_scene : 1;
    std::vector<int> vol((size_t)batch * shape[0] * shape[1] * shape[2], -1), flags((size_t)T, -7), rank((size_t)T, 0);
    emul::launch(blocks, [&] { index_volume_kernel((const int4*)coords, n, nullptr, g, vol.data()); });
    emul::launch(blocks, [&] {
        tri_kernel<false>(feats, zyx, b_target, g, vol.data(), flags.data(), nullptr, 1, 1, 1, 0, nullptr, nullptr, nullptr);
    });
    int total = 0;
    for (long long t = 0; t < T; ++t) {
        if (flags[t] != 0 && flags[t] != 1) return -1000;   // a target the flag pass never wrote
        rank[t] = total;
        total += flags[t];
    }
    emul::launch(blocks, [&] {
        tri_kernel<true>(feats, zyx, b_target, g, vol.data(), flags.data(), rank.data(), P, local_shape[1], local_shape[2],
                         out_cap, out_feats, (int4*)out_coords, out_target);
    });
    return total;
}

// adjoint of the emitted rows: grad_feats [n, C] (zeroed by the caller) += w_j * grad_out[o]
int emul_trilinear_sparse_grad(const float* grad_out, const long long* out_target, int n_out, int C, const float* /*feats*/,
                               const int* coords, int n, int batch, int /*unused*/, const int* shape, const float* zyx,
                               const long long* b_target, long long T, long long per_scene, int normalize, float* grad_feats,
                               int blocks) {
    TriGeom g;
    g.B = batch; g.Z = shape[0]; g.Y = shape[1]; g.X = shape[2]; g.C = C; g.normalize = normalize; g.T = T;
    g.per_scene = per_scene > 0 ? per_scene : 1;
    std::vector<int> vol((size_t)batch * shape[0] * shape[1] * shape[2], -1);
    emul::launch(blocks, [&] { index_volume_kernel((const int4*)coords, n, nullptr, g, vol.data()); });
    emul::launch(blocks, [&] { tri_grad_kernel(grad_out, out_target, n_out, nullptr, zyx, b_target, g, vol.data(), grad_feats); });
    return 0;
}

}  // extern "C"
