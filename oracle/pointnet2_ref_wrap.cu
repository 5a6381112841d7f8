// TEST INFRASTRUCTURE ONLY — never linked into or imported by the product (btcdet_b200/, spconv/).
//
// Compiles the REFERENCE'S OWN stacked ball-query / grouping CUDA kernels
// (btcdet/ops/pointnet2/pointnet2_stack/src/ball_query_gpu.cu, group_points_gpu.cu) from where they lie under the
// reference checkout — the sources are #included by path (given by oracle/Makefile), nothing of them is copied into this
// repository — with nvcc's default flags (as the reference's setup.py builds them: -fmad=true) for sm_100a, and exposes
// their launchers through a plain C ABI, so that the GPU tests can use the reference kernels themselves as the oracle
// for SURVEY §8(f) N1 (`oracle/_ref/libpointnet2_ref.so`, git-ignored, travels to the GPU box).  The reference's torch
// wrappers (ball_query.cpp, group_points.cpp) need THC headers torch 2.11 no longer ships; the kernels do not.
#include BTC_REF_BALL_QUERY
#include BTC_REF_GROUP_POINTS

extern "C" {

void ref_ball_query_stack(int B, int M, float radius, int nsample, const float* new_xyz, const int* new_xyz_batch_cnt,
                          const float* xyz, const int* xyz_batch_cnt, int* idx) {
    ball_query_kernel_launcher_stack(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx);
}

void ref_group_points_stack(int B, int M, int C, int nsample, const float* features, const int* features_batch_cnt,
                            const int* idx, const int* idx_batch_cnt, float* out) {
    group_points_kernel_launcher_stack(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out);
}

void ref_group_points_grad_stack(int B, int M, int C, int N, int nsample, const float* grad_out, const int* idx,
                                 const int* idx_batch_cnt, const int* features_batch_cnt, float* grad_features) {
    group_points_grad_kernel_launcher_stack(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt,
                                            grad_features);
}

}  // extern "C"
