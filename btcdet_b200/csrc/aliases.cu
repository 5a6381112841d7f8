// Entry points under the names SURVEY.md §8(b) lists as the minimum export set where this library's own name differs:
// thin forwards, no behaviour of their own.
#include "common.cuh"

extern "C" {

int btc_voxelize_cuda(const float* points, int n_points, int n_feat, const int* scene_offsets, int n_scenes,
                      const float* voxel_size, const float* range, const int* grid, int max_points, int max_voxels,
                      float* voxels, int* coords, int* num_points, float* voxel_mean, int* n_voxels, void* workspace,
                      int64_t workspace_bytes, void* stream) {
    return btc_voxelize(points, n_points, n_feat, scene_offsets, n_scenes, voxel_size, range, grid, max_points, max_voxels,
                        voxels, coords, num_points, voxel_mean, n_voxels, workspace, workspace_bytes, stream);
}

int btc_rulebook_pool(const int* coords_in, int n_in_cap, const int* n_in_dev, int batch, const int* in_shape,
                      const int* out_shape, const int* ksize, const int* stride, const int* padding, const int* dilation,
                      uint64_t* out_index, int64_t out_entries, int* out_coords, int out_cap, int* n_out, int* nbr_out,
                      int* nbr_in, void* workspace, int64_t workspace_bytes, void* stream) {
    return btc_rulebook_conv(coords_in, n_in_cap, n_in_dev, batch, in_shape, out_shape, ksize, stride, padding, dilation, 0,
                             out_index, out_entries, out_coords, out_cap, n_out, nbr_out, nbr_in, workspace, workspace_bytes,
                             stream);
}

int btc_occ_inject_revoxelize(const int* pt_coords, int n_cap, const int* n_dev, int batch, const int* shape, uint64_t* index,
                              int64_t n_entries, int* vox_coords, int vox_cap, int* vox_count, int* slots, int* pt_voxel,
                              int* n_voxels, int* max_count, void* workspace, int64_t workspace_bytes, void* stream) {
    return btc_revoxelize(pt_coords, n_cap, n_dev, batch, shape, index, n_entries, vox_coords, vox_cap, vox_count, slots,
                          pt_voxel, n_voxels, max_count, workspace, workspace_bytes, stream);
}

}  // extern "C"
