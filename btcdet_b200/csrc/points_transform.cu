// Dataset-side coordinate transforms of the occupancy branch on the GPU:
//   absxyz_2_cylinxyz_np   btcdet/utils/coords_utils.py:282-292   (rho, phi [deg], z)
//   absxyz_2_spherexyz_np  btcdet/utils/coords_utils.py:268-279   (r, azimuth [deg], elevation [deg])
// The reference runs them with numpy on a DataLoader worker before the cylindrical VoxelGeneratorV2 call
// (btcdet/datasets/processor/data_processor.py:128-136); here raw points stay on the device and feed both voxelisers.
//
// Arithmetic follows numpy's float32 op order exactly: np.linalg.norm(p[:, :2], axis=1) = sqrt(x*x + y*y) with separately
// rounded products and sum (no FMA contraction), `arctan2(-y, x) * 180. / np.pi` = TWO roundings (multiply by 180, then
// divide by float32(pi)) — unlike the GPU-side torch variant inside occ_masks.cu, which multiplies once by 180/pi.
// rho, r, z and the extra feature columns are therefore bit-identical to numpy.  The angle goes through atan2f: CUDA's
// libm and numpy's (SVML / glibc, host-CPU dependent) both stay within a few ulp of the true value but are not the same
// function, so phi agrees to <= 4 ulp and a point lying within that distance of a bin edge may quantise differently
// (tests/test_points_transform_gpu.py counts them).
#include "common.cuh"

namespace btc {

__global__ void points_to_angular_kernel(const float* __restrict__ pts, int n_cap, const int* __restrict__ n_dev, int n_feat,
                                         int sphere, float* __restrict__ out) {
    const int n = live_count(n_cap, n_dev);
    const float k180 = 180.0f, kpi = 3.14159274101257324f;   // float32(np.pi)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* p = pts + (int64_t)i * n_feat;
        const float x = p[0], y = p[1], z = p[2];
        const float xx = __fmul_rn(x, x), yy = __fmul_rn(y, y);
        const float sxy = __fadd_rn(xx, yy);
        const float rho = __fsqrt_rn(sxy);
        const float phi = __fdiv_rn(__fmul_rn(atan2f(-y, x), k180), kpi);
        float* o = out + (int64_t)i * n_feat;
        if (sphere) {
            o[0] = __fsqrt_rn(__fadd_rn(sxy, __fmul_rn(z, z)));      // add.reduce over the three squares, in order
            o[1] = phi;
            o[2] = __fdiv_rn(__fmul_rn(atan2f(z, rho), k180), kpi);
        } else {
            o[0] = rho;
            o[1] = phi;
            o[2] = z;
        }
        for (int j = 3; j < n_feat; ++j) o[j] = p[j];
    }
}

}  // namespace btc

using namespace btc;

extern "C" int btc_points_to_cylinder(const float* points, int n_cap, const int* n_dev, int n_feat, int sphere, float* out,
                                      void* stream) {
    if (n_cap < 0 || n_feat < 3 || (sphere != 0 && sphere != 1)) return badarg("btc_points_to_cylinder: bad sizes");
    if (n_cap == 0) return BTC_OK;
    if (!points || !out) return badarg("btc_points_to_cylinder: null argument");
    points_to_angular_kernel<<<grid_for(n_cap, 256), 256, 0, (cudaStream_t)stream>>>(points, n_cap, n_dev, n_feat, sphere, out);
    BTC_CHECK_LAUNCH("points_to_angular");
    return BTC_OK;
}
