// Occupancy-grid geometry shared by occ_masks.cu (a5-a8) and box_masks.cu (a9-a12).
#pragma once
#include "common.cuh"

namespace btc {

struct OccGeom {
    float vs[3], lo[3], hi[3];        // occ grid (rho, phi, z): voxel size, range min / max
    float svs[3], slo[3], shi[3];     // support sphere grid (r, az, el)
    float empt_thresh, det_zmin, det_zmax;
    int g[3];                         // nx, ny, nz
    int sg[3];                        // snx, sny, snz
    int kern[3];                      // dist kern (z, y, x)
    int concede_x, use_empty;
    int batch;
};

__device__ __forceinline__ float deg2rad_like_torch(float deg) {
    // coords_utils.py: `x * np.pi / 180.`.  On CUDA, torch evaluates tensor / python_scalar as a multiplication by
    // the reciprocal computed once on the host in fp32 (ATen BinaryDivTrueKernel.cu, "is_cpu_scalar" fast path):
    // (x * fp32(pi)) * fp32(1/180) — NOT an IEEE division.  The reference only runs on CUDA, so that is the
    // behaviour to reproduce (a CPU run of the same code divides, and differs at bin edges).
    return __fmul_rn(__fmul_rn(deg, 3.14159274101257324f), 1.0f / 180.0f);
}
constexpr float kRad2Deg = 57.2957801818847656f;   // fp32(180. / np.pi)

// point2coords_inrange (:82-90): inclusive range test, trunc((p - origin) / vs), clamp into the grid
__device__ __forceinline__ bool quantize_inrange(const float p[3], const float lo[3], const float hi[3],
                                                 const float vs[3], const int n[3], int c[3]) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (!(p[a] >= lo[a] && p[a] <= hi[a])) return false;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        long long q = (long long)__fdiv_rn(__fsub_rn(p[a], lo[a]), vs[a]);
        q = q < (long long)(n[a] - 1) ? q : (long long)(n[a] - 1);
        q = q > 0 ? q : 0;
        c[a] = (int)q;
    }
    return true;
}

inline int parse_geom(OccGeom& g, int batch, const float* gf, const int* gi) {
    if (!gf || !gi || batch < 1) return BTC_E_BADARG;
    for (int a = 0; a < 3; ++a) {
        g.vs[a] = gf[a]; g.lo[a] = gf[3 + a]; g.hi[a] = gf[6 + a];
        g.svs[a] = gf[9 + a]; g.slo[a] = gf[12 + a]; g.shi[a] = gf[15 + a];
        g.g[a] = gi[a]; g.sg[a] = gi[3 + a]; g.kern[a] = gi[6 + a];
        if (g.g[a] < 1 || g.sg[a] < 1 || g.kern[a] < 1) return BTC_E_BADARG;
    }
    g.empt_thresh = gf[18]; g.det_zmin = gf[19]; g.det_zmax = gf[20];
    g.concede_x = gi[9]; g.use_empty = gi[10];
    g.batch = batch;
    return BTC_OK;
}


}  // namespace btc
