// Occupancy / occlusion mask generation — SURVEY §8 rows a5-a8 and the mask algebra of a12.
// Replaces, with a handful of fused kernels and no host synchronisation, the torch index-op chain of
//   btcdet/models/occ_pnt/occ_training_targets/occ_targets_template.py
//     get_valid / get_voxelwise_mask :194-202, create_predict_area3d :432-447 (225x index blow-up),
//     occ_from_cylin_ocp :136-155 + point2coords_inrange :82-90 + occ_from_sphere_ocp :110-134 +
//     get_empty_mask :186-191 (nonzero / cumsum over a promoted int64 volume), filter_occ :249-255,
//     prepare_cls_loss_map :333 (general_cls_loss_mask).
//
// Bit-exactness: torch eager evaluates every *, +, -, / as its own fp32 kernel — no FMA contraction.
// All arithmetic that feeds a floor/trunc below therefore uses __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn
// and the same CUDA libm entry points torch uses (sqrtf, atan2f, sinf, cosf); every mask write is an
// idempotent "store 1", so the result is independent of thread order.
#include "common.cuh"
#include "occ_geom.cuh"

namespace btc {

// a5 + the sphere scatter of a7: one thread per (voxel, slot)
__global__ void occ_points_kernel(const float* __restrict__ voxels, int P, int C, const int4* __restrict__ coords,
                                  const int* __restrict__ num_points, int m_cap, const int* __restrict__ m_dev,
                                  const float* __restrict__ rot_z, OccGeom g, unsigned char* __restrict__ voxelwise,
                                  unsigned char* __restrict__ sphere_map) {
    const int m_live = live_count(m_cap, m_dev);
    const int64_t work = (int64_t)m_live * P;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(t / P), p = (int)(t - (int64_t)m * P);
        if (p >= __ldg(num_points + m)) continue;
        const int4 c = __ldg(coords + m);      // (b, z, y, x)
        if ((unsigned)c.x >= (unsigned)g.batch) continue;
        if (p == 0 && (unsigned)c.y < (unsigned)g.g[2] && (unsigned)c.z < (unsigned)g.g[1] && (unsigned)c.w < (unsigned)g.g[0])
            voxelwise[(((int64_t)c.x * g.g[2] + c.y) * g.g[1] + c.z) * g.g[0] + c.w] = 1;
        const float* v = voxels + ((int64_t)m * P + p) * C;
        const float rho = __ldg(v), phi = __ldg(v + 1), z = __ldg(v + 2);
        // cylinder_uvd2absxyz (coords_utils.py:198-204)
        const float u = deg2rad_like_torch(phi);
        const float x = __fmul_rn(rho, cosf(u));
        const float y = __fmul_rn(-rho, sinf(u));
        // cartesian_sphere_coords (coords_utils.py:216-226) on (p + sphere_offset), sphere_offset = 0
        const float xo = __fadd_rn(x, 0.0f), yo = __fadd_rn(y, 0.0f), zo = __fadd_rn(z, 0.0f);
        const float sx = __fmul_rn(xo, xo), sy = __fmul_rn(yo, yo), sz = __fmul_rn(zo, zo);
        const float sxy = __fadd_rn(sx, sy);
        float sp[3];
        sp[0] = sqrtf(__fadd_rn(sxy, sz));
        sp[1] = __fmul_rn(atan2f(-yo, xo), kRad2Deg);
        sp[2] = __fmul_rn(atan2f(zo, sqrtf(sxy)), kRad2Deg);
        if (rot_z) sp[1] = __fadd_rn(sp[1], __ldg(rot_z + c.x));
        int sc[3];
        if (quantize_inrange(sp, g.slo, g.shi, g.svs, g.sg, sc))
            sphere_map[(((int64_t)c.x * g.sg[2] + sc[2]) * g.sg[1] + sc[1]) * g.sg[0] + sc[0]] = 1;
    }
}

// a6 create_predict_area3d: one thread per (voxel, window cell); out-of-grid cells clamp ONTO the border
__global__ void occ_dilate_kernel(const int4* __restrict__ coords, const int* __restrict__ num_points, int m_cap,
                                  const int* __restrict__ m_dev, OccGeom g, unsigned char* __restrict__ vcc) {
    const int m_live = live_count(m_cap, m_dev);
    const int win = g.kern[0] * g.kern[1] * g.kern[2];
    const int64_t work = (int64_t)m_live * win;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(t / win), w = (int)(t - (int64_t)m * win);
        if (__ldg(num_points + m) <= 0) continue;
        const int4 c = __ldg(coords + m);
        const int dx = w % g.kern[2], dy = (w / g.kern[2]) % g.kern[1], dz = w / (g.kern[2] * g.kern[1]);
        int b = min(max(c.x, 0), g.batch - 1);
        int z = min(max(c.y + dz - g.kern[0] / 2, 0), g.g[2] - 1);
        int y = min(max(c.z + dy - g.kern[1] / 2, 0), g.g[1] - 1);
        int x = min(max(c.w + dx - g.kern[2] / 2 + g.concede_x, 0), g.g[0] - 1);
        vcc[(((int64_t)b * g.g[2] + z) * g.g[1] + y) * g.g[0] + x] = 1;
    }
}

// MeanVFE of the occupancy branch on absolute coordinates (occ_targets_3d.py:45-47 with USE_ABSXYZ rewrites
// batch_dict['voxels'] to cylinder_uvd2absxyz(voxels) + extra columns for EVERY slot, padding included; mean_vfe.py:27-44
// then sums all P slots and divides by clamp_min(count, 1)).  One thread per (voxel, column); the slot sum runs in slot
// order (torch's reduction order over a 12-element strided axis is an implementation detail: compared at 1e-6).
__global__ void occ_abs_vfe_kernel(const float* __restrict__ voxels, int P, int C, const int* __restrict__ num_points, int m_cap,
                                   const int* __restrict__ m_dev, float* __restrict__ voxels_abs, float* __restrict__ mean) {
    const int m_live = live_count(m_cap, m_dev);
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < m_live; m += gridDim.x * blockDim.x) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int p = 0; p < P; ++p) {
            const float* v = voxels + ((int64_t)m * P + p) * C;
            const float rho = __ldg(v), phi = __ldg(v + 1);
            const float u = deg2rad_like_torch(phi);
            float o[8];
            o[0] = __fmul_rn(rho, cosf(u));
            o[1] = __fmul_rn(-rho, sinf(u));
            for (int j = 2; j < C; ++j) o[j] = __ldg(v + j);
            for (int j = 0; j < C; ++j) acc[j] = __fadd_rn(acc[j], o[j]);
            if (voxels_abs)
                for (int j = 0; j < C; ++j) voxels_abs[((int64_t)m * P + p) * C + j] = o[j];
        }
        const int cnt = __ldg(num_points + m);
        const float norm = cnt > 1 ? (float)cnt : 1.0f;
        for (int j = 0; j < C; ++j) mean[(int64_t)m * C + j] = __fdiv_rn(acc[j], norm);
    }
}

// per (b, el, az) column of the sphere map: number of returns and first return at range bin >= 1
__global__ void sphere_columns_kernel(const unsigned char* __restrict__ sphere_map, OccGeom g, int* __restrict__ counts,
                                      int* __restrict__ first1) {
    const int64_t cols = (int64_t)g.batch * g.sg[2] * g.sg[1];
    for (int64_t col = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; col < cols; col += (int64_t)gridDim.x * blockDim.x) {
        const unsigned char* p = sphere_map + col * g.sg[0];
        int cnt = 0, first = g.sg[0];
        for (int r = 0; r < g.sg[0]; ++r) {
            int v = p[r] != 0;
            cnt += v;
            if (v && r >= 1 && first == g.sg[0]) first = r;
        }
        counts[col] = cnt;
        first1[col] = p[0] ? -(first + 1) : first;   // sign bit: range bin 0 holds a return of its own
    }
}

// occ_from_sphere_ocp + back-projection: one thread per sphere cell (b, el, az, r)
__global__ void sphere_occlusion_kernel(const int* __restrict__ counts, const int* __restrict__ first1, OccGeom g,
                                        unsigned char* __restrict__ occ_raw) {
    const int snx = g.sg[0], sny = g.sg[1], snz = g.sg[2];
    const int64_t cells = (int64_t)g.batch * snz * sny * snx;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cells; t += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(t % snx);
        const int64_t col = t / snx;
        const int az = (int)(col % sny);
        const int el = (int)((col / sny) % snz);
        const int b = (int)(col / ((int64_t)sny * snz));
        int first = __ldg(first1 + col);
        const bool own_bin0 = first < 0;
        if (own_bin0) first = -first - 1;
        if (g.use_empty) {
            // range bin 0 <- empty column whose 3x3 (el, az) neighbourhood holds > thresh returns (get_empty_mask)
            bool bin0 = false;
            if (__ldg(counts + col) == 0) {
                float s = 0.f;   // fixed all-ones 3x3 conv over float counts, zero padded (exact small integers)
                for (int de = -1; de <= 1; ++de)
                    for (int da = -1; da <= 1; ++da) {
                        int e2 = el + de, a2 = az + da;
                        if (e2 >= 0 && e2 < snz && a2 >= 0 && a2 < sny)
                            s += (float)__ldg(counts + ((int64_t)b * snz + e2) * sny + a2);
                    }
                bin0 = s > g.empt_thresh;
            }
            if (bin0) first = 0;
        } else if (own_bin0) {
            first = 0;   // EMPT_SUR_THRESH >= 9: bin 0 is not overwritten and keeps its own return
        }
        if (r < first) continue;   // cumsum(dim=r) > 0.9: at or behind the first return
        // bin lower corner -> Cartesian -> cylinder (occ_targets_template.py:147-149)
        const float sr = __fadd_rn(__fmul_rn((float)r, g.svs[0]), g.slo[0]);
        const float sa = __fadd_rn(__fmul_rn((float)az, g.svs[1]), g.slo[1]);
        const float se = __fadd_rn(__fmul_rn((float)el, g.svs[2]), g.slo[2]);
        const float te = deg2rad_like_torch(se), ta = deg2rad_like_torch(sa);
        const float xyd = __fmul_rn(sr, cosf(te));
        const float x = __fmul_rn(xyd, cosf(ta));
        const float y = __fmul_rn(-xyd, sinf(ta));
        const float z = __fmul_rn(sr, sinf(te));
        const float xo = __fsub_rn(x, 0.0f), yo = __fsub_rn(y, 0.0f), zo = __fsub_rn(z, 0.0f);
        float cp[3];
        cp[0] = sqrtf(__fadd_rn(__fmul_rn(xo, xo), __fmul_rn(yo, yo)));
        cp[1] = __fmul_rn(atan2f(-yo, xo), kRad2Deg);
        cp[2] = zo;
        int cc[3];
        if (quantize_inrange(cp, g.lo, g.hi, g.vs, g.g, cc))
            occ_raw[(((int64_t)b * g.g[2] + cc[2]) * g.g[1] + cc[1]) * g.g[0] + cc[0]] = 1;
    }
}

// filter_occ part 1: per (b, x) the minimum over (z, y) of (1 - occupied) * 100 + z_centre.
// One block per (b, x) column: the threads sweep y, every z layer's "some cell occupied" / "some cell free" facts are
// OR-reduced over the block and thread 0 applies the reference's arithmetic (a thread per column walked 9 x 157 dependent
// byte loads: 127 us of a 3.8 ms config-3 step).  nz <= 32 (9 in the shipped configuration).
__global__ void __launch_bounds__(160) occ_floor_kernel(const unsigned char* __restrict__ voxelwise, OccGeom g,
                                                        float* __restrict__ floor_z) {
    __shared__ unsigned s_occ[8], s_free[8];
    const int nx = g.g[0], ny = g.g[1], nz = g.g[2];
    const int t = blockIdx.x;
    const int x = t % nx, b = t / nx;
    unsigned occ_bits = 0, free_bits = 0;
    for (int y = threadIdx.x; y < ny; y += blockDim.x) {
        for (int z = 0; z < nz; ++z) {
            const unsigned char v = voxelwise[(((int64_t)b * nz + z) * ny + y) * nx + x];
            if (v) occ_bits |= 1u << z; else free_bits |= 1u << z;
        }
    }
    occ_bits = __reduce_or_sync(0xffffffffu, occ_bits);
    free_bits = __reduce_or_sync(0xffffffffu, free_bits);
    if ((threadIdx.x & 31) == 0) { s_occ[threadIdx.x >> 5] = occ_bits; s_free[threadIdx.x >> 5] = free_bits; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 1; w < nw; ++w) { occ_bits |= s_occ[w]; free_bits |= s_free[w]; }
        float m = 3.4e38f;
        for (int z = 0; z < nz; ++z) {
            const float zc = __fadd_rn(__fmul_rn(__fadd_rn(0.5f, (float)z), g.vs[2]), g.lo[2]);
            const float free_v = __fadd_rn(100.0f, zc);
            if ((occ_bits >> z) & 1u) m = fminf(m, zc);
            if ((free_bits >> z) & 1u) m = fminf(m, free_v);
        }
        if (m > 20.0f) m = __fsub_rn(m, 200.0f);
        floor_z[t] = fmaxf(m, g.det_zmin);
    }
}

// filter_occ part 2 + general_cls_loss_mask
__global__ void occ_filter_kernel(const unsigned char* __restrict__ occ_raw, const unsigned char* __restrict__ vcc,
                                  const float* __restrict__ floor_z, OccGeom g, unsigned char* __restrict__ occ_mask,
                                  unsigned char* __restrict__ general_mask) {
    const int nx = g.g[0], ny = g.g[1], nz = g.g[2];
    const int64_t cells = (int64_t)g.batch * nz * ny * nx;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cells; t += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(t % nx);
        const int z = (int)((t / ((int64_t)nx * ny)) % nz);
        const int b = (int)(t / ((int64_t)nx * ny * nz));
        const float zc = __fadd_rn(__fmul_rn(__fadd_rn(0.5f, (float)z), g.vs[2]), g.lo[2]);
        const unsigned char o = (occ_raw[t] && zc > floor_z[(int64_t)b * nx + x] && zc < g.det_zmax) ? 1 : 0;
        occ_mask[t] = o;
        if (general_mask) general_mask[t] = o & vcc[t];
    }
}

static int64_t occ_ws(const OccGeom& g, int64_t* o_sphere, int64_t* o_counts, int64_t* o_first, int64_t* o_floor,
                      int64_t* o_raw) {
    int64_t off = 0;
    int64_t sph = (int64_t)g.batch * g.sg[0] * g.sg[1] * g.sg[2];
    int64_t cols = (int64_t)g.batch * g.sg[1] * g.sg[2];
    int64_t cells = (int64_t)g.batch * g.g[0] * g.g[1] * g.g[2];
    *o_sphere = off; off += align_up(sph, 256);
    *o_raw = off; off += align_up(cells, 256);
    *o_counts = off; off += align_up(cols * 4, 256);
    *o_first = off; off += align_up(cols * 4, 256);
    *o_floor = off; off += align_up((int64_t)g.batch * g.g[0] * 4, 256);
    return off;
}

// ---- occupancy-point injection: threshold -> ordered compaction -> pseudo points -----------
// Replaces AddOccTemplate.filter_occ_points (nonzero + per-scene python loop, add_occ_template.py:94-128, without
// the top-k branch), occ_coords2absxyz (:131-146), trans_voxel_grid (:78-88) and assemble_occ_points (:149-165).
struct InjectGeom {
    float ovs[3], oorg[3];     // occ grid voxel size / origin (rho, phi, z)
    float dvs[3], dmin[3];     // det grid voxel size / range min (x, y, z)
    int g[3];                  // occ grid nx, ny, nz
    int dg[3];                 // det grid nx, ny, nz
    float thresh, inten;
    int batch;
};

__global__ void occ_flag_kernel(const float* __restrict__ probs, int64_t cells, float thresh, int* __restrict__ flags) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cells; t += (int64_t)gridDim.x * blockDim.x)
        flags[t] = __ldg(probs + t) > thresh ? 1 : 0;
}

__global__ void occ_emit_kernel(const float* __restrict__ probs, const float* __restrict__ residuals,
                                const int* __restrict__ flags, const int* __restrict__ rank,
                                const int* __restrict__ total, const float* __restrict__ rot_z, InjectGeom g, int cap,
                                int4* __restrict__ occ_coords, float* __restrict__ occ_probs, float* __restrict__ occ_xyz,
                                int4* __restrict__ det_coords, float* __restrict__ occ_points, int* __restrict__ counts) {
    const int nx = g.g[0], ny = g.g[1], nz = g.g[2];
    const int64_t per_scene = (int64_t)nz * ny * nx;
    const int64_t cells = per_scene * g.batch;
    if (blockIdx.x == 0 && threadIdx.x <= g.batch) {      // per-scene counts + total
        int b = threadIdx.x;
        int r1 = b < g.batch ? __ldg(rank + (int64_t)b * per_scene) : 0;
        if (b == g.batch) counts[b] = *total;
        else counts[b] = ((b + 1 < g.batch) ? __ldg(rank + (int64_t)(b + 1) * per_scene) : *total) - r1;
    }
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cells; t += (int64_t)gridDim.x * blockDim.x) {
        if (!flags[t]) continue;
        const int row = rank[t];
        if (row >= cap) continue;
        const int x = (int)(t % nx), y = (int)((t / nx) % ny), z = (int)((t / ((int64_t)nx * ny)) % nz);
        const int b = (int)(t / per_scene);
        const float p = __ldg(probs + t);
        // voxel centre in cylinder coordinates: origin + (idx + 0.5) * vs   (python-float scalars -> fp32)
        float rho = __fadd_rn(g.oorg[0], __fmul_rn(__fadd_rn((float)x, 0.5f), g.ovs[0]));
        float phi = __fadd_rn(g.oorg[1], __fmul_rn(__fadd_rn((float)y, 0.5f), g.ovs[1]));
        float zc = __fadd_rn(g.oorg[2], __fmul_rn(__fadd_rn((float)z, 0.5f), g.ovs[2]));
        if (rot_z) phi = __fsub_rn(phi, __ldg(rot_z + b));
        const float u = deg2rad_like_torch(phi);
        float px = __fmul_rn(rho, cosf(u));
        float py = __fmul_rn(-rho, sinf(u));
        float pz = zc;
        if (residuals) {
            const float* r = residuals + (int64_t)b * 3 * per_scene + (t - (int64_t)b * per_scene);
            px = __fadd_rn(px, __ldg(r));
            py = __fadd_rn(py, __ldg(r + per_scene));
            pz = __fadd_rn(pz, __ldg(r + 2 * per_scene));
        }
        // trans_voxel_grid: floor((p - min) / vs) clamped into the det grid
        const float q[3] = {px, py, pz};
        int dc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float f = floorf(__fdiv_rn(__fsub_rn(q[a], g.dmin[a]), g.dvs[a]));
            f = fminf(fmaxf(f, 0.0f), (float)(g.dg[a] - 1));
            dc[a] = (int)f;
        }
        occ_coords[row] = make_int4(b, z, y, x);
        occ_probs[row] = p;
        occ_xyz[row * 3 + 0] = px; occ_xyz[row * 3 + 1] = py; occ_xyz[row * 3 + 2] = pz;
        det_coords[row] = make_int4(b, dc[2], dc[1], dc[0]);
        float* o = occ_points + (int64_t)row * 6;
        o[0] = px; o[1] = py; o[2] = pz; o[3] = g.inten; o[4] = p; o[5] = 1.0f;
    }
}

// OccHead3D.forward tail (occ_head_3D.py:46-49) without the dense logits volume: prob = softmax(dense(logits), dim=1)[:, -1]
// * mask.  A cell without an active site has logits (0, .., 0) -> softmax 1 / n_cls (0.5 for the two-class head), so the
// volume is `mask / n_cls` everywhere and the active rows overwrite their cells with their own softmax (max-subtracted,
// expf, sum in class order, one division — the op order of torch's softmax kernel).
__global__ void occ_head_fill_kernel(const unsigned char* __restrict__ mask, int64_t cells, float base, float* __restrict__ prob) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cells; t += (int64_t)gridDim.x * blockDim.x)
        prob[t] = mask ? __fmul_rn(base, (float)mask[t]) : base;
}
__global__ void occ_head_rows_kernel(const float* __restrict__ logits, const int4* __restrict__ coords, int n_cap,
                                     const int* __restrict__ n_dev, int n_cls, int nx, int ny, int nz,
                                     const unsigned char* __restrict__ mask, float* __restrict__ prob) {
    const int n = live_count(n_cap, n_dev);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const float* l = logits + (int64_t)r * n_cls;
        float m = l[0];
        for (int c = 1; c < n_cls; ++c) m = fmaxf(m, l[c]);
        float sum = 0.f, last = 0.f;
        for (int c = 0; c < n_cls; ++c) {
            last = expf(__fsub_rn(l[c], m));
            sum = __fadd_rn(sum, last);
        }
        const int4 q = coords[r];          // (b, z, y, x)
        const int64_t cell = (((int64_t)q.x * nz + q.y) * ny + q.z) * nx + q.w;
        const float p = __fdiv_rn(last, sum);
        prob[cell] = mask ? __fmul_rn(p, (float)mask[cell]) : p;
    }
}

// OccVFE (occ_vfe.py:24-55): slots with code < 0.05 are raw points, the others injected occupancy points.
__global__ void occ_vfe_kernel(const float* __restrict__ voxels, const int* __restrict__ num_points, int m_cap,
                               const int* __restrict__ m_dev, int P, int C, int n_raw, float* __restrict__ feats,
                               float* __restrict__ occ_feats) {
    const int m = live_count(m_cap, m_dev);
    const int n_code = C - n_raw;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < m; v += gridDim.x * blockDim.x) {
        const float* vox = voxels + (int64_t)v * P * C;
        const int np = __ldg(num_points + v);
        int raw_n = 0, occ_n = 0;
        for (int p = 0; p < P && p < np; ++p) {
            if (__ldg(vox + p * C + C - 1) < 0.05f) ++raw_n; else ++occ_n;
        }
        const bool occ_only = occ_n > 0 && raw_n == 0;
        const float raw_d = (float)max(raw_n, 1), occ_d = (float)max(occ_n, 1);
        for (int c = 0; c < n_raw; ++c) {
            float sr = 0.f, so = 0.f;
            for (int p = 0; p < P; ++p) {
                const bool valid = p < np;
                const float val = __ldg(vox + p * C + c);
                const bool is_raw = __ldg(vox + p * C + C - 1) < 0.05f;
                sr = __fadd_rn(sr, (valid && is_raw) ? val : 0.0f);
                so = __fadd_rn(so, (valid && !is_raw) ? val : 0.0f);
            }
            const float fr = __fdiv_rn(sr, raw_d), fo = __fdiv_rn(so, occ_d);
            feats[(int64_t)v * C + c] = __fadd_rn(fr, occ_only ? fo : 0.0f);
        }
        for (int c = 0; c < n_code; ++c) {   // max over ALL slots, zero padding included (as the reference)
            float mx = __ldg(vox + n_raw + c);
            for (int p = 1; p < P; ++p) mx = fmaxf(mx, __ldg(vox + p * C + n_raw + c));
            feats[(int64_t)v * C + n_raw + c] = mx;
            if (occ_feats) occ_feats[(int64_t)v * n_code + c] = mx;
        }
    }
}

}  // namespace btc

using namespace btc;

extern "C" {

int64_t btc_occ_targets_workspace_bytes(int batch, const float* geom_f, const int* geom_i) {
    OccGeom g;
    if (parse_geom(g, batch, geom_f, geom_i)) return BTC_E_BADARG;
    int64_t a, b, c, d, e;
    return occ_ws(g, &a, &b, &c, &d, &e);
}

int btc_occ_targets(const float* voxels, int P, int C, const int* voxel_coords, const int* num_points, int m_cap,
                    const int* m_dev, int batch, const float* rot_z, const float* geom_f, const int* geom_i,
                    uint8_t* voxelwise_mask, uint8_t* vcc_mask, uint8_t* occ_mask, uint8_t* general_mask,
                    uint8_t* sphere_map_out, void* workspace, int64_t workspace_bytes, void* stream) {
    OccGeom g;
    if (parse_geom(g, batch, geom_f, geom_i)) return badarg("btc_occ_targets: bad geometry");
    if (!voxelwise_mask || !vcc_mask || !occ_mask || !workspace) return badarg("btc_occ_targets: null argument");
    if (m_cap > 0 && (!voxels || !voxel_coords || !num_points)) return badarg("btc_occ_targets: null inputs");
    if (P < 1 || C < 3) return badarg("btc_occ_targets: bad voxel layout");
    int64_t o_sphere, o_counts, o_first, o_floor, o_raw;
    if (workspace_bytes < occ_ws(g, &o_sphere, &o_counts, &o_first, &o_floor, &o_raw)) return badarg("btc_occ_targets: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    unsigned char* sphere = (unsigned char*)(ws + o_sphere);
    unsigned char* raw = (unsigned char*)(ws + o_raw);
    int* counts = (int*)(ws + o_counts);
    int* first1 = (int*)(ws + o_first);
    float* floor_z = (float*)(ws + o_floor);
    const int64_t cells = (int64_t)batch * g.g[0] * g.g[1] * g.g[2];
    const int64_t sph = (int64_t)batch * g.sg[0] * g.sg[1] * g.sg[2];
    BTC_CUDA(cudaMemsetAsync(voxelwise_mask, 0, cells, st), "occ memset");
    BTC_CUDA(cudaMemsetAsync(vcc_mask, 0, cells, st), "occ memset");
    BTC_CUDA(cudaMemsetAsync(sphere, 0, align_up(sph, 256) + align_up(cells, 256), st), "occ memset");   // sphere + raw
    const int T = 256;
    if (m_cap > 0) {
        occ_points_kernel<<<grid_for((int64_t)m_cap * P, T), T, 0, st>>>(voxels, P, C, (const int4*)voxel_coords, num_points,
                                                                         m_cap, m_dev, rot_z, g, voxelwise_mask, sphere);
        int win = g.kern[0] * g.kern[1] * g.kern[2];
        occ_dilate_kernel<<<grid_for((int64_t)m_cap * win, T), T, 0, st>>>((const int4*)voxel_coords, num_points, m_cap, m_dev,
                                                                           g, vcc_mask);
    }
    sphere_columns_kernel<<<grid_for((int64_t)batch * g.sg[1] * g.sg[2], T), T, 0, st>>>(sphere, g, counts, first1);
    sphere_occlusion_kernel<<<grid_for(sph, T), T, 0, st>>>(counts, first1, g, raw);
    if (g.g[2] > 32) return badarg("btc_occ_targets: more than 32 z layers are not supported");
    occ_floor_kernel<<<batch * g.g[0], 160, 0, st>>>(voxelwise_mask, g, floor_z);
    occ_filter_kernel<<<grid_for(cells, T), T, 0, st>>>(raw, vcc_mask, floor_z, g, occ_mask, general_mask);
    if (sphere_map_out) BTC_CUDA(cudaMemcpyAsync(sphere_map_out, sphere, sph, cudaMemcpyDeviceToDevice, st), "occ copy");
    BTC_CHECK_LAUNCH("occ_targets");
    return BTC_OK;
}

int64_t btc_occ_select_workspace_bytes(int batch, const int* grid) {
    if (!grid || batch < 1) return BTC_E_BADARG;
    int64_t cells = (int64_t)batch * grid[0] * grid[1] * grid[2];
    return align_up(cells * 4, 256) * 2 + align_up((int64_t)(scan_num_blocks(cells) + 2) * 4, 256) + 256;
}

int btc_occ_select(const float* probs, const float* residuals, int batch, const int* grid, float thresh,
                   const float* rot_z, const float* geom_f, const int* det_grid, float inten, int cap, int* occ_coords,
                   float* occ_probs, float* occ_xyz, int* det_coords, float* occ_points, int* counts, void* workspace,
                   int64_t workspace_bytes, void* stream) {
    if (!probs || !grid || !geom_f || !det_grid || !counts || !workspace) return badarg("btc_occ_select: null argument");
    if (cap > 0 && (!occ_coords || !occ_probs || !occ_xyz || !det_coords || !occ_points)) return badarg("btc_occ_select: null outputs");
    if (workspace_bytes < btc_occ_select_workspace_bytes(batch, grid)) return badarg("btc_occ_select: workspace too small");
    InjectGeom g;
    for (int a = 0; a < 3; ++a) {
        g.ovs[a] = geom_f[a]; g.oorg[a] = geom_f[3 + a]; g.dvs[a] = geom_f[6 + a]; g.dmin[a] = geom_f[9 + a];
        g.g[a] = grid[a]; g.dg[a] = det_grid[a];
    }
    g.thresh = thresh; g.inten = inten; g.batch = batch;
    const int64_t cells = (int64_t)batch * grid[0] * grid[1] * grid[2];
    if (cells > 0x7fffffff) return badarg("btc_occ_select: grid too large");
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)workspace;
    int* flags = (int*)ws;
    int* rank = (int*)(ws + align_up(cells * 4, 256));
    int* block_sums = (int*)(ws + 2 * align_up(cells * 4, 256));
    int* total = (int*)(ws + 2 * align_up(cells * 4, 256) + align_up((int64_t)(scan_num_blocks(cells) + 2) * 4, 256));
    occ_flag_kernel<<<grid_for(cells, 256), 256, 0, st>>>(probs, cells, thresh, flags);
    int rc = launch_flag_scan(flags, rank, (int)cells, block_sums, total, st);
    if (rc) return rc;
    occ_emit_kernel<<<grid_for(cells, 256), 256, 0, st>>>(probs, residuals, flags, rank, total, rot_z, g, cap, (int4*)occ_coords,
                                                         occ_probs, occ_xyz, (int4*)det_coords, occ_points, counts);
    BTC_CHECK_LAUNCH("occ_select");
    return BTC_OK;
}

int btc_occ_vfe(const float* voxels, const int* num_points, int m_cap, const int* m_dev, int P, int C, int num_raw,
                float* feats, float* occ_feats, void* stream) {
    if (!feats) return badarg("btc_occ_vfe: null argument");
    if (P < 1 || C < 2 || num_raw < 1 || num_raw >= C) return badarg("btc_occ_vfe: bad layout");
    if (m_cap <= 0) return BTC_OK;
    if (!voxels || !num_points) return badarg("btc_occ_vfe: null inputs");
    occ_vfe_kernel<<<grid_for(m_cap, 128), 128, 0, (cudaStream_t)stream>>>(voxels, num_points, m_cap, m_dev, P, C, num_raw, feats,
                                                                          occ_feats);
    BTC_CHECK_LAUNCH("occ_vfe");
    return BTC_OK;
}

int btc_occ_abs_mean_vfe(const float* voxels, int max_points, int n_feat, const int* num_points, int m_cap, const int* m_dev,
                         float* voxels_abs, float* voxel_mean, void* stream) {
    if (m_cap < 0 || max_points < 1 || n_feat < 3 || n_feat > 8) return badarg("btc_occ_abs_mean_vfe: bad sizes");
    if (m_cap == 0) return BTC_OK;
    if (!voxels || !num_points || !voxel_mean) return badarg("btc_occ_abs_mean_vfe: null argument");
    occ_abs_vfe_kernel<<<grid_for(m_cap, 128), 128, 0, (cudaStream_t)stream>>>(voxels, max_points, n_feat, num_points, m_cap, m_dev,
                                                                              voxels_abs, voxel_mean);
    BTC_CHECK_LAUNCH("occ_abs_vfe");
    return BTC_OK;
}

int btc_occ_head_prob(const float* logits, const int* coords, int n_cap, const int* n_dev, int n_cls, int batch, const int* grid,
                      const unsigned char* mask, float* prob, void* stream) {
    if (!grid || batch < 1 || n_cls < 1 || n_cap < 0) return badarg("btc_occ_head_prob: bad sizes");
    if (!prob) return badarg("btc_occ_head_prob: null output");
    const int64_t cells = (int64_t)batch * grid[0] * grid[1] * grid[2];
    cudaStream_t st = (cudaStream_t)stream;
    occ_head_fill_kernel<<<grid_for(cells, 256), 256, 0, st>>>(mask, cells, 1.0f / (float)n_cls, prob);
    if (n_cap > 0) {
        if (!logits || !coords) return badarg("btc_occ_head_prob: null inputs");
        occ_head_rows_kernel<<<grid_for(n_cap, 256), 256, 0, st>>>(logits, (const int4*)coords, n_cap, n_dev, n_cls, grid[0], grid[1],
                                                                   grid[2], mask, prob);
    }
    BTC_CHECK_LAUNCH("occ_head_prob");
    return BTC_OK;
}

}  // extern "C"
