"""Restatement of the reference's occupancy / occlusion mask generation.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header): imported by tests/, smoke and bench baselines.

Written with plain torch ops and an explicit `device`, op for op in the reference's order (torch eager
never contracts a*b+c into an FMA, App. C), so that
  * on the CPU it is the oracle, pinned bit-for-bit against the reference's own code executed in this
    container (tests/test_occ_oracle_cpu.py, tests/golden/ref_loader.py) and against the committed
    fixture tests/golden/occ_masks.npz generated from that reference run;
  * on the GPU box the same code runs on CUDA tensors and shares CUDA libm with the kernels, which is
    the strict (bit-exact) reference for `btc_occ_*` where a CPU libm could differ by an ulp at a bin edge.

Reference functions followed (paths relative to the reference checkout):
  get_paddings_indicator / get_valid / get_voxelwise_mask  btcdet/models/occ_pnt/occ_training_targets/
                                                           occ_targets_template.py:73-80,194-202
  create_predict_area3d                                    occ_targets_template.py:432-447
  occ_from_cylin_ocp / point2coords_inrange /
  occ_from_sphere_ocp / get_empty_mask / create_predict_area2d   :136-155, :82-90, :110-134, :186-191, :404-407
  filter_occ                                               :249-255
  prepare_cls_loss_map (mask algebra only)                 :330-380
  uvd2absxyz / cartesian_sphere_coords / sphere_uvd2absxyz / cartesian_cylinder_coords
                                                           btcdet/utils/coords_utils.py:198-204,216-226,180-186,229-239
  all_voxel_centers                                        btcdet/models/detectors/detector3d_template.py:52-63
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OccGeometry:
    """Geometry of DATA_CONFIG.OCC (tools/cfgs/model_configs/btcdet_kitti_car.yaml:72-92)."""
    voxel_size: List[float] = field(default_factory=lambda: [0.32, 0.5184, 0.36])             # (rho, phi deg, z)
    point_cloud_range: List[float] = field(default_factory=lambda: [2.24, -40.6944, -2.6, 69.12, 40.6944, 0.64])
    support_sphere_range: List[float] = field(default_factory=lambda: [2.24, -40.6944, -16.5953125, 70.72, 40.6944, 4.0,
                                                                       0.4203125])
    dist_kern: List[int] = field(default_factory=lambda: [5, 9, 5])                            # (z, y, x)
    half_x: bool = True
    empt_sur_thresh: float = 1
    det_point_cloud_range: List[float] = field(default_factory=lambda: [0, -40, -3, 70.4, 40, 1])

    @property
    def grid_size(self):
        r, v = np.array(self.point_cloud_range), np.array(self.voxel_size)
        return [int(g) for g in np.round((r[3:6] - r[0:3]) / v).astype(np.int64)]               # data_processor.py:119-120

    @property
    def sphere_voxel_size(self):
        return [self.voxel_size[0], self.voxel_size[1], self.support_sphere_range[6]]            # template.py:48

    @property
    def sphere_grid_size(self):
        sr = np.asarray(self.support_sphere_range)
        return [int(g) for g in ((sr[3:6] - sr[:3]) / np.array(self.sphere_voxel_size)).astype(int)]   # template.py:51 (trunc)

    @property
    def concede_x(self):
        return self.dist_kern[-1] // 2 if self.half_x else 0                                      # template.py:68


def _f32(v, device):
    return torch.as_tensor(v, dtype=torch.float32, device=device)


def all_voxel_centers_z(geo: OccGeometry, device):
    """z of every occ-voxel centre, [nz, ny, nx] (detector3d_template.py:52-63, coords_utils.py:166-177)."""
    nx, ny, nz = geo.grid_size
    vz = _f32(geo.voxel_size[2], device)
    oz = _f32(geo.point_cloud_range[2], device)
    z_ind = torch.arange(nz, device=device)
    zc = (0.5 + z_ind.to(torch.float32)) * vz + oz
    return zc.view(nz, 1, 1).expand(nz, ny, nx).contiguous()


# `tensor / python_scalar` is device dependent in torch: the CPU kernel divides, the CUDA kernel multiplies by the
# fp32 reciprocal computed on the host (ATen BinaryDivTrueKernel.cu, is_cpu_scalar fast path).  The mask code
# back-projects sphere-bin LOWER CORNERS, which land exactly on cylinder-bin edges, so this one-ulp difference
# moves ~10 % of the occluded cells (measured: 20 623 of 204 511 on the fixture scene).  The reference only runs
# on CUDA; `CUDA_SCALAR_DIV` = True makes CPU tensors follow the CUDA rule (CUDA tensors always do).
CUDA_SCALAR_DIV = False


def _deg2rad(x):
    """`x * np.pi / 180.` (coords_utils.py:181-185,200-202) under the active scalar-division rule."""
    if CUDA_SCALAR_DIV and not x.is_cuda:
        return (x * np.pi) * float(np.float32(1.0) / np.float32(180.0))
    return x * np.pi / 180.


def cylinder_uvd2absxyz(rho, phi, z):
    """coords_utils.py:198-204: u = (phi*pi)/180 ; x = rho*cos(u) ; y = (-rho)*sin(u)."""
    u = _deg2rad(phi)
    return torch.stack([rho * torch.cos(u), -rho * torch.sin(u), z], dim=-1)


def sphere_uvd2absxyz(r, az, el):
    """coords_utils.py:180-186."""
    xydist = r * torch.cos(_deg2rad(el))
    x = xydist * torch.cos(_deg2rad(az))
    y = -xydist * torch.sin(_deg2rad(az))
    z = r * torch.sin(_deg2rad(el))
    return torch.stack([x, y, z], dim=-1)


def cartesian_sphere_coords(p):
    """coords_utils.py:216-226 (perm xyz): (r, az deg, el deg)."""
    sq = torch.square(p)
    dist = torch.sqrt(torch.sum(sq, dim=1))
    xydist = torch.sqrt(torch.sum(sq[..., 0:2], dim=-1))
    az = torch.atan2(-p[..., 1], p[..., 0]) * (180. / np.pi)
    el = torch.atan2(p[..., 2], xydist) * (180. / np.pi)
    return torch.stack([dist, az, el], dim=-1)


def cartesian_cylinder_coords(p):
    """coords_utils.py:229-239 (perm xyz): (rho, phi deg, z)."""
    sq = torch.square(p)
    xydist = torch.sqrt(torch.sum(sq[..., 0:2], dim=-1))
    phi = torch.atan2(-p[..., 1], p[..., 0]) * (180. / np.pi)
    return torch.stack([xydist, phi, p[..., 2]], dim=-1)


def point2coords_inrange(points, origin, pmax, max_grid, voxel_size):
    """occ_targets_template.py:82-90: inclusive range test, trunc((p-origin)/vs), clamp to the grid."""
    inrange = torch.cat([points[:, :3] >= origin, points[:, :3] <= pmax], dim=-1).all(-1)
    inds = torch.nonzero(inrange)[..., 0]
    pts = points[inds, :]
    coords = ((pts - origin) / voxel_size).to(torch.int64)
    coords = torch.minimum(coords, max_grid)
    coords = torch.maximum(coords, torch.zeros_like(max_grid))
    return coords, inds


def get_valid(voxels, voxel_coords, voxel_num_points):
    """get_paddings_indicator + get_valid (:73-80, :194-197)."""
    P = voxels.shape[1]
    mask = voxel_num_points.int().unsqueeze(1) > torch.arange(P, dtype=torch.int, device=voxels.device).view(1, -1)
    idx = torch.nonzero(mask)
    return voxel_coords[idx[:, 0]].to(torch.int64), voxels[idx[:, 0], idx[:, 1]], mask


def occ_targets(voxels, voxel_coords, voxel_num_points, batch_size, geo: OccGeometry, rot_z=None,
                fore_mask=None, mirr_mask=None, bm_mask=None):
    """a5-a8 (+ the mask algebra of a12 when the box masks are supplied).

    voxels [M,P,C>=3] cylindrical (rho, phi deg, z, ...), voxel_coords [M,4] (b,z,y,x), voxel_num_points [M].
    Returns uint8/bool tensors [B, nz, ny, nx].
    """
    dev = voxels.device
    nx, ny, nz = geo.grid_size
    bs = int(batch_size)
    # occ_targets_3d.py:45  cylinder -> Cartesian of every slot (USE_ABSXYZ)
    occ_pnts = torch.cat([cylinder_uvd2absxyz(voxels[..., 0], voxels[..., 1], voxels[..., 2]), voxels[..., 3:]], dim=-1)
    valid_coords, valid_feats, point_mask = get_valid(occ_pnts, voxel_coords, voxel_num_points)
    # a5 get_voxelwise_mask
    voxelwise_mask = torch.zeros([bs, nz, ny, nx], dtype=torch.uint8, device=dev)
    voxelwise_mask[valid_coords[:, 0], valid_coords[:, 1], valid_coords[:, 2], valid_coords[:, 3]] = 1
    # a6 create_predict_area3d (:432-447): window z,y centred, x shifted by concede_x; border-clamped
    kz, ky, kx = geo.dist_kern
    startz, starty, startx = -(kz // 2), -(ky // 2), -(kx // 2) + geo.concede_x
    z, y, x = torch.meshgrid(torch.arange(startz, startz + kz, device=dev), torch.arange(starty, starty + ky, device=dev),
                             torch.arange(startx, startx + kx, device=dev), indexing="ij")
    bzyx = torch.stack([torch.zeros_like(z), z, y, x], dim=-1).view(1, -1, 4)
    vc = (valid_coords.view(-1, 1, 4) + bzyx).view(-1, 4)
    vcc_mask = torch.zeros([bs, nz, ny, nx], dtype=torch.uint8, device=dev)
    vcc_mask[torch.clamp(vc[:, 0], 0, bs - 1), torch.clamp(vc[:, 1], 0, nz - 1), torch.clamp(vc[:, 2], 0, ny - 1),
             torch.clamp(vc[:, 3], 0, nx - 1)] = 1
    # a7 occ_from_cylin_ocp (:136-155)
    sr = geo.support_sphere_range
    s_origin, s_max = _f32([sr[:3]], dev), _f32([sr[3:6]], dev)
    s_vs = _f32(geo.sphere_voxel_size, dev)
    snx, sny, snz = geo.sphere_grid_size
    s_maxgrid = torch.as_tensor([[snx - 1, sny - 1, snz - 1]], dtype=torch.int64, device=dev)
    pts = valid_feats[:, :3]
    occ_b = valid_coords[:, 0]
    sph = cartesian_sphere_coords(pts + _f32([[0.0, 0.0, 0.0]], dev))
    if rot_z is not None:
        sph[..., 1] += rot_z[occ_b]
    sph_map = torch.zeros([bs, snz, sny, snx], dtype=torch.uint8, device=dev)
    c, inds = point2coords_inrange(sph, s_origin, s_max, s_maxgrid, s_vs)
    b_in = occ_b[inds]
    sph_map[b_in, c[:, 2], c[:, 1], c[:, 0]] = 1
    # occ_from_sphere_ocp (:110-134) with EMPT_SUR_THRESH < 9: range bin 0 <- empty column with >thr returns around
    if geo.empt_sur_thresh != "None" and geo.empt_sur_thresh < 9:
        occ_2d = torch.sum(sph_map, dim=3)                                   # [B, snz, sny] returns per (el, az) column
        empty_2d = occ_2d == 0
        neigh = F.conv2d(occ_2d.unsqueeze(1).to(torch.float32), torch.ones(1, 1, 3, 3, device=dev), padding=1)
        sph_map[:, :, :, 0] = (empty_2d & (neigh.squeeze(1) > geo.empt_sur_thresh)).to(torch.uint8)
    sph_occ = torch.cumsum(sph_map, dim=3) > 0.9                             # at or behind the first return
    ind = torch.nonzero(sph_occ)                                             # [K, 4] (b, el, az, r)
    rev_vs = _f32([s_vs[2].item(), s_vs[1].item(), s_vs[0].item()], dev)
    rev_origin = _f32([[sr[2], sr[1], sr[0]]], dev)
    sp = ind[:, 1:] * rev_vs + rev_origin                                    # lower corner (el, az, r) (:147)
    carte = sphere_uvd2absxyz(sp[..., 2], sp[..., 1], sp[..., 0])
    cyl = cartesian_cylinder_coords(carte - _f32([[0.0, 0.0, 0.0]], dev))
    p_origin, p_max = _f32([geo.point_cloud_range[:3]], dev), _f32([geo.point_cloud_range[3:]], dev)
    p_vs = _f32([geo.voxel_size], dev)
    p_maxgrid = torch.as_tensor([[nx - 1, ny - 1, nz - 1]], dtype=torch.int64, device=dev)
    cc, ii = point2coords_inrange(cyl, p_origin, p_max, p_maxgrid, p_vs)
    occ_mask = torch.zeros_like(voxelwise_mask)
    occ_mask[ind[ii, 0], cc[:, 2], cc[:, 1], cc[:, 0]] = 1
    occ_mask = occ_mask > 0.9
    # a8 filter_occ (:249-255)
    zc = all_voxel_centers_z(geo, dev)
    vz = (1 - voxelwise_mask) * 100.0 + zc.unsqueeze(0)
    vz = torch.min(vz.view(bs, nz * ny, nx), dim=1, keepdim=True)[0].unsqueeze(1)
    vz = vz - (vz > 20.0) * 200
    occ_voxelwise_mask = occ_mask & (zc.unsqueeze(0) > torch.clamp(vz, min=geo.det_point_cloud_range[2], max=None)) \
        & (zc.unsqueeze(0) < geo.det_point_cloud_range[5])
    out = {"voxelwise_mask": voxelwise_mask, "vcc_mask": vcc_mask, "sphere_map": sph_map, "occ_voxelwise_mask": occ_voxelwise_mask,
           "valid_coords": valid_coords, "valid_feats": valid_feats, "voxel_point_mask": point_mask, "occ_pnts": occ_pnts}
    out["general_cls_loss_mask"] = vcc_mask & occ_voxelwise_mask            # :333
    if fore_mask is not None:
        g = out["general_cls_loss_mask"]
        fore = fore_mask & g
        mirr = mirr_mask & g
        bm = bm_mask & g
        pos = fore | mirr | bm
        out.update({"occ_fore_cls_mask": fore, "occ_mirr_cls_mask": mirr, "occ_bm_cls_mask": bm, "pos_mask": pos,
                    "neg_mask": g & (1 - pos)})
    return out
