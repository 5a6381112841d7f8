"""Restatement of the reference's occupancy-point injection, torch ops, device-agnostic.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  Pinned bit-for-bit against the reference's own
`PassOccVox.forward` / `OccVFE.forward` executed on CPU (tests/test_occ_inject_cpu.py via tests/golden/ref_loader.py).

Reference code followed (paths relative to the reference checkout):
  AddOccTemplate.filter_occ_points     btcdet/models/occ_pnt/add_occ_template.py:94-128   (threshold, nonzero, top-k)
  occ_coords2absxyz                    :131-146   (voxel centre, -rot_z, cylinder -> Cartesian)
  trans_voxel_grid                     :78-88     (floor((p - min)/vs) clamped to the det grid)
  assemble_occ_points / assemble_gt_vox_points   :149-190
  combine_gt_occ_voxel_point / voxelize_pad      :248-268
  PassOccVox.forward                   btcdet/models/occ_pnt/pass_occ_vox.py:10-59
  OccVFE.forward                       btcdet/models/backbones_3d/vfe/occ_vfe.py:24-55
"""
import numpy as np
import torch

from . import occ_masks


def filter_occ_points(probs, residuals, thresh, max_points):
    """Per scene: cells with prob > thresh in row-major order (torch.nonzero); if more than max_points, the
    max_points most probable (torch.topk(sorted=False): the reference's order is unspecified -> compare as sets)."""
    coords, ps, rs = [], [], []
    for b in range(probs.shape[0]):
        mask = probs[b] > thresh
        c = torch.nonzero(mask)
        if c.shape[0] == 0:
            continue
        p = probs[b][mask]
        r = residuals[b][:, mask].permute(1, 0) if residuals is not None else None
        if p.shape[0] > max_points:
            p, top = torch.topk(p, max_points, largest=True, sorted=False)
            c = c[top]
            r = r[top] if r is not None else None
        coords.append(torch.cat([torch.full_like(c[:, :1], b), c], dim=1))
        ps.append(p)
        if r is not None:
            rs.append(r)
    if not coords:
        return None, None, None
    return torch.cat(coords), torch.cat(ps), (torch.cat(rs) if rs else None)


def occ_coords2absxyz(occ_coords, geo: occ_masks.OccGeometry, rot_z=None):
    """Voxel centre in cylinder coordinates -> (undo the augmentation rotation) -> Cartesian (:131-146)."""
    ox, oy, oz = geo.point_cloud_range[0:3]
    vx, vy, vz = geo.voxel_size
    cx = ox + (occ_coords[..., 3] + 0.5) * vx
    cy = oy + (occ_coords[..., 2] + 0.5) * vy
    cz = oz + (occ_coords[..., 1] + 0.5) * vz
    if rot_z is not None:
        cy = cy - rot_z[occ_coords[..., 0]]
    return occ_masks.cylinder_uvd2absxyz(cx, cy, cz)


def trans_voxel_grid(xyz, b_inds, voxel_size, grid_size, point_cloud_range):
    """:78-88 — det-grid coordinates (b, z, y, x), clamped (not dropped)."""
    dev = xyz.device
    rng = torch.tensor(point_cloud_range, dtype=torch.float32, device=dev)
    vs = torch.tensor(voxel_size, dtype=torch.float32, device=dev)
    nx, ny, nz = grid_size
    c = torch.div(xyz - rng[0:3].unsqueeze(0), vs.unsqueeze(0))
    cx = torch.clamp(torch.floor(c[..., 0]), min=0, max=nx - 1).to(torch.int64)
    cy = torch.clamp(torch.floor(c[..., 1]), min=0, max=ny - 1).to(torch.int64)
    cz = torch.clamp(torch.floor(c[..., 2]), min=0, max=nz - 1).to(torch.int64)
    return torch.stack([b_inds, cz, cy, cx], dim=-1)


def pass_occ_vox(probs, residuals, det_voxels, det_voxel_num_points, det_voxel_coords, geo, det_voxel_size, det_grid_size,
                 det_range, thresh=0.3, max_points=2048, rot_z=None, inten=0.0):
    """PassOccVox.forward for COORD_TYPE cylinder, REG True, CODE_NUM_DIM 2, 4-feature points.
    Returns dict(voxels [M',Pmax,6], voxel_num_points [M'], voxel_coords [M',4] i64, occ_pnts, occ_coords, ...)."""
    occ_coords, occ_probs, occ_res = filter_occ_points(probs, residuals, thresh, max_points)
    M, P, C = det_voxels.shape
    mask = det_voxel_num_points.int().unsqueeze(1) > torch.arange(P, dtype=torch.int, device=det_voxels.device).view(1, -1)
    inds = mask.nonzero()
    gt_points = det_voxels[inds[:, 0], inds[:, 1], :]
    gt_coords = det_voxel_coords[inds[:, 0], :].to(torch.int64)
    zeros = torch.zeros_like(gt_points[..., :1])
    gt_points = torch.cat([gt_points, zeros, zeros], dim=-1)                 # two zero code channels (:186-189)
    if occ_coords is None:
        return None
    xyz = occ_coords2absxyz(occ_coords, geo, rot_z)
    if occ_res is not None:
        xyz = xyz + occ_res
    det_coords = trans_voxel_grid(xyz, occ_coords[..., 0], det_voxel_size, det_grid_size, det_range)
    ones = torch.ones_like(xyz[..., :1])
    occ_pnts = torch.cat([xyz, ones * inten, occ_probs.unsqueeze(-1), ones], dim=-1)     # [x,y,z,inten,prob,1]
    points = torch.cat([gt_points, occ_pnts], dim=0)
    coords = torch.cat([gt_coords, det_coords], dim=0)
    vox_coords, inverse, counts = torch.unique(coords, dim=0, sorted=True, return_inverse=True, return_counts=True)
    return {"occ_coords": occ_coords, "occ_probs": occ_probs, "occ_xyz": xyz, "occ_det_coords": det_coords, "occ_pnts": occ_pnts,
            "points": points, "coords": coords, "voxel_coords": vox_coords, "voxel_num_points": counts, "inverse": inverse}


def occ_vfe(voxels, voxel_num_points, num_raw_features=4):
    """OccVFE.forward (occ_vfe.py:24-55): raw / occupancy slots split by the code channel, mean + max."""
    P = voxels.shape[1]
    mask = torch.arange(P, dtype=torch.int, device=voxels.device).view(1, -1) < voxel_num_points.int().unsqueeze(1)
    raw_mask = (voxels[:, :, -1] < 0.05) & mask
    occ_mask = (voxels[:, :, -1] >= 0.05) & mask
    raw_n = raw_mask.sum(dim=1).view(-1, 1)
    occ_n = occ_mask.sum(dim=1).view(-1, 1)
    occ_only = (occ_n > 0.5) & (raw_n < 0.5)
    raw_nf = torch.clamp_min(raw_n, min=1.0).type_as(voxels)
    occ_nf = torch.clamp_min(occ_n, min=1.0).type_as(voxels)
    f_raw = (raw_mask.unsqueeze(-1) * voxels[:, :, :num_raw_features]).sum(dim=1) / raw_nf
    f_occ = (occ_mask.unsqueeze(-1) * voxels[:, :, :num_raw_features]).sum(dim=1) / occ_nf
    feats = f_raw + occ_only * f_occ
    occ_max = voxels[:, :, num_raw_features:].max(dim=1)[0]
    return torch.cat([feats, occ_max], dim=-1), occ_max
