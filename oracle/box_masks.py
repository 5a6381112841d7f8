"""Restatement of the reference's box-driven occupancy targets (SURVEY §8 rows a9-a11 and the rest of a12).

TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  torch ops, device-agnostic; pinned bit-for-bit against the
reference's own `OccTargets3D.create_voxel_res_label` executed on CPU (tests/test_box_masks_cpu.py).

Reference code followed:
  torch_points_and_sym_in_box_3d_batch / torch_points_in_box_3d_label_mirr_points   btcdet/utils/point_box_utils.py:70-97,252-306
  torch_points_in_box_2d_mask / torch_points_in_box_3d_label / rotatez               :332-365, :198-238, :241-250
  get_fore_mirr_voxelwise_mask_res / get_mean_res / get_voxel_center_xyz             btcdet/models/occ_pnt/occ_training_targets/occ_targets_3d.py:122-171
  forebox label loop                                                                 occ_targets_3d.py:70-86
  prepare_cls_loss_map / prepare_reg_loss_map                                        occ_targets_template.py:330-401

Note on reproducibility: the box-frame coordinates go through `torch.inverse` (LU) of a 4x4 transform, so points within
rounding distance of a box face may be classified differently by LAPACK (CPU), by the CUDA batched LU and by an analytic
inverse; `box_frame()` therefore exposes the margins so that tests can exclude the ambiguous points.
"""
import numpy as np
import torch

from . import occ_masks


def yaw_rotation(yaw):
    c, s = torch.cos(yaw), torch.sin(yaw)
    one, zero = torch.ones_like(yaw), torch.zeros_like(yaw)
    return torch.stack([torch.stack([c, -1.0 * s, zero], dim=-1), torch.stack([s, c, zero], dim=-1),
                        torch.stack([zero, zero, one], dim=-1)], dim=-2)


def transform4(rotation, translation):
    t = torch.cat([rotation, translation.unsqueeze(-1)], dim=-1)
    last = torch.cat([torch.zeros_like(translation), torch.ones_like(translation[..., 0:1])], dim=-1)
    return torch.cat([t, last.unsqueeze(-2)], dim=-2)


def box_frame(points, boxes):
    """q[n, m, :] = coordinates of point n in the frame of box m (through torch.inverse, as the reference)."""
    rot = yaw_rotation(boxes[:, 6])
    inv = torch.inverse(transform4(rot, boxes[:, :3]))
    q = torch.einsum("nj,mij->nmi", points, inv[:, :3, :3]) + inv[:, :3, 3]
    return q, rot


def points_in_boxes_mirror(points, boxes, mirr_flag):
    """point_box_utils.py:252-306 — per-point label (max box label over containing boxes) and mirrored points of the
    (point, box) pairs whose box has the mirror flag, in torch.nonzero order."""
    n, m = points.shape[0], boxes.shape[0]
    if m == 0:
        return torch.zeros(n, dtype=torch.int8, device=points.device), None
    dim = boxes[:, 3:6]
    label = boxes[:, 7].to(torch.int8)
    q, rot = box_frame(points, boxes)
    inside = torch.prod((q <= dim * 0.5) & (q >= -dim * 0.5), dim=-1, dtype=torch.int8)
    mirr_pairs = torch.nonzero(inside * (mirr_flag > 0.5).to(torch.int8).unsqueeze(0))
    qm = q.clone()
    qm[:, :, 1] = -qm[:, :, 1]
    back = torch.einsum("nmj,mij->nmi", qm, rot) + boxes[:, :3]
    mirror_points = back[mirr_pairs[:, 0], mirr_pairs[:, 1], :]
    point_label = torch.max(inside * label.unsqueeze(0), dim=1)[0]
    return point_label, mirror_points


def points_in_boxes_label(points, boxes):
    """torch_points_in_box_3d_label (:198-238) for a single scene: max label over containing boxes."""
    if boxes.shape[0] == 0:
        return torch.zeros(points.shape[0], dtype=torch.int8, device=points.device)
    q, _ = box_frame(points, boxes)
    dim = boxes[:, 3:6]
    inside = torch.prod((q <= dim * 0.5) & (q >= -dim * 0.5), dim=-1, dtype=torch.int8)
    return torch.max(inside * boxes[:, 7].to(torch.int8).unsqueeze(0), dim=1)[0]


def points_in_boxes_2d(points, boxes):
    """torch_points_in_box_2d_mask (:332-365)."""
    if boxes.shape[0] == 0:
        return torch.zeros(points.shape[0], dtype=torch.bool, device=points.device)
    yaw = boxes[:, 6]
    c, s = torch.cos(yaw), torch.sin(yaw)
    rot = torch.stack([torch.stack([c, -1.0 * s], dim=-1), torch.stack([s, c], dim=-1)], dim=-2)
    t = torch.cat([rot, boxes[:, :2].unsqueeze(-1)], dim=-1)
    last = torch.cat([torch.zeros_like(boxes[:, :2]), torch.ones_like(boxes[:, 0:1])], dim=-1)
    inv = torch.inverse(torch.cat([t, last.unsqueeze(-2)], dim=-2))
    q = torch.einsum("nj,mij->nmi", points, inv[:, :2, :2]) + inv[:, :2, 2]
    dim = boxes[:, 3:5]
    inside = torch.prod((q <= dim * 0.5) & (q >= -dim * 0.5), dim=-1, dtype=torch.int8) > 0
    return torch.any(inside, dim=-1)


def rotatez(points, zyaw_deg):
    """point_box_utils.py:241-250 (degrees)."""
    yaw = zyaw_deg * np.pi / 180.
    c, s = torch.cos(yaw), torch.sin(yaw)
    if points.shape[-1] == 3:
        rot = torch.stack([torch.stack([c, -1.0 * s, torch.zeros_like(c)]), torch.stack([s, c, torch.zeros_like(c)]),
                           torch.stack([torch.zeros_like(c), torch.zeros_like(c), torch.ones_like(c)])])
    else:
        rot = torch.stack([torch.stack([c, -1.0 * s]), torch.stack([s, c])])
    return torch.matmul(points, rot.transpose(0, 1))


def voxel_center_xyz(coords, geo, rot_z=None):
    """get_voxel_center_xyz (occ_targets_3d.py:133-145), cylinder grid."""
    dev = coords.device
    vs = torch.as_tensor([geo.voxel_size], dtype=torch.float32, device=dev)
    org = torch.as_tensor([geo.point_cloud_range[:3]], dtype=torch.float32, device=dev)
    c = (coords[:, [3, 2, 1]].float() + 0.5) * vs + org
    if rot_z is not None:
        c[..., 1] -= rot_z[coords[:, 0]]
    return occ_masks.cylinder_uvd2absxyz(c[..., 0], c[..., 1], c[..., 2])


def mean_res(feat, coords, bs, geo, rot_z=None):
    """get_mean_res (occ_targets_3d.py:122-130): per-voxel mean of the points minus the voxel centre."""
    nx, ny, nz = geo.grid_size
    out = torch.zeros([bs, 3, nz, ny, nx], dtype=torch.float32, device=feat.device)
    if len(coords) > 0:
        uni, inv, cnt = torch.unique(coords, return_inverse=True, return_counts=True, dim=0)
        mean = torch.zeros([uni.shape[0], 3], dtype=feat.dtype, device=feat.device).scatter_add_(
            0, inv.view(-1, 1).expand(-1, 3), feat[..., :3]) / cnt.float().unsqueeze(1)
        mean = mean - voxel_center_xyz(uni, geo, rot_z)
        out[uni[:, 0], :, uni[:, 1], uni[:, 2], uni[:, 3]] = mean
    return out


def cyl_coords_inrange(points, b_inds, geo, rot_z=None):
    dev = points.device
    nx, ny, nz = geo.grid_size
    cyl = occ_masks.cartesian_cylinder_coords(points)
    if rot_z is not None:
        cyl[..., 1] += rot_z[b_inds]
    org = torch.as_tensor([geo.point_cloud_range[:3]], dtype=torch.float32, device=dev)
    pmax = torch.as_tensor([geo.point_cloud_range[3:]], dtype=torch.float32, device=dev)
    vs = torch.as_tensor([geo.voxel_size], dtype=torch.float32, device=dev)
    mg = torch.as_tensor([[nx - 1, ny - 1, nz - 1]], dtype=torch.int64, device=dev)
    c, inds = occ_masks.point2coords_inrange(cyl, org, pmax, mg, vs)
    coords = torch.cat([b_inds[inds].unsqueeze(-1), torch.stack([c[:, 2], c[:, 1], c[:, 0]], dim=-1)], dim=-1)
    return coords, inds


def box_targets(valid_coords, valid_feats, gt_boxes, gt_boxes_num, mirr_flag, bs, geo, rot_z=None, num_class=1,
                box_weight=0.2, bm_points=None):
    """fore / mirror masks + residual matrices (a9), best-match template points (a10, bm_points [P,4] = b,x,y,z) and
    the forebox label (a11) for every scene of the batch."""
    dev = valid_feats.device
    nx, ny, nz = geo.grid_size
    if num_class == 1:
        gt_boxes = torch.cat([gt_boxes[..., :-1], (gt_boxes[..., -1:] > 1e-2).to(torch.float32)], dim=-1)
    pts = valid_feats[..., :3]
    label = torch.zeros(pts.shape[0], dtype=torch.int8, device=dev)
    mirr_pts, mirr_b = [], []
    for i in range(bs):
        sel = torch.nonzero(valid_coords[:, 0] == i)[:, 0]
        if sel.numel() == 0:
            continue
        lab, mp = points_in_boxes_mirror(pts[sel], gt_boxes[i, :gt_boxes_num[i]], mirr_flag[i, :gt_boxes_num[i]])
        label[sel] = lab
        if mp is not None:
            mirr_pts.append(mp)
            mirr_b.append(torch.full((mp.shape[0],), i, dtype=torch.int64, device=dev))
    fore_inds = label > 0
    fore_coords = valid_coords[fore_inds]
    fore_mask = torch.zeros([bs, nz, ny, nx], dtype=torch.uint8, device=dev)
    fore_mask[fore_coords[:, 0], fore_coords[:, 1], fore_coords[:, 2], fore_coords[:, 3]] = 1
    fore_res = mean_res(valid_feats[fore_inds], fore_coords, bs, geo, rot_z)
    mirr_mask = torch.zeros_like(fore_mask)
    mirr_res = torch.zeros([bs, 3, nz, ny, nx], dtype=torch.float32, device=dev)
    if mirr_pts:
        mp, mb = torch.cat(mirr_pts), torch.cat(mirr_b)
        mc, inds = cyl_coords_inrange(mp, mb, geo, rot_z)
        mirr_res = mean_res(mp[inds], mc, bs, geo, rot_z)
        mirr_mask[mc[:, 0], mc[:, 1], mc[:, 2], mc[:, 3]] = 1
    # best-match template points (get_bm_voxelwise_mask_res, occ_targets_3d.py:95-119)
    bm_mask = torch.zeros_like(fore_mask)
    bm_res = torch.zeros([bs, 3, nz, ny, nx], dtype=torch.float32, device=dev)
    if bm_points is not None and len(bm_points) > 0:
        bb, bp = bm_points[:, 0].to(torch.int64), bm_points[:, 1:4]
        lab = torch.zeros(bp.shape[0], dtype=torch.int8, device=dev)
        for i in range(bs):
            sel = torch.nonzero(bb == i)[:, 0]
            if sel.numel() > 0:
                lab[sel] = points_in_boxes_label(bp[sel], gt_boxes[i, :gt_boxes_num[i]])
        keep = torch.nonzero(lab)[:, 0]
        bb, bp = bb[keep], bp[keep]
        bc, inds = cyl_coords_inrange(bp, bb, geo, rot_z)
        bm_res = mean_res(bp[inds], bc, bs, geo, rot_z)
        bm_mask[bc[:, 0], bc[:, 1], bc[:, 2], bc[:, 3]] = 1
    # forebox label (occ_targets_3d.py:70-86)
    forebox = None
    if box_weight != 1.0:
        forebox = torch.zeros([bs, nz, ny, nx], dtype=torch.int8, device=dev)
        centers = all_voxel_centers(geo, dev)                                  # [nz, ny, nx, 3]
        centers2d = torch.mean(centers[:, :, :, :2], dim=0).view(-1, 2)
        for i in range(bs):
            boxes = gt_boxes[i, :gt_boxes_num[i]]
            c2 = rotatez(centers2d, rot_z[i]) if rot_z is not None else centers2d
            hit = points_in_boxes_2d(c2, boxes).view(ny, nx).nonzero()
            if hit.shape[0] > 0:
                cf = centers[:, hit[:, 0], hit[:, 1], :].reshape(-1, 3)
                if rot_z is not None:
                    cf = rotatez(cf, rot_z[i])
                forebox[i, :, hit[:, 0], hit[:, 1]] = points_in_boxes_label(cf, boxes).view(nz, -1)
    return {"fore_voxelwise_mask": fore_mask, "fore_res_mtrx": fore_res, "mirr_fore_voxelwise_mask": mirr_mask,
            "mirr_res_mtrx": mirr_res, "bm_voxelwise_mask": bm_mask, "bm_res_mtrx": bm_res, "forebox_label": forebox,
            "point_label": label}


def all_voxel_centers(geo, device):
    """detector3d_template.py:52-63: Cartesian centres of every cylinder voxel, [nz, ny, nx, 3]."""
    nx, ny, nz = geo.grid_size
    vs = torch.tensor([geo.voxel_size[2], geo.voxel_size[1], geo.voxel_size[0]], device=device)
    org = torch.tensor([geo.point_cloud_range[2], geo.point_cloud_range[1], geo.point_cloud_range[0]], device=device)
    z, y, x = torch.meshgrid(torch.arange(nz, device=device), torch.arange(ny, device=device),
                             torch.arange(nx, device=device), indexing="ij")
    c = (0.5 + torch.stack([z, y, x], dim=0).to(torch.float32)) * vs.view(3, 1, 1, 1) + org.view(3, 1, 1, 1)
    return occ_masks.cylinder_uvd2absxyz(c[2], c[1], c[0])


def loss_maps(occ, box, weights=None):
    """prepare_cls_loss_map + prepare_reg_loss_map (occ_targets_template.py:330-401) without dropout."""
    w = weights or {"fore_cls": 1.0, "mirr_cls": 1.0, "bm_cls": 1.0, "neg_cls": 1.0, "fore_res": 0.1, "mirr_res": 0.0,
                    "bm_res": 0.0, "box_weight": 0.2}
    vm = occ["voxelwise_mask"]
    g = occ["general_cls_loss_mask"]
    fore = box["fore_voxelwise_mask"] & g
    mirr_m = box["mirr_fore_voxelwise_mask"] * (1 - vm)
    mirr = mirr_m & g
    bm_m = box["bm_voxelwise_mask"] * (1 - vm) * (1 - mirr_m)
    bm = bm_m & g
    pos = fore | mirr | bm
    neg = g & (1 - pos)
    fl = fore.to(torch.float32) * w["fore_cls"] + mirr.to(torch.float32) * w["mirr_cls"] + bm.to(torch.float32) * w["bm_cls"] \
        + neg.to(torch.float32) * w["neg_cls"]
    if box["forebox_label"] is not None:
        fl = fl + (neg & (box["forebox_label"] > 1e-3)).to(torch.float32) * (w["box_weight"] - w["neg_cls"])
    reg_f = fore.to(torch.float32) * w["fore_res"] + mirr.to(torch.float32) * w["mirr_res"] + bm.to(torch.float32) * w["bm_res"]
    reg_m = (reg_f > 0).to(torch.uint8)
    mirr_res = box["mirr_res_mtrx"] * (1 - vm).unsqueeze(1)
    bm_res = box["bm_res_mtrx"] * (1 - vm).unsqueeze(1) * (1 - mirr_m).unsqueeze(1)
    res = box["fore_res_mtrx"] * reg_m.unsqueeze(1) + mirr_res * reg_m.unsqueeze(1) + bm_res * reg_m.unsqueeze(1)
    return {"occ_fore_cls_mask": fore, "occ_mirr_cls_mask": mirr, "occ_bm_cls_mask": bm, "pos_mask": pos, "neg_mask": neg,
            "general_cls_loss_mask_float": fl, "general_reg_loss_mask": reg_m, "general_reg_loss_mask_float": reg_f,
            "res_mtrx": res}
