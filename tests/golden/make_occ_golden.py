"""Generates tests/golden/occ_masks.npz by running the REFERENCE's own OccTargets3D code
(/root/reference/btcdet/models/occ_pnt/occ_training_targets/occ_targets_3d.py, unchanged, executed on CPU through
tests/golden/ref_loader.py) on seeded synthetic scenes.  The fixture pins oracle/occ_masks.py (and through it the
CUDA kernels) to the reference for SURVEY §8 rows a5-a8/a12 where the GPU box has no reference checkout.
Run from the repo root in the build container:  python tests/golden/make_occ_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_loader  # noqa: E402
from btcdet_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle.occ_masks import OccGeometry  # noqa: E402

MODEL_OCC_CFG = {
    "PARAMS": {"OCC_THRESH": 0.3, "REG": True},
    "TARGETS": {"NAME": "OccTargets3D", "TMPLT": True},
    "OCC_DENSE_HEAD": {"LOSS_CONFIG": {"LOSS_WEIGHTS": {
        'occ_fore_cls_weight': 1.0, 'occ_mirr_cls_weight': 1.0, 'occ_bm_cls_weight': 1.0, 'occ_neg_cls_weight': 1.0,
        'occ_fore_res_weight': 0.1, 'occ_mirr_res_weight': 0.0, 'occ_bm_res_weight': 0.0, 'res_beta': 0.025,
        'cls_alpha': 0.5, 'fore_dropout_cls_weight': 1.0, 'fore_dropout_reg_weight': 1.0}}},
}


def data_cfg(geo):
    return {"POINT_CLOUD_RANGE": geo.det_point_cloud_range,
            "OCC": {"VOXEL_SIZE": geo.voxel_size, "DIST_KERN": geo.dist_kern, "HALF_X": geo.half_x,
                    "EMPT_SUR_THRESH": geo.empt_sur_thresh, "POINT_CLOUD_RANGE": geo.point_cloud_range,
                    "SUPPORT_SPHERE_RANGE": geo.support_sphere_range, "BOX_WEIGHT": 0.2, "RES_NUM_DIM": 3,
                    "CODE_NUM_DIM": 2, "INTEN": 0.0, "DROPOUT_RATE": 0.0, "REAL_DROP": False, "COORD_TYPE": "cylinder",
                    "USE_ABSXYZ": True, "NOLOC": False, "MAX_VFE": True, "USEOCC_PERCENTAGE": 1.1}}


def make_inputs(seeds, n_points=6000, with_rot=False):
    """Synthetic batch in the reference's collate layout: cylindrical occ voxels (data_processor.py:105-145)."""
    geo = OccGeometry()
    gen = O.VoxelGeneratorV2(geo.voxel_size, geo.point_cloud_range, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"])
    vox, coords, nums, boxes = [], [], [], []
    for b, seed in enumerate(seeds):
        pts, bx = S.lidar_like(n_points, seed=seed, return_boxes=True)
        rho = np.linalg.norm(pts[:, :2], axis=1)
        phi = np.arctan2(-pts[:, 1], pts[:, 0]) * 180. / np.pi           # coords_utils.py:282-292 (numpy, dataset side)
        cyl = np.stack([rho, phi, pts[:, 2], pts[:, 3]], axis=-1).astype(np.float32)
        r = gen.generate(cyl)
        vox.append(r["voxels"])
        coords.append(np.pad(r["coordinates"], ((0, 0), (1, 0)), constant_values=b))
        nums.append(r["num_points_per_voxel"])
        boxes.append(bx[:12])
    gt = np.stack(boxes).astype(np.float32)
    out = {"voxels": np.concatenate(vox), "voxel_coords": np.concatenate(coords).astype(np.float32),
           "voxel_num_points": np.concatenate(nums).astype(np.float32), "gt_boxes": gt,
           "gt_boxes_num": [gt.shape[1]] * len(seeds), "box_mirr_flag": np.ones(gt.shape[:2], np.float32),
           "batch_size": len(seeds)}
    if with_rot:
        out["rot_z"] = np.array([7.5, -11.25][:len(seeds)], np.float32)
    return out, geo


def run_reference(inp, geo, device="cpu", keep_all=False):
    """device="cuda": the reference's code runs unchanged on the GPU (O3) — same CUDA libm, ATen kernels and
    torch.inverse (batched LU) as a real BtcDet run."""
    mods = ref_loader.load_reference_modules(device)
    with ref_loader.cuda_as_cpu():
        nx, ny, nz = geo.grid_size
        vs = torch.tensor(geo.voxel_size, dtype=torch.float32, device=device)
        centers = mods["coords_utils"].get_all_voxel_centers_zyx(1, torch.tensor([nx, ny, nz], dtype=torch.int32, device=device),
                                                                 geo.point_cloud_range[:3], vs)[0]
        centers = mods["coords_utils"].uvd2absxyz(centers[2], centers[1], centers[0], "cylinder", dim=-1)
        vc = {"all_voxel_centers": centers, "all_voxel_centers_2d": torch.mean(centers[:, :, :, :2], dim=0).view(-1, 2)}
        tgt = mods["occ_targets_3d"].OccTargets3D(ref_loader.Cfg.wrap(MODEL_OCC_CFG), voxel_size=geo.voxel_size,
                                                  point_cloud_range=geo.point_cloud_range,
                                                  data_cfg=ref_loader.Cfg.wrap(data_cfg(geo)), grid_size=geo.grid_size,
                                                  num_class=1, voxel_centers=vc).to(device)   # as model.cuda() does
        bd = {k: (torch.from_numpy(v).to(device) if isinstance(v, np.ndarray) else v) for k, v in inp.items()}
        bd["is_train"] = True
        out = tgt(bd)
    keep = ["voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "fore_voxelwise_mask", "general_cls_loss_mask", "pos_mask",
            "occ_fore_cls_mask", "occ_mirr_cls_mask", "occ_bm_cls_mask", "general_cls_loss_mask_float", "forebox_label",
            "res_mtrx", "general_reg_loss_mask"]
    if keep_all:
        keep = [k for k, v in out.items() if torch.is_tensor(v)]
    return {k: out[k].cpu().numpy() for k in keep if out.get(k) is not None}


def main():
    inp, geo = make_inputs([3, 4], with_rot=True)
    ref = run_reference(inp, geo)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "occ_masks.npz")
    packed = {"ref_" + k: np.packbits(v.astype(bool)) if v.dtype in (np.uint8, np.bool_) else v.astype(np.float16)
              for k, v in ref.items() if k in ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask",
                                               "fore_voxelwise_mask", "pos_mask")}
    np.savez_compressed(path, shape=np.array(ref["voxelwise_mask"].shape), seeds=np.array([3, 4]), n_points=6000,
                        rot_z=inp["rot_z"], **packed)
    print(path, os.path.getsize(path), {k: int(v.sum()) for k, v in ref.items() if v.dtype in (np.uint8, np.bool_)})


if __name__ == "__main__":
    main()
