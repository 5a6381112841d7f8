"""Pins oracle/box_masks.py against the reference's own OccTargets3D.create_voxel_res_label
executed on CPU in this container (tests/golden/ref_loader.py + make_occ_golden.run_reference); skipped without a
reference checkout.  Same device + same torch ops => bit-exact."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


def bm_points_for(inp, seed):
    """Template ('best match') points: jittered copies of box-interior samples + some outside (b, x, y, z)."""
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(inp["gt_boxes"].shape[0]):
        for bx in inp["gt_boxes"][b][:6]:
            u = rng.uniform(-0.6, 0.6, (40, 3)) * bx[3:6]
            c, s = np.cos(bx[6]), np.sin(bx[6])
            xy = np.stack([c * u[:, 0] - s * u[:, 1], s * u[:, 0] + c * u[:, 1]], 1) + bx[:2]
            rows.append(np.concatenate([np.full((40, 1), b), xy, u[:, 2:3] + bx[2]], 1))
    return np.concatenate(rows).astype(np.float32)


@pytest.mark.parametrize("with_rot,with_bm", [(True, False), (False, True), (True, True)])
def test_box_targets_match_reference(with_rot, with_bm):
    import make_occ_golden as G
    from oracle import box_masks, occ_masks
    inp, geo = G.make_inputs([3, 4], n_points=6000, with_rot=with_rot)
    if with_bm:
        inp["bm_points"] = bm_points_for(inp, 5)
    ref = G.run_reference({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in inp.items()}, geo)
    t = {k: torch.from_numpy(v) for k, v in inp.items() if isinstance(v, np.ndarray)}
    rot = t.get("rot_z")
    occ = occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], 2, geo, rot_z=rot)
    box = box_masks.box_targets(occ["valid_coords"], occ["valid_feats"], t["gt_boxes"], inp["gt_boxes_num"], t["box_mirr_flag"],
                                2, geo, rot_z=rot, bm_points=t.get("bm_points"))
    maps = box_masks.loss_maps(occ, box)
    np.testing.assert_array_equal(box["fore_voxelwise_mask"].numpy(), ref["fore_voxelwise_mask"])
    np.testing.assert_array_equal(box["forebox_label"].numpy(), ref["forebox_label"])
    for k in ("occ_fore_cls_mask", "occ_mirr_cls_mask", "occ_bm_cls_mask", "pos_mask", "general_reg_loss_mask"):
        np.testing.assert_array_equal(maps[k].numpy(), ref[k], err_msg=k)
    np.testing.assert_array_equal(maps["general_cls_loss_mask_float"].numpy(), ref["general_cls_loss_mask_float"])
    np.testing.assert_array_equal(maps["res_mtrx"].numpy(), ref["res_mtrx"])
    assert ref["fore_voxelwise_mask"].sum() > 50 and ref["occ_mirr_cls_mask"].sum() > 10 and ref["forebox_label"].sum() > 100
    if with_bm:
        assert ref["occ_bm_cls_mask"].sum() > 10
    assert np.abs(ref["res_mtrx"]).sum() > 0
