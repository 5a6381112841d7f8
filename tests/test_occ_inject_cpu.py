"""Pins oracle/occ_inject.py against the reference's own PassOccVox.forward and
OccVFE.forward executed on CPU in this container (tests/golden/ref_loader.py); skipped without a reference."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")

MODEL_CFG = {"PARAMS": {"OCC_THRESH": 0.3, "EVAL_OCC_THRESH": 0.57, "MAX_NUM_OCC_PNTS": 2048, "EVAL_MAX_NUM_OCC_PNTS": 40000,
                        "REG": True}, "OCC_PNT_UPDATE": {"PASS_GRAD": False}}


def make_case(seed, batch=2, dense=0.004, with_rot=True):
    """Synthetic occupancy-head outputs + det voxels in the reference's batch_dict layout."""
    from btcdet_b200 import synthetic as S
    from oracle import oracle as O
    from oracle.occ_masks import OccGeometry
    geo = OccGeometry()
    nx, ny, nz = geo.grid_size
    g = torch.Generator().manual_seed(seed)
    probs = torch.rand(batch, nz, ny, nx, generator=g)
    keep = torch.rand(batch, nz, ny, nx, generator=g) < dense
    probs = torch.where(keep, 0.3 + 0.7 * probs, 0.25 * probs)
    res = (torch.rand(batch, 3, nz, ny, nx, generator=g) - 0.5) * 0.3
    scenes = [S.lidar_like(8000, seed=seed * 10 + b) for b in range(batch)]
    v, c, n = O.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    case = {"probs": probs, "res": res, "det_voxels": torch.from_numpy(v), "det_voxel_coords": torch.from_numpy(c).float(),
            "det_voxel_num_points": torch.from_numpy(n).float(), "batch": batch,
            "rot_z": torch.tensor([5.0, -12.5][:batch]) if with_rot else None,
            "points": torch.from_numpy(np.concatenate([np.pad(s, ((0, 0), (1, 0)), constant_values=b) for b, s in enumerate(scenes)]))}
    return case, geo


def run_reference(case, geo, is_train=True):
    import make_occ_golden
    from btcdet_b200 import synthetic as S
    mods = ref_loader.load_reference_modules()
    with ref_loader.cuda_as_cpu():
        data_cfg = ref_loader.Cfg.wrap(make_occ_golden.data_cfg(geo))
        mod = mods["pass_occ_vox"].PassOccVox(ref_loader.Cfg.wrap(MODEL_CFG), data_cfg, S.KITTI_RANGE, geo.voxel_size,
                                              geo.grid_size, S.DET_VOXEL_SIZE, [1408, 1600, 40], "train",
                                              {"all_voxel_centers": torch.zeros(1)})
        bd = {"voxels": torch.zeros(3, 12, 4), "voxel_num_points": torch.ones(3), "voxel_coords": torch.zeros(3, 4),
              "batch_size": case["batch"], "batch_pred_occ_prob": case["probs"].clone(), "pred_sem_residuals": case["res"].clone(),
              "points": case["points"], "det_voxels": case["det_voxels"].clone(), "det_voxel_coords": case["det_voxel_coords"],
              "det_voxel_num_points": case["det_voxel_num_points"], "is_train": is_train,
              "use_occ_prob": [True] * case["batch"]}
        if case["rot_z"] is not None:
            bd["rot_z"] = case["rot_z"]
        out = mod(bd)
        vfe = mods["occ_vfe"].OccVFE(ref_loader.Cfg(), 6, ref_loader.Cfg.wrap(
            {"POINT_FEATURE_ENCODING": {"used_feature_list": ["x", "y", "z", "intensity"]}}), maxprob=True)
        out = vfe(out)
    return out


def _rows_by_voxel(voxels, counts):
    return [sorted(map(tuple, voxels[i, :int(counts[i])].tolist())) for i in range(voxels.shape[0])]


@pytest.mark.parametrize("seed,with_rot,is_train,dense", [(1, True, False, 0.004), (2, False, False, 0.004), (3, True, True, 0.05)])
def test_oracle_equals_reference_pass_occ_vox_and_occ_vfe(seed, with_rot, is_train, dense):
    from btcdet_b200 import synthetic as S
    from oracle import occ_inject
    case, geo = make_case(seed, dense=dense, with_rot=with_rot)
    ref = run_reference(case, geo, is_train=is_train)
    got = occ_inject.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                  case["det_voxel_coords"], geo, S.DET_VOXEL_SIZE, [1408, 1600, 40], S.KITTI_RANGE, thresh=0.3,
                                  max_points=2048 if is_train else 40000, rot_z=case["rot_z"])
    if is_train:   # top-k (sorted=False) picks the same SET; its order is unspecified in the reference
        assert (case["probs"][0] > 0.3).sum() > 2048 and got["occ_coords"].shape[0] == 2 * 2048
        assert sorted(map(tuple, got["occ_pnts"].tolist())) == sorted(map(tuple, torch.cat(
            [ref["occ_pnts"][:, :3], torch.zeros(len(ref["occ_pnts"]), 1), ref["occ_pnts"][:, 3:4],
             torch.ones(len(ref["occ_pnts"]), 1)], 1).tolist()))
    else:
        assert torch.equal(got["occ_xyz"], ref["added_occ_xyz"])
        assert torch.equal(got["occ_coords"][:, 0], ref["added_occ_b_ind"])
    assert torch.equal(got["voxel_coords"], ref["voxel_coords"])
    assert torch.equal(got["voxel_num_points"], ref["voxel_num_points"])
    # slot order inside a voxel follows an (unstable) sort in the reference: compare each voxel's rows as a set
    pmax = int(got["voxel_num_points"].max())
    vox = torch.zeros(got["voxel_coords"].shape[0], pmax, 6)
    order = torch.argsort(got["inverse"], stable=True)
    inv_sorted = got["inverse"][order]
    start = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long), got["voxel_num_points"][:-1]]), 0)
    vox[inv_sorted, torch.arange(len(order)) - start[inv_sorted]] = got["points"][order]
    assert _rows_by_voxel(vox, got["voxel_num_points"]) == _rows_by_voxel(ref["voxels"], ref["voxel_num_points"])
    feats, occ_max = occ_inject.occ_vfe(vox, got["voxel_num_points"], 4)
    torch.testing.assert_close(feats, ref["voxel_features"], rtol=1e-6, atol=1e-6)
    assert torch.equal(occ_max, ref["occ_voxel_features"])
