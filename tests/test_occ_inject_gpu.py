"""GPU parity of btc_occ_select / btc_occ_vfe / the PassOccVox composition against
oracle/occ_inject.py run on the same device (strict) — the oracle itself is pinned to the reference on CPU."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _case(seed, dense, with_rot):
    import test_occ_inject_cpu as T
    case, geo = T.make_case(seed, dense=dense, with_rot=with_rot)
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}, geo


@pytest.mark.parametrize("seed,with_rot,max_points,dense", [(1, True, 40000, 0.004), (2, False, 40000, 0.004), (3, True, 2048, 0.05)])
def test_pass_occ_vox_matches_oracle_on_device(cuda, oracle, seed, with_rot, max_points, dense):
    from btcdet_b200 import ops, synthetic as S
    from oracle import occ_inject
    case, geo = _case(seed, dense, with_rot)
    det_grid = [1408, 1600, 40]
    ref = occ_inject.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                  case["det_voxel_coords"], geo, S.DET_VOXEL_SIZE, det_grid, S.KITTI_RANGE, thresh=0.3,
                                  max_points=max_points, rot_z=case["rot_z"])
    vox, cnt, vc, sel = ops.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                         case["det_voxel_coords"], case["batch"], 0.3, max_points, geo.voxel_size,
                                         geo.point_cloud_range[:3], S.DET_VOXEL_SIZE, S.KITTI_RANGE, det_grid,
                                         rot_z=case["rot_z"])
    if max_points >= 40000:    # no top-k: ordered compaction == torch.nonzero order, everything bit-exact
        assert torch.equal(sel["occ_coords"].long(), ref["occ_coords"])
        assert torch.equal(sel["occ_probs"], ref["occ_probs"])
        assert torch.equal(sel["occ_xyz"], ref["occ_xyz"])
        assert torch.equal(sel["det_coords"].long(), ref["occ_det_coords"])
        assert torch.equal(sel["occ_points"], ref["occ_pnts"])
    else:                      # top-k order is unspecified (sorted=False): same set of pseudo points
        assert sorted(map(tuple, sel["occ_points"].tolist())) == sorted(map(tuple, ref["occ_pnts"].tolist()))
    assert torch.equal(vc, ref["voxel_coords"]) and torch.equal(cnt, ref["voxel_num_points"])
    # voxel contents: same multiset of rows per voxel (the reference's slot order comes from an unstable sort)
    order = torch.argsort(ref["inverse"], stable=True)
    inv_sorted = ref["inverse"][order]
    start = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), ref["voxel_num_points"][:-1]]), 0)
    want = torch.zeros_like(vox)
    want[inv_sorted, torch.arange(len(order), device="cuda") - start[inv_sorted]] = ref["points"][order]
    if max_points >= 40000:
        assert torch.equal(vox, want)          # stable slots: identical to a stable sort of the reference formulation
    feats, occ = ops.occ_vfe(vox, cnt, 4)
    rf, ro = occ_inject.occ_vfe(vox, cnt, 4)
    torch.testing.assert_close(feats, rf, rtol=1e-6, atol=1e-6)
    assert torch.equal(occ, ro)


def test_occ_select_nothing_above_threshold(cuda):
    from btcdet_b200 import ops, synthetic as S
    from oracle.occ_masks import OccGeometry
    geo = OccGeometry()
    probs = torch.zeros(2, 9, 157, 209, device="cuda")
    assert ops.occ_select(probs, None, 0.3, 2048, geo.voxel_size, geo.point_cloud_range[:3], S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                          [1408, 1600, 40]) is None
