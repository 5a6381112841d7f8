"""The size-independent property checkers (tests/properties.py) accept the oracle's voxel groups / neighbour tables and
reject corrupted ones — so that the GPU suite can rely on them at BASELINE config-5 sizes where the oracle is too slow."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import properties as P  # noqa: E402


@pytest.fixture(scope="module")
def scene(oracle):
    from btcdet_b200 import synthetic as S
    scenes = [S.lidar_like(6000, seed=5), S.lidar_like(4000, seed=6)]
    vox, coords, num = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    return scenes, vox, coords, num


def test_voxelization_checker(scene, oracle):
    from btcdet_b200 import ops, synthetic as S
    scenes, vox, coords, num = scene
    pts, offs = S.batch_points(scenes)
    grid = ops.voxel_grid_size(S.DET_VOXEL_SIZE, S.KITTI_RANGE)
    counts = np.bincount(coords[:, 0], minlength=2)
    nv = torch.tensor([counts[0], counts[1], coords.shape[0]], dtype=torch.int32)
    args = (torch.from_numpy(pts), torch.from_numpy(offs), torch.from_numpy(vox), torch.from_numpy(coords),
            torch.from_numpy(num.astype(np.int32)), nv, S.DET_VOXEL_SIZE, S.KITTI_RANGE, grid, 5, 16000)
    assert P.check_voxelization(*args) == coords.shape[0]
    bad = torch.from_numpy(coords).clone()
    bad[7, 3] += 1                                    # a voxel row claims the neighbouring cell
    with pytest.raises(AssertionError):
        P.check_voxelization(*(args[:3] + (bad,) + args[4:]))
    badv = torch.from_numpy(vox).clone()
    badv[3, 4, 0] = 1.0                               # garbage in a padding slot (voxel 3 holds < 5 points?)
    if int(num[3]) < 5:
        with pytest.raises(AssertionError):
            P.check_voxelization(*(args[:2] + (badv,) + args[3:]))


def test_subm_table_checker(scene, oracle):
    _, _, coords, _ = scene
    shape = [41, 1600, 1408]
    outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 2, shape, 3, subm=True)
    nbr_out, _ = oracle.pairs_to_tables(pairs, pair_num, coords.shape[0], coords.shape[0])
    # the oracle's pair list follows spconv's offset convention: entry [o][k] = input feeding o through offset k
    c, t = torch.from_numpy(coords), torch.from_numpy(nbr_out)
    n_pairs = P.check_subm_table(c, t, shape, [3, 3, 3])
    assert n_pairs == int(pair_num.sum())
    i, k = np.argwhere(nbr_out >= 0)[11]
    for mutate in ("drop", "wrong", "asym"):
        bad = t.clone()
        if mutate == "drop":
            j = int(bad[i, k])
            bad[i, k] = -1
            bad[j, 26 - k] = -1                        # symmetric removal: only the completeness count can catch it
        elif mutate == "wrong":
            bad[i, k] = (int(bad[i, k]) + 1) % coords.shape[0]
        else:
            bad[i, k] = -1
        with pytest.raises(AssertionError):
            P.check_subm_table(c, bad, shape, [3, 3, 3])


@pytest.mark.parametrize("ksize,stride,pad", [([3, 3, 3], [2, 2, 2], [1, 1, 1]), ([3, 3, 3], [2, 2, 2], [0, 1, 1]),
                                               ([3, 1, 1], [2, 1, 1], [0, 0, 0])])
def test_conv_table_checker(scene, oracle, ksize, stride, pad):
    _, _, coords, _ = scene
    shape = [41, 1600, 1408]
    outids, pairs, pair_num, out_shape = oracle.get_indice_pairs(coords, 2, shape, ksize, stride, pad)
    nbr_out, nbr_in = oracle.pairs_to_tables(pairs, pair_num, coords.shape[0], outids.shape[0])
    ci, co = torch.from_numpy(coords), torch.from_numpy(outids)
    to, ti = torch.from_numpy(nbr_out), torch.from_numpy(nbr_in)
    n = P.check_conv_tables(ci, co, to, ti, shape, list(out_shape), ksize, stride, pad)
    assert n == int(pair_num.sum())
    o, k = np.argwhere(nbr_out >= 0)[5]
    bad = to.clone()
    bad[o, k] = -1
    with pytest.raises(AssertionError):
        P.check_conv_tables(ci, co, bad, ti, shape, list(out_shape), ksize, stride, pad)
    swapped = co.clone()
    swapped[[0, 1]] = swapped[[1, 0]]                  # outputs out of flat-key order
    with pytest.raises(AssertionError):
        P.check_conv_tables(ci, swapped, to, ti, shape, list(out_shape), ksize, stride, pad)
