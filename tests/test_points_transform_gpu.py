"""GPU parity of btc_points_to_cylinder and of the GPU-resident input pipeline (N3) against the numpy
restatement (oracle/coords.py, pinned to the reference's functions on CPU).

rho / r / z / extra columns: bit-exact (IEEE mul, add, sqrt in numpy's order).  Angles: CUDA atan2f and numpy's float32
arctan2 (SVML or glibc, host dependent) are different libm implementations, both within a few ulp of the true value —
asserted <= 4 ulp apart; the voxel a point falls into may then differ only for points within that distance of a bin edge,
and the test states how many there are on the seeded scenes (none)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ANGLE_ULP = 4


@pytest.mark.parametrize("seed,n", [(0, 20000), (5, 100000)])
def test_points_to_cylinder_and_sphere(cuda, seed, n):
    from btcdet_b200 import ops, synthetic as S
    from oracle import coords
    pts = S.lidar_like(n, seed=seed, az_density=4.0 if n > 50000 else 1.0)
    t = torch.from_numpy(pts).cuda()
    for sphere, fn in ((False, coords.absxyz_2_cylinxyz), (True, coords.absxyz_2_spherexyz)):
        got = ops.points_to_cylinder(t, sphere=sphere).cpu().numpy()
        want = fn(pts)
        np.testing.assert_array_equal(got[:, 0], want[:, 0])                      # rho / r
        np.testing.assert_array_equal(got[:, 3:], want[:, 3:])                    # intensity
        if not sphere:
            np.testing.assert_array_equal(got[:, 2], want[:, 2])                  # z
        ang_cols = (1, 2) if sphere else (1,)
        for c in ang_cols:
            assert int(coords.ulp_distance(got[:, c], want[:, c]).max()) <= ANGLE_ULP
    empty = ops.points_to_cylinder(torch.zeros(0, 4, device="cuda"))
    assert empty.shape == (0, 4)


def test_gpu_resident_input_pipeline_matches_dataset_side_path(cuda, oracle):
    """data_processor.py:105-190 on the device: a2 + cylindrical VoxelGeneratorV2 + det VoxelGeneratorV2 from raw points,
    against numpy a2 + the CPU voxeliser oracle per scene."""
    from btcdet_b200 import ops, synthetic as S
    from oracle import coords
    scenes = [S.lidar_like(20000, seed=40 + b) for b in range(3)]
    pts, offs = S.batch_points(scenes)
    occ, det = ops.voxelize_occ_and_det(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(),
                                        S.OCC_VOXEL_SIZE, S.OCC_RANGE, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"],
                                        S.DET_VOXEL_SIZE, S.KITTI_RANGE, S.DET_MAX_POINTS, S.DET_MAX_VOXELS["train"])
    gen_occ = oracle.VoxelGeneratorV2(S.OCC_VOXEL_SIZE, S.OCC_RANGE, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"])
    gen_det = oracle.VoxelGeneratorV2(S.DET_VOXEL_SIZE, S.KITTI_RANGE, S.DET_MAX_POINTS, S.DET_MAX_VOXELS["train"])
    for (vox, crd, cnt, mean, nv), gen, xform in ((occ, gen_occ, coords.absxyz_2_cylinxyz), (det, gen_det, None)):
        nv = nv.cpu().numpy()
        base = 0
        for b, sc in enumerate(scenes):
            want = gen.generate(xform(sc) if xform else sc)
            m = int(nv[b])
            if xform is not None:
                # points whose phi bin depends on the last ANGLE_ULP ulps (numpy's and CUDA's atan2f are different libms)
                phi = xform(sc)[:, 1]
                lo = np.floor((np.nextafter(phi, -np.inf, dtype=np.float32) - ANGLE_ULP * np.spacing(phi) - np.float32(S.OCC_RANGE[1]))
                              / np.float32(S.OCC_VOXEL_SIZE[1]))
                hi = np.floor((phi + ANGLE_ULP * np.spacing(phi) - np.float32(S.OCC_RANGE[1])) / np.float32(S.OCC_VOXEL_SIZE[1]))
                ambiguous = int((lo != hi).sum())
                if ambiguous:      # an edge point may move one voxel over: bounded, not bit-comparable
                    assert abs(m - want["voxel_num"]) <= ambiguous
                    base += m
                    continue
            assert m == want["voxel_num"], (b, m, want["voxel_num"])
            c = crd[base:base + m].cpu().numpy()
            assert (c[:, 0] == b).all()
            np.testing.assert_array_equal(c[:, 1:], want["coordinates"])
            np.testing.assert_array_equal(cnt[base:base + m].cpu().numpy(), want["num_points_per_voxel"])
            g = vox[base:base + m].cpu().numpy()
            if xform is None:
                np.testing.assert_array_equal(g, want["voxels"])
            else:   # cylindrical contents: rho, z, intensity exact; phi within the libm distance
                np.testing.assert_array_equal(g[..., [0, 2, 3]], want["voxels"][..., [0, 2, 3]])
                assert int(coords.ulp_distance(g[..., 1], want["voxels"][..., 1]).max()) <= ANGLE_ULP
            base += m
        assert base == int(nv[-1])
