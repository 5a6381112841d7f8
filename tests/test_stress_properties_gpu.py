"""BASELINE config 5 (stress): 100 k-point dense clouds at voxel [0.025, 0.025, 0.05] — grid [81, 3200, 2816] per scene
(730 M cells, 64-bit cell keys once batched), ~46 k voxels per scene.  The CPU oracle needs minutes here, so parity is
asserted through the size-independent properties of tests/properties.py (whose checkers are pinned against the oracle at
small sizes in tests/test_properties_cpu.py), plus linearity of the convolution and agreement of the two conv tiles."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_config5_integer_path_properties_and_conv_linearity(cuda):
    import properties as P
    from btcdet_b200 import ops, synthetic as S
    B = 2
    vs, rng = [0.025, 0.025, 0.05], S.KITTI_RANGE
    grid = ops.voxel_grid_size(vs, rng)
    sparse_shape = [grid[2] + 1, grid[1], grid[0]]
    assert B * sparse_shape[0] * sparse_shape[1] * sparse_shape[2] > 2 ** 30      # keys need more than 31 bits from B = 3 on
    scenes = [S.lidar_like(100000, seed=50 + b, az_density=1.5, point_range=[0, -20, -3, 35.2, 20, 1]) for b in range(B)]
    pts, offs = S.batch_points(scenes)
    pts_d, offs_d = torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda()
    max_vox, max_pts = 150000, 5
    v, c, npnt, mean, nv = ops.voxelize(pts_d, offs_d, vs, rng, max_pts, max_vox, want_mean=True, grid=grid)
    m = P.check_voxelization(pts_d, offs_d, v, c, npnt, nv, vs, rng, grid, max_pts, max_vox)
    assert m > 50000                                  # ~46 k voxels per 100 k-point scene at this voxel size
    want_mean = v[:m].sum(1) / npnt[:m].clamp(min=1).view(-1, 1).float()
    torch.testing.assert_close(mean[:m], want_mean, rtol=1e-5, atol=1e-5)

    coords = c[:m].contiguous()
    idx = ops.build_hash(coords, B, sparse_shape)
    rb1 = ops.rulebook_subm(coords, B, sparse_shape, 3, index=idx)                  # hash probes, unsorted rows
    p1 = P.check_subm_table(coords, rb1.nbr_out, sparse_shape, [3, 3, 3])
    rb2 = ops.rulebook_conv(coords, B, sparse_shape, 3, 2, 1)                       # rank bitmap over 2 x 91 M cells
    p2 = P.check_conv_tables(coords, rb2.out_coords, rb2.nbr_out, rb2.nbr_in, sparse_shape, rb2.out_shape,
                             [3, 3, 3], [2, 2, 2], [1, 1, 1])
    rb3 = ops.rulebook_subm(rb2.out_coords, B, rb2.out_shape, 3, index=rb2.out_index)   # bitmap probes, sorted rows
    p3 = P.check_subm_table(rb2.out_coords, rb3.nbr_out, rb2.out_shape, [3, 3, 3])
    assert p1 > m and p2 >= m and p3 > rb2.n_out

    # convolution at this size: linear in its input, and the tcgen05 tile agrees with the fp32 FFMA tile
    g = torch.Generator(device="cpu").manual_seed(0)
    n2 = rb2.n_out
    f1 = torch.randn(n2, 32, generator=g).cuda()
    f2 = torch.randn(n2, 32, generator=g).cuda()
    w = (torch.randn(27, 32, 32, generator=g) * 0.1).cuda()
    pk = ops.tc_pack_weight(w)
    y1 = ops.sparse_conv_fwd_tc(f1, rb3.nbr_out, pk, 32, 32)
    y2 = ops.sparse_conv_fwd_tc(f2, rb3.nbr_out, pk, 32, 32)
    y12 = ops.sparse_conv_fwd_tc(2.0 * f1 + 3.0 * f2, rb3.nbr_out, pk, 32, 32)
    assert _rel(y12, 2.0 * y1 + 3.0 * y2) < 1e-4
    ffma = ops.sparse_conv_fwd(f1, rb3.nbr_out, w, algo=1)
    assert _rel(y1, ffma) < 2e-5
    # a row without neighbours other than itself sees only the centre weight
    lonely = torch.nonzero((rb3.nbr_out >= 0).sum(1) == 1)[:, 0]
    if lonely.numel():
        torch.testing.assert_close(y1[lonely], f1[lonely] @ w[13], rtol=1e-4, atol=1e-4)
