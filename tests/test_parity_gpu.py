"""GPU parity tests: the CUDA library (through its C ABI / the spconv shim) against the CPU oracle on
the same seeded inputs.  Integer outputs (voxel ids, coordinates, rulebooks) bit-exact; sparse-conv
activations within 1e-4 relative fp32 (north_star tolerance)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # max |a-b| / max |b|, the north_star's "1e-4 relative fp32"


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _gpu_voxelize(scenes, voxel_size, prange, max_points, max_voxels, want_mean=False):
    from btcdet_b200 import ops, synthetic as S
    pts, offs = S.batch_points(scenes)
    v, c, n, mean, nv = ops.voxelize(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(), voxel_size, prange,
                                     max_points, max_voxels, want_mean=want_mean)
    nv = nv.cpu().numpy()
    m = int(nv[-1])
    return v[:m].cpu().numpy(), c[:m].cpu().numpy(), n[:m].cpu().numpy(), (mean[:m].cpu().numpy() if want_mean else None), nv


# ---------------------------------------------------------------------------------------------
# voxelisation
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["config1_uniform2k", "lidar20k", "batch3", "cap_voxels", "occ_grid"])
def test_voxelize_bit_exact(cuda, oracle, case):
    from btcdet_b200 import synthetic as S
    vs, rg, mp, mv = S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000
    if case == "config1_uniform2k":
        scenes = [S.uniform(2000, seed=0)]
    elif case == "lidar20k":
        scenes = [S.lidar_like(20000, seed=1)]
    elif case == "batch3":
        scenes = [S.lidar_like(20000, seed=2), S.uniform(5000, seed=3), S.lidar_like(12000, seed=4)]
    elif case == "cap_voxels":
        scenes, mv = [S.uniform(30000, seed=5), S.lidar_like(20000, seed=6)], 4000
    else:  # coarse cylindrical-size grid: many points per voxel, max_points cap active
        scenes, vs, rg, mp, mv = [S.lidar_like(20000, seed=7), S.lidar_like(20000, seed=8)], [0.32, 0.5184, 0.36], \
            [0.0, -40.6944, -3.0, 70.4, 40.6944, 0.96], 12, 20000
    ov, oc, on = oracle.voxelize_batch(scenes, vs, rg, mp, mv)
    gv, gc, gn, gmean, nv = _gpu_voxelize(scenes, vs, rg, mp, mv, want_mean=True)
    assert gc.shape == oc.shape
    np.testing.assert_array_equal(gc, oc)
    np.testing.assert_array_equal(gn, on)
    np.testing.assert_array_equal(gv, ov)  # point rows are copies: bit-exact including zero padding
    mean = ov.sum(1) / np.maximum(on, 1)[:, None].astype(np.float32)
    np.testing.assert_allclose(gmean, mean, rtol=1e-6, atol=1e-6)
    if case == "cap_voxels":
        assert nv[0] == 4000  # the cap bites on the uniform scene


def test_voxelize_edge_cases(cuda, oracle):
    vs, rg = [0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 1]
    pts = np.array([[0.5, 0.0, 0.0, 1], [2.0, 0.1, 0.1, 2], [0.6, 0.1, 0.1, 3], [0.7, 0.2, 0.2, 4],
                    [-0.01, 0.1, 0.1, 5], [1.9, 1.9, 0.9, 6], [0.1, 1.1, 0.6, 7], [1.1, 1.1, 0.1, 8],
                    [1.95, 1.95, 0.95, 9], [np.nan, 0.1, 0.1, 10], [np.inf, 0.1, 0.1, 11]], np.float32)
    gv, gc, gn, _, nv = _gpu_voxelize([pts], vs, rg, 2, 3)
    np.testing.assert_array_equal(gc, [[0, 0, 0, 1], [0, 1, 3, 3], [0, 1, 2, 0]])
    np.testing.assert_array_equal(gn, [2, 2, 1])
    np.testing.assert_array_equal(gv[:, :, 3], [[1, 3], [6, 9], [7, 0]])
    # empty scene and all-out-of-range scene
    empty = np.zeros((0, 4), np.float32)
    far = np.full((10, 4), 100.0, np.float32)
    gv, gc, gn, _, nv = _gpu_voxelize([empty, far, pts], vs, rg, 2, 3)
    np.testing.assert_array_equal(nv, [0, 0, 3, 3])
    np.testing.assert_array_equal(gc[:, 0], [2, 2, 2])


def test_voxelgenerator_shim_matches_oracle(cuda, oracle):
    import spconv
    from btcdet_b200 import synthetic as S
    pts = S.lidar_like(20000, seed=11)
    gen = spconv.utils.VoxelGeneratorV2(voxel_size=S.DET_VOXEL_SIZE, point_cloud_range=S.KITTI_RANGE,
                                        max_num_points=5, max_voxels=16000)
    ref = oracle.VoxelGeneratorV2(S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000).generate(pts)
    out = gen.generate(pts)
    for k in ("voxels", "coordinates", "num_points_per_voxel"):
        assert isinstance(out[k], np.ndarray)
        np.testing.assert_array_equal(out[k], ref[k])
    assert out["voxel_num"] == ref["voxel_num"]
    assert list(gen.grid_size) == [1408, 1600, 40]


# ---------------------------------------------------------------------------------------------
# rulebooks
# ---------------------------------------------------------------------------------------------
def _scene_coords(oracle, seed, n=20000, batch=1, uniform=False):
    from btcdet_b200 import synthetic as S
    scenes = [(S.uniform if uniform else S.lidar_like)(n, seed=seed + b) for b in range(batch)]
    _, coords, _ = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    return coords


def _check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, padding, subm, transposed):
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(coords, batch, shape, ksize, stride, padding, 1, 0, subm,
                                                              transposed)
    assert rb.out_shape == oshape and rb.n_out == outids.shape[0]
    np.testing.assert_array_equal(rb.out_coords.cpu().numpy(), outids)
    nbr_out, nbr_in = oracle.pairs_to_tables(pairs, pair_num, coords.shape[0], outids.shape[0])
    np.testing.assert_array_equal(rb.nbr_out.cpu().numpy(), nbr_out)
    if rb.nbr_in is not None:
        np.testing.assert_array_equal(rb.nbr_in.cpu().numpy(), nbr_in)
    gp, gn = rb.pairs()  # spconv-format pair list, canonical order: bit-exact
    np.testing.assert_array_equal(gn.cpu().numpy(), pair_num)
    np.testing.assert_array_equal(gp.cpu().numpy(), pairs)
    return outids


@pytest.mark.parametrize("batch,uniform", [(1, False), (2, False), (1, True)])
def test_rulebooks_det_pyramid_bit_exact(cuda, oracle, batch, uniform):
    """subm1 / spconv2 / subm2 / spconv3 / subm3 / spconv4 / subm4 / conv_out geometry of
    VoxelBackBone8xOcc (spconv_backbone.py:657-707) on the KITTI det grid."""
    from btcdet_b200 import ops
    coords = _scene_coords(oracle, 20, batch=batch, uniform=uniform, n=8000 if uniform else 20000)
    shape = [41, 1600, 1408]
    levels = [(3, 2, 1), (3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)]
    for ksize, stride, pad in levels:
        c = torch.from_numpy(coords).cuda()
        rb = ops.rulebook_subm(c, batch, shape, 3)   # coordinate-hash lookup (unsorted rows)
        _check_rulebook(oracle, rb, coords, batch, shape, 3, 1, 1, True, False)
        rb_bitmap = ops.rulebook_subm(c, batch, shape, 3, index=ops.build_index(c, batch, shape, need_perm=True))
        assert torch.equal(rb_bitmap.nbr_out, rb.nbr_out)   # rank-bitmap + permutation lookup agrees
        rb = ops.rulebook_conv(c, batch, shape, ksize, stride, pad)
        coords = _check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, pad, False, False)
        # the output index of a strided conv doubles as the (sorted) index of the next subm layer
        rb2 = ops.rulebook_subm(rb.out_coords, batch, rb.out_shape, 3, index=rb.out_index)
        _check_rulebook(oracle, rb2, coords, batch, rb.out_shape, 3, 1, 1, True, False)
        shape = rb.out_shape
    assert shape == [2, 200, 176]


def test_rulebooks_occ_pyramid_with_transposed_bit_exact(cuda, oracle):
    """conv1 (dilating k3 s1 p1) / conv2 / conv3 / deconv4 / deconv5 of VoxelBackBoneDeconv
    (spconv_backbone.py:106-128) on the cylindrical occ grid [9,157,209]."""
    from btcdet_b200 import ops
    rng = np.random.default_rng(0)
    batch, shape = 2, [9, 157, 209]
    cells = batch * int(np.prod(shape))
    flat = rng.choice(cells, 9000, replace=False)
    coords = np.stack([flat // (9 * 157 * 209), (flat // (157 * 209)) % 9, (flat // 209) % 157, flat % 209], 1).astype(np.int32)
    for ksize, stride, pad, tr in [(3, 1, 1, False), (3, 2, 1, False), (3, 2, 1, False), (3, 2, 1, True), (3, 2, 1, True)]:
        rb = ops.rulebook_conv(torch.from_numpy(coords).cuda(), batch, shape, ksize, stride, pad, transposed=tr)
        coords = _check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, pad, False, tr)
        shape = rb.out_shape
    assert shape == [9, 157, 209]


def test_rulebook_empty_and_single(cuda, oracle):
    from btcdet_b200 import ops
    one = torch.tensor([[0, 2, 2, 2]], dtype=torch.int32, device="cuda")
    rb = ops.rulebook_conv(one, 1, [5, 5, 5], 3, 1, 1)
    assert rb.n_out == 27
    rb = ops.rulebook_subm(one, 1, [5, 5, 5], 3)
    assert rb.nbr_out.cpu().numpy().tolist() == [[-1] * 13 + [0] + [-1] * 13]
    empty = torch.zeros((0, 4), dtype=torch.int32, device="cuda")
    rb = ops.rulebook_conv(empty, 1, [5, 5, 5], 3, 2, 1)
    assert rb.n_out == 0


# ---------------------------------------------------------------------------------------------
# arithmetic
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cin,cout", [(16, 16), (4, 16), (6, 16), (34, 32), (32, 64), (64, 64), (64, 128), (256, 128),
                                      (32, 2), (32, 3)])
@pytest.mark.parametrize("kind", ["subm", "conv"])
def test_sparse_conv_forward_within_tolerance(cuda, oracle, cin, cout, kind):
    from btcdet_b200 import ops
    coords = _scene_coords(oracle, 30, n=6000)
    rng = np.random.default_rng(cin * 1000 + cout)
    shape = [41, 1600, 1408]
    if kind == "subm":
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, subm=True)
        rb = ops.rulebook_subm(torch.from_numpy(coords).cuda(), 1, shape, 3)
    else:
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, 2, 1)
        rb = ops.rulebook_conv(torch.from_numpy(coords).cuda(), 1, shape, 3, 2, 1)
    feat = rng.standard_normal((coords.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    ref = oracle.indice_conv(feat, w, pairs, pair_num, outids.shape[0], subm=(kind == "subm"), bias=bias)
    out = ops.sparse_conv_fwd(torch.from_numpy(feat).cuda(), rb.nbr_out, torch.from_numpy(w).cuda(),
                              torch.from_numpy(bias).cuda(), algo=1)
    assert rel_err(out.cpu().numpy(), ref) < REL_TOL
    # fused affine + relu epilogue
    scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.standard_normal(cout).astype(np.float32)
    out2 = ops.sparse_conv_fwd(torch.from_numpy(feat).cuda(), rb.nbr_out, torch.from_numpy(w).cuda(),
                               torch.from_numpy(bias).cuda(), torch.from_numpy(scale).cuda(),
                               torch.from_numpy(shift).cuda(), relu=True, algo=1)
    assert rel_err(out2.cpu().numpy(), np.maximum(ref * scale + shift, 0)) < REL_TOL


@pytest.mark.parametrize("cin,cout", [(32, 32), (16, 32), (32, 64), (64, 64), (64, 128), (128, 128), (256, 128),
                                      (48, 40), (20, 16), (4, 16), (16, 16), (8, 4), (6, 16)])
@pytest.mark.parametrize("kind", ["subm", "conv"])
def test_tensor_core_conv_within_tolerance(cuda, oracle, cin, cout, kind):
    """tcgen05 3xTF32 tile vs the oracle (and vs the fp32 FFMA tile): same 1e-4 bar."""
    from btcdet_b200 import ops
    if not ops.tc_supported(27, cin, cout):
        pytest.skip("shape not covered by the tensor-core tile")
    coords = _scene_coords(oracle, 31, n=9000)
    rng = np.random.default_rng(cin * 1000 + cout + 7)
    shape = [41, 1600, 1408]
    if kind == "subm":
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, subm=True)
        rb = ops.rulebook_subm(torch.from_numpy(coords).cuda(), 1, shape, 3)
    else:
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, 2, 1)
        rb = ops.rulebook_conv(torch.from_numpy(coords).cuda(), 1, shape, 3, 2, 1)
    feat = rng.standard_normal((coords.shape[0], cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.standard_normal(cout).astype(np.float32)
    ref = oracle.indice_conv(feat, w, pairs, pair_num, outids.shape[0], subm=(kind == "subm"), bias=bias)
    wt = torch.from_numpy(w).cuda()
    packed = ops.tc_pack_weight(wt)
    f = torch.from_numpy(feat).cuda()
    out = ops.sparse_conv_fwd_tc(f, rb.nbr_out, packed, cin, cout, torch.from_numpy(bias).cuda())
    assert rel_err(out.cpu().numpy(), ref) < REL_TOL
    out2 = ops.sparse_conv_fwd_tc(f, rb.nbr_out, packed, cin, cout, torch.from_numpy(bias).cuda(),
                                  torch.from_numpy(scale).cuda(), torch.from_numpy(shift).cuda(), relu=True)
    assert rel_err(out2.cpu().numpy(), np.maximum(ref * scale + shift, 0)) < REL_TOL
    ffma = ops.sparse_conv_fwd(f, rb.nbr_out, wt, torch.from_numpy(bias).cuda(), algo=1)
    assert rel_err(out.cpu().numpy(), ffma.cpu().numpy()) < 2e-5
    # mask-sorted rows (block skipping): per-row arithmetic is unchanged -> bit-identical results
    nbr_sorted, out_rows = rb.sorted_rows()
    out3 = ops.sparse_conv_fwd_tc(f, nbr_sorted, packed, cin, cout, torch.from_numpy(bias).cuda(), out_rows=out_rows)
    assert torch.equal(out3, out)


@pytest.mark.parametrize("n_out,K,cin,cout,density,run", [
    (30000, 27, 64, 64, 0.35, 97), (45000, 27, 16, 32, 0.08, 97), (20000, 3, 64, 128, 0.45, 97), (52000, 27, 4, 16, 0.13, 97),
    (19000, 8, 32, 32, 0.5, 97), (300, 27, 32, 32, 0.02, 97),
    # one or two active chunks per tile, several tiles per CTA: producer groups own no stage in consecutive tiles
    # (regression: a producer blocking on a later tile's chunk list while holding gathers in flight deadlocked)
    (70000, 27, 4, 16, 0.03, 1024), (70000, 27, 16, 16, 0.03, 1024), (60000, 27, 32, 32, 0.02, 640)])
def test_tensor_core_multi_tile_persistent(cuda, n_out, K, cin, cout, density, run):
    """More tiles than SMs (each persistent CTA walks several tiles, variable active-chunk counts per tile), capacity
    larger than the live count, clustered neighbourhood patterns: tcgen05 tile == FFMA tile within 2e-5, and the
    mask-sorted table gives bit-identical rows."""
    from btcdet_b200 import ops
    rng = np.random.default_rng(n_out + K)
    n_in = 40000
    pat = rng.random((64, K)) < density
    valid = pat[(np.arange(n_out) // run) % 64] ^ (rng.random((n_out, K)) < 0.01)   # runs of equal patterns + noise
    valid[:, K // 2] |= ~valid.any(1)
    nbr = np.where(valid, rng.integers(0, n_in, (n_out, K)), -1).astype(np.int32)
    cap = n_out + 1000
    table = torch.full((cap, K), 123456789, dtype=torch.int32, device="cuda")     # garbage past the live rows
    table[:n_out] = torch.from_numpy(nbr).cuda()
    n_dev = torch.tensor([n_out], dtype=torch.int32, device="cuda")
    feat = torch.from_numpy(rng.standard_normal((n_in, cin)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((K, cin, cout)) * 0.1).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).cuda()
    packed = ops.tc_pack_weight(w)
    ffma = ops.sparse_conv_fwd(feat, table[:n_out].contiguous(), w, bias, algo=1)
    out = torch.zeros((cap, cout), device="cuda")
    ops.sparse_conv_fwd_tc(feat, table, packed, cin, cout, bias, n_out_dev=n_dev, out=out)
    torch.cuda.synchronize()
    assert rel_err(out[:n_out].cpu().numpy(), ffma.cpu().numpy()) < 2e-5
    assert float(out[n_out:].abs().sum()) == 0.0                                   # rows past the live count untouched
    nbr_sorted, out_rows = ops.rulebook_sort_rows(table, n_out_dev=n_dev)
    out2 = torch.zeros((cap, cout), device="cuda")
    ops.sparse_conv_fwd_tc(feat, nbr_sorted, packed, cin, cout, bias, n_out_dev=n_dev, out=out2, out_rows=out_rows)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("npw,cat,dyn", [(16, 0, 1), (8, 0, 1), (8, 0, 0), (16, 0, 0)])
@pytest.mark.parametrize("n_out,K,cin,cout,density,run", [
    (45000, 27, 32, 32, 0.27, 97), (33000, 27, 64, 64, 0.35, 300), (70000, 27, 16, 16, 0.03, 1024),
    (21000, 3, 64, 128, 0.45, 97), (130, 27, 32, 64, 0.2, 7)])
def test_tensor_core_tile_variants(cuda, npw, cat, dyn, n_out, K, cin, cout, density, run):
    """Every variant of the tcgen05 tile (8 / 16 producer warps, 3-MMA / concatenated [B_hi|B_lo] 2-MMA k-steps, static /
    dynamic tile scheduling) against the fp32 FFMA tile; the dynamic scheduler's counter must return to zero after
    every launch (three launches, identical bits), and scheduling must not change a single bit."""
    from btcdet_b200 import ops
    rng = np.random.default_rng(n_out + K + cin)
    n_in = 40000
    pat = rng.random((64, K)) < density
    valid = pat[(np.arange(n_out) // run) % 64] ^ (rng.random((n_out, K)) < 0.01)
    valid[:, K // 2] |= ~valid.any(1)
    nbr = np.where(valid, rng.integers(0, n_in, (n_out, K)), -1).astype(np.int32)
    cap = n_out + 777
    table = torch.full((cap, K), 123456789, dtype=torch.int32, device="cuda")
    table[:n_out] = torch.from_numpy(nbr).cuda()
    n_dev = torch.tensor([n_out], dtype=torch.int32, device="cuda")
    feat = torch.from_numpy(rng.standard_normal((n_in, cin)).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((K, cin, cout)) * 0.1).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).cuda()
    packed = ops.tc_pack_weight(w)
    ffma = ops.sparse_conv_fwd(feat, table[:n_out].contiguous(), w, bias, algo=1)
    try:
        ops.tc_config(npw, cat, dyn)
        outs = []
        for rep in range(3):
            out = torch.zeros((cap, cout), device="cuda")
            ops.sparse_conv_fwd_tc(feat, table, packed, cin, cout, bias, n_out_dev=n_dev, out=out)
            outs.append(out)
        torch.cuda.synchronize()
        assert rel_err(outs[0][:n_out].cpu().numpy(), ffma.cpu().numpy()) < 2e-5
        assert float(outs[0][n_out:].abs().sum()) == 0.0
        assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
        ops.tc_config(npw, cat, 0)
        static = torch.zeros((cap, cout), device="cuda")
        ops.sparse_conv_fwd_tc(feat, table, packed, cin, cout, bias, n_out_dev=n_dev, out=static)
        assert torch.equal(static, outs[0])
    finally:
        ops.tc_config(16, 0, 1)


@pytest.mark.parametrize("n,K", [(1, 27), (127, 27), (2048, 27), (2049, 8), (9000, 27), (5000, 64), (3000, 3), (4100, 33)])
def test_rulebook_sort_rows(cuda, n, K):
    """btc_rulebook_sort_rows: a permutation inside 2048-row windows, ordered by (valid-offset mask, original row)."""
    from btcdet_b200 import ops
    rng = np.random.default_rng(n * 100 + K)
    pat = rng.integers(0, 2, (40, K))                            # few distinct neighbourhood patterns + noise
    valid = pat[rng.integers(0, 40, n)] ^ (rng.random((n, K)) < 0.02)
    nbr = np.where(valid, rng.integers(0, 100000, (n, K)), -1).astype(np.int32)
    cap = n + 37                                                 # capacity larger than the live count
    buf = np.full((cap, K), 7, np.int32)
    buf[:n] = nbr
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    nbr_sorted, out_rows = ops.rulebook_sort_rows(torch.from_numpy(buf).cuda(), n_out_dev=n_dev)
    rows = out_rows[:n].cpu().numpy()
    got = nbr_sorted[:n].cpu().numpy()
    np.testing.assert_array_equal(got, nbr[rows])
    mask = [int(sum(1 << k for k in range(K) if nbr[r, k] >= 0)) for r in range(n)]
    for w0 in range(0, n, 2048):
        w = rows[w0:w0 + 2048]
        assert sorted(w.tolist()) == list(range(w0, min(w0 + 2048, n)))
        keys = [(mask[r], r) for r in w]
        assert keys == sorted(keys)


def test_config1_golden(cuda, oracle):
    """BASELINE.json configs[0]: 2k uniform points -> voxelize -> one SubMConv3d(16->16,k3); compared
    with the committed golden fixture (generated by tests/golden/make_golden.py from the oracle)."""
    import os
    from btcdet_b200 import ops, synthetic as S
    path = os.path.join(os.path.dirname(__file__), "golden", "config1.npz")
    g = np.load(path)
    pts = S.uniform(2000, seed=0)
    gv, gc, gn, _, _ = _gpu_voxelize([pts], S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    np.testing.assert_array_equal(gc[:, 1:], g["coordinates"])
    np.testing.assert_array_equal(gn, g["num_points_per_voxel"])
    rb = ops.rulebook_subm(torch.from_numpy(gc).cuda(), 1, [41, 1600, 1408], 3)
    gp, gpn = rb.pairs()
    np.testing.assert_array_equal(gp.cpu().numpy(), g["indice_pairs"])
    np.testing.assert_array_equal(gpn.cpu().numpy(), g["indice_pair_num"])
    out = ops.sparse_conv_fwd(torch.from_numpy(g["features"]).cuda(), rb.nbr_out,
                              torch.from_numpy(g["weight"]).cuda().reshape(27, 16, 16), algo=1)
    assert rel_err(out.cpu().numpy(), g["out_features"]) < REL_TOL


def test_sparse_conv_backward_matches_oracle_autograd(cuda, oracle):
    """dX, dW, db of the CUDA library vs autograd through the oracle formulation (CPU torch)."""
    from btcdet_b200 import ops
    coords = _scene_coords(oracle, 40, n=3000)
    shape = [41, 1600, 1408]
    rng = np.random.default_rng(0)
    for kind, cin, cout, algo in [("subm", 16, 32, 1), ("conv", 32, 16, 1), ("conv", 6, 20, 1),
                                  ("subm", 32, 32, 0), ("subm", 64, 64, 0), ("conv", 32, 64, 0), ("conv", 64, 128, 0)]:
        # algo 0: forward and dX on the tcgen05 tile (3xTF32), dW / db on the fp32 kernels
        if kind == "subm":
            outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, subm=True)
            rb = ops.rulebook_subm(torch.from_numpy(coords).cuda(), 1, shape, 3)
        else:
            outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, 2, 1)
            rb = ops.rulebook_conv(torch.from_numpy(coords).cuda(), 1, shape, 3, 2, 1)
        feat = torch.from_numpy(rng.standard_normal((coords.shape[0], cin)).astype(np.float32))
        w = torch.from_numpy((rng.standard_normal((27, cin, cout)) * 0.1).astype(np.float32))
        b = torch.from_numpy(rng.standard_normal(cout).astype(np.float32))
        go = torch.from_numpy(rng.standard_normal((outids.shape[0], cout)).astype(np.float32))
        f1, w1, b1 = feat.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
        out = torch.zeros(outids.shape[0], cout)
        pr = torch.from_numpy(pairs).long()
        for k in range(27):
            nh = int(pair_num[k])
            if nh:
                out = out.index_add(0, pr[1, k, :nh], f1[pr[0, k, :nh]] @ w1[k])
        (out + b1).backward(go)
        f2, w2, b2 = (t.clone().cuda().requires_grad_() for t in (feat, w, b))
        y = ops.SparseConvFunction.apply(f2, w2, b2, rb, algo)
        y.backward(go.cuda())
        assert rel_err(f2.grad.cpu().numpy(), f1.grad.numpy()) < REL_TOL
        assert rel_err(w2.grad.cpu().numpy(), w1.grad.numpy()) < REL_TOL
        assert rel_err(b2.grad.cpu().numpy(), b1.grad.numpy()) < REL_TOL


def test_maxpool_and_dense(cuda, oracle):
    from btcdet_b200 import ops
    coords = _scene_coords(oracle, 50, n=8000, batch=2)
    shape = [41, 1600, 1408]
    rng = np.random.default_rng(1)
    feat = rng.standard_normal((coords.shape[0], 2)).astype(np.float32)
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(coords, 2, shape, 3, 2, 1)
    rb = ops.rulebook_conv(torch.from_numpy(coords).cuda(), 2, shape, 3, 2, 1)
    ref = oracle.indice_maxpool(feat, pairs, pair_num, outids.shape[0])
    f = torch.from_numpy(feat).cuda().requires_grad_()
    out = ops.SparseMaxPoolFunction.apply(f, rb)
    np.testing.assert_array_equal(out.detach().cpu().numpy(), ref)  # max is exact
    out.sum().backward()
    # gradient goes to inputs equal to the pooled maximum; pooled zeros (all-negative windows) get none
    g = f.grad.cpu().numpy()
    assert g.shape == feat.shape and np.all(g >= 0) and g.sum() > 0
    # dense(): exact scatter, channels first
    small_shape = [5, 200, 176]
    rng2 = np.random.default_rng(2)
    flat = rng2.choice(2 * 5 * 200 * 176, 5000, replace=False)
    idx = np.stack([flat // (5 * 200 * 176), (flat // (200 * 176)) % 5, (flat // 176) % 200, flat % 176], 1).astype(np.int32)
    fd = rng2.standard_normal((5000, 7)).astype(np.float32)
    fd_t = torch.from_numpy(fd).cuda().requires_grad_()
    d = ops.ToDenseFunction.apply(fd_t, torch.from_numpy(idx).cuda(), 2, small_shape)
    np.testing.assert_array_equal(d.detach().cpu().numpy(), oracle.dense(fd, idx, small_shape, 2))
    wgt = torch.randn_like(d)
    (d * wgt).sum().backward()
    li = torch.from_numpy(idx).long()
    np.testing.assert_array_equal(fd_t.grad.cpu().numpy(), wgt.cpu()[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]].numpy())


def test_revoxelize_sorted_matches_torch_unique(cuda):
    """add_occ_template.py:262-268 semantics: unique rows sorted lexicographically, counts, stable slots."""
    from btcdet_b200 import ops
    g = torch.Generator().manual_seed(0)
    batch, shape = 2, [40, 1600, 1408]
    base = torch.stack([torch.randint(0, batch, (4000,), generator=g), torch.randint(0, 40, (4000,), generator=g),
                        torch.randint(700, 760, (4000,), generator=g), torch.randint(600, 650, (4000,), generator=g)], 1)
    coords = torch.cat([base, base[:1500], base[:300]], 0)  # duplicates -> multi-point voxels
    coords = coords[torch.randperm(coords.shape[0], generator=g)]
    feat = torch.randn(coords.shape[0], 6, generator=g)
    vox, cnt, vc = ops.revoxelize_sorted(coords.cuda(), feat.cuda(), batch, shape)
    u, inv, counts = torch.unique(coords, dim=0, sorted=True, return_inverse=True, return_counts=True)
    assert torch.equal(vc.cpu(), u) and torch.equal(cnt.cpu(), counts)
    assert vox.shape == (u.shape[0], int(counts.max()), 6)
    # stable: slot = number of earlier points in the same voxel
    ref = torch.zeros_like(vox.cpu())
    seen = {}
    for i in range(coords.shape[0]):
        v = int(inv[i])
        s = seen.get(v, 0)
        ref[v, s] = feat[i]
        seen[v] = s + 1
    assert torch.equal(vox.cpu(), ref)


@pytest.mark.parametrize("n_out,K,cin,cout,density,run", [
    (45000, 27, 32, 32, 0.27, 97), (33000, 27, 64, 64, 0.35, 300), (21000, 3, 64, 128, 0.45, 97), (130, 27, 32, 64, 0.2, 7),
    (52000, 27, 128, 64, 0.2, 50), (9000, 27, 32, 128, 0.3, 11)])
def test_tensor_core_split_format(cuda, n_out, K, cin, cout, density, run):
    """Split (bf16 hi / lo) feature format on the input and / or output side of the tcgen05 tile against the fp32 FFMA tile:
    every combination inside the 1e-4 bar; conversion round trip to 2^-16."""
    from btcdet_b200 import ops
    rng = np.random.default_rng(n_out + K + cin + cout)
    n_in = 40000
    pat = rng.random((64, K)) < density
    valid = pat[(np.arange(n_out) // run) % 64] ^ (rng.random((n_out, K)) < 0.01)
    valid[:, K // 2] |= ~valid.any(1)
    nbr = torch.from_numpy(np.where(valid, rng.integers(0, n_in, (n_out, K)), -1).astype(np.int32)).cuda()
    n_dev = torch.tensor([n_out], dtype=torch.int32, device="cuda")
    feat = torch.from_numpy((rng.standard_normal((n_in, cin)) * np.exp(rng.uniform(-3, 3, (n_in, 1)))).astype(np.float32)).cuda()
    w = torch.from_numpy((rng.standard_normal((K, cin, cout)) * 0.1).astype(np.float32)).cuda()
    bias = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).cuda()
    scale = torch.from_numpy(rng.uniform(0.5, 1.5, cout).astype(np.float32)).cuda()
    shift = torch.from_numpy(rng.standard_normal(cout).astype(np.float32)).cuda()
    want = ops.sparse_conv_fwd(feat, nbr, w, bias, scale, shift, relu=True, algo=1).cpu().numpy()
    fs = ops.features_to_split(feat)
    back = ops.features_from_split(fs)
    assert float(((back - feat).abs() / feat.abs().clamp_min(1e-30)).max()) <= 2.0 ** -16
    pk32, pks = ops.tc_pack_weight(w), ops.tc_pack_weight_split(w)
    from btcdet_b200 import _lib as _lib_mod
    tmask = torch.zeros((n_out + 127) // 128, dtype=torch.int64, device="cuda")
    tiles = (n_out + 127) // 128
    torder = torch.zeros(int(_lib_mod.load().btc_rulebook_tile_order_ints(n_out)), dtype=torch.int32, device="cuda")
    from btcdet_b200 import _lib
    import ctypes
    _lib.check(_lib.load().btc_rulebook_tile_meta(ctypes.c_void_p(nbr.data_ptr()), n_out, ctypes.c_void_p(n_dev.data_ptr()), K,
                                                  ctypes.c_void_p(tmask.data_ptr()), ctypes.c_void_p(torder.data_ptr()),
                                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "tile_meta")
    to = torder.cpu().numpy()
    counts, buckets = to[:65], to[65:].reshape(65, tiles)
    assert int(counts.sum()) == tiles
    listed = sorted(int(t) for c in range(65) for t in buckets[c, :counts[c]])
    assert listed == list(range(tiles))                                   # every live tile in exactly one class bucket
    pop = np.array([bin(int(v) & (2 ** 64 - 1)).count("1") for v in tmask.cpu().numpy()])
    assert all(pop[t] == c for c in range(65) for t in buckets[c, :counts[c]])
    for in_split, out_split in ((True, False), (True, True), (False, True), (False, False)):
        if not ops.tc_split_supported(K, cin, cout, in_split, out_split):
            continue
        for meta in (False, True):
            out = ops.sparse_conv_fwd_tc_split(fs if in_split else feat, nbr, pks if in_split else pk32, cin, cout, in_split,
                                               out_split, bias, scale, shift, relu=True, n_out_dev=n_dev,
                                               tile_mask=tmask if meta else None, tile_order=torder if meta else None)
            got = (ops.features_from_split(out) if out_split else out).cpu().numpy()
            assert rel_err(got, want) < 1e-4, (in_split, out_split, meta, rel_err(got, want))
