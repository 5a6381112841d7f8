"""Pins the CPU oracle (oracle/) — the reference ships no golden vectors, so the
restatement of spconv 1.2.1 is pinned against an independent dense formulation with stock torch
ops (O2) and hand-derived known-answer tests (KATs i-iv of SURVEY §8c)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def _random_sparse(rng, batch, shape, n, c):
    cells = batch * int(np.prod(shape))
    flat = rng.choice(cells, size=min(n, cells), replace=False)
    rng.shuffle(flat)
    x = flat % shape[2]
    y = (flat // shape[2]) % shape[1]
    z = (flat // (shape[2] * shape[1])) % shape[0]
    b = flat // (shape[2] * shape[1] * shape[0])
    idx = np.stack([b, z, y, x], 1).astype(np.int32)
    feat = rng.standard_normal((idx.shape[0], c)).astype(np.float32)
    return idx, feat


def _dense(orc, feat, idx, shape, batch):
    return torch.from_numpy(orc.dense(feat, idx, shape, batch))


def _sample(dense, idx):
    idx = torch.as_tensor(idx).long()
    return dense[idx[:, 0], :, idx[:, 1], idx[:, 2], idx[:, 3]].numpy()


@pytest.mark.parametrize("ksize,dil", [(3, 1), ((3, 1, 1), 1), ((1, 3, 3), 1), (3, 2)])
def test_subm_equals_dense_conv_at_active_sites(oracle, ksize, dil):
    rng = np.random.default_rng(1)
    shape, batch = [7, 12, 10], 2
    idx, feat = _random_sparse(rng, batch, shape, 300, 5)
    ks = [ksize] * 3 if isinstance(ksize, int) else list(ksize)
    w = (rng.standard_normal((*ks, 5, 6)) * 0.2).astype(np.float32)
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(idx, batch, shape, ks, dilation=dil, subm=True)
    assert np.array_equal(outids, idx) and oshape == shape
    out = oracle.indice_conv(feat, w.reshape(-1, 5, 6), pairs, pair_num, idx.shape[0], subm=True)
    out_c = oracle.indice_conv(feat, w.reshape(-1, 5, 6), pairs, pair_num, idx.shape[0], subm=True, use_c=True)
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous()
    pad = [dil * (k // 2) for k in ks]
    ref = F.conv3d(_dense(oracle, feat, idx, shape, batch), wt, padding=pad, dilation=dil)
    np.testing.assert_allclose(out, _sample(ref, idx), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out_c, out, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("ksize,stride,pad", [(3, 2, 1), (3, 1, 1), ((3, 1, 1), (2, 1, 1), 0), (3, 2, (0, 1, 1)),
                                                (2, 2, 1), ((2, 1, 1), (2, 1, 1), 0), (3, (1, 2, 2), 1)])
def test_sparse_conv_equals_dense_conv_at_touched_sites(oracle, ksize, stride, pad):
    rng = np.random.default_rng(2)
    shape, batch = [9, 14, 11], 2
    idx, feat = _random_sparse(rng, batch, shape, 250, 4)
    ks = [ksize] * 3 if isinstance(ksize, int) else list(ksize)
    st = [stride] * 3 if isinstance(stride, int) else list(stride)
    pd = [pad] * 3 if isinstance(pad, int) else list(pad)
    w = (rng.standard_normal((*ks, 4, 3)) * 0.2).astype(np.float32)
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(idx, batch, shape, ks, st, pd)
    out = oracle.indice_conv(feat, w.reshape(-1, 4, 3), pairs, pair_num, outids.shape[0])
    wt = torch.from_numpy(w).permute(4, 3, 0, 1, 2).contiguous()
    ref = F.conv3d(_dense(oracle, feat, idx, shape, batch), wt, stride=st, padding=pd)
    assert list(ref.shape[2:]) == oshape
    np.testing.assert_allclose(out, _sample(ref, outids), rtol=1e-4, atol=1e-5)
    # active output sites == sites whose receptive field holds >= 1 active input, ascending flat key
    occ = F.conv3d(_dense(oracle, np.ones((idx.shape[0], 1), np.float32), idx, shape, batch),
                   torch.ones(1, 1, *ks), stride=st, padding=pd)
    want = torch.nonzero(occ[:, 0] > 0.5).numpy().astype(np.int32)  # nonzero() is row-major ascending
    assert np.array_equal(outids, want)


@pytest.mark.parametrize("stride,pad", [(2, 1), (1, 1), (2, 0)])
def test_transposed_conv_equals_dense_conv_transpose(oracle, stride, pad):
    rng = np.random.default_rng(3)
    shape, batch = [3, 7, 6], 2
    idx, feat = _random_sparse(rng, batch, shape, 60, 4)
    w = (rng.standard_normal((3, 3, 3, 4, 5)) * 0.2).astype(np.float32)
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(idx, batch, shape, 3, stride, pad, transpose=True)
    out = oracle.indice_conv(feat, w.reshape(-1, 4, 5), pairs, pair_num, outids.shape[0])
    wt = torch.from_numpy(w).permute(3, 4, 0, 1, 2).contiguous()
    ref = F.conv_transpose3d(_dense(oracle, feat, idx, shape, batch), wt, stride=stride, padding=pad)
    assert list(ref.shape[2:]) == oshape
    np.testing.assert_allclose(out, _sample(ref, outids), rtol=1e-4, atol=1e-5)
    occ = F.conv_transpose3d(_dense(oracle, np.ones((idx.shape[0], 1), np.float32), idx, shape, batch),
                             torch.ones(1, 1, 3, 3, 3), stride=stride, padding=pad)
    want = torch.nonzero(occ[:, 0] > 0.5).numpy().astype(np.int32)
    assert np.array_equal(outids, want)


def test_maxpool_equals_dense_maxpool_for_nonnegative_features_and_clamps_negatives(oracle):
    rng = np.random.default_rng(4)
    shape, batch = [9, 12, 10], 1
    idx, feat = _random_sparse(rng, batch, shape, 200, 2)
    feat = np.abs(feat)
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(idx, batch, shape, 3, 2, 1)
    out = oracle.indice_maxpool(feat, pairs, pair_num, outids.shape[0])
    ref = F.max_pool3d(_dense(oracle, feat, idx, shape, batch), 3, 2, 1)
    np.testing.assert_array_equal(out, _sample(ref, outids))
    # zero-initialised output: an all-negative input pools to 0
    out_neg = oracle.indice_maxpool(-feat - 1.0, pairs, pair_num, outids.shape[0])
    assert np.all(out_neg == 0.0)


def test_kat_impulse_reproduces_unflipped_kernel(oracle):
    """KAT (ii): a single active voxel through SparseConv3d(k3,s1,p1) gives W[k] at out = in + pad - k."""
    shape = [5, 5, 5]
    idx = np.array([[0, 2, 2, 2]], np.int32)
    feat = np.ones((1, 1), np.float32)
    w = np.arange(27, dtype=np.float32).reshape(27, 1, 1) + 1
    outids, pairs, pair_num, _ = oracle.get_indice_pairs(idx, 1, shape, 3, 1, 1)
    out = oracle.indice_conv(feat, w, pairs, pair_num, outids.shape[0])
    assert outids.shape[0] == 27
    for (b, z, y, x), v in zip(outids, out[:, 0]):
        kz, ky, kx = 2 + 1 - z, 2 + 1 - y, 2 + 1 - x
        assert v == w[(kz * 3 + ky) * 3 + kx, 0, 0]
    # transposed: out = in*s - p + k
    outids, pairs, pair_num, oshape = oracle.get_indice_pairs(idx, 1, shape, 3, 2, 1, transpose=True)
    out = oracle.indice_conv(feat, w, pairs, pair_num, outids.shape[0])
    assert oshape == [9, 9, 9]
    for (b, z, y, x), v in zip(outids, out[:, 0]):
        kz, ky, kx = z - 3, y - 3, x - 3
        assert v == w[(kz * 3 + ky) * 3 + kx, 0, 0]


def test_kat_fix_conv_counts_neighbours(oracle):
    """KAT (i): fixSparseConv3d weight 1/27 on all-ones features = (#inputs in the window)/27
    (btcdet/models/backbones_3d/spconv_backbone.py:45-48,812-828)."""
    rng = np.random.default_rng(5)
    shape = [9, 16, 16]
    idx, _ = _random_sparse(rng, 1, shape, 400, 1)
    feat = np.ones((idx.shape[0], 1), np.float32)
    w = np.full((27, 1, 1), 1.0 / 27, np.float32)
    outids, pairs, pair_num, _ = oracle.get_indice_pairs(idx, 1, shape, 3, 2, 1)
    out = oracle.indice_conv(feat, w, pairs, pair_num, outids.shape[0])
    nbr_out, _ = oracle.pairs_to_tables(pairs, pair_num, idx.shape[0], outids.shape[0])
    np.testing.assert_allclose(out[:, 0], (nbr_out >= 0).sum(1) / 27.0, rtol=1e-6)


def test_kat_output_shapes_of_both_backbones(oracle):
    """KAT (iii): spconv_backbone.py:996-1000 comments and the occ pyramid."""
    s = [41, 1600, 1408]
    s = oracle.conv_output_shape(s, [3] * 3, [2] * 3, [1] * 3, [1] * 3)
    assert s == [21, 800, 704]
    s = oracle.conv_output_shape(s, [3] * 3, [2] * 3, [1] * 3, [1] * 3)
    assert s == [11, 400, 352]
    s = oracle.conv_output_shape(s, [3] * 3, [2] * 3, [0, 1, 1], [1] * 3)
    assert s == [5, 200, 176]
    assert oracle.conv_output_shape(s, [3, 1, 1], [2, 1, 1], [0] * 3, [1] * 3) == [2, 200, 176]
    o = [9, 157, 209]
    o2 = oracle.conv_output_shape(o, [3] * 3, [2] * 3, [1] * 3, [1] * 3)
    o3 = oracle.conv_output_shape(o2, [3] * 3, [2] * 3, [1] * 3, [1] * 3)
    assert (o2, o3) == ([5, 79, 105], [3, 40, 53])
    d4 = oracle.deconv_output_shape(o3, [3] * 3, [2] * 3, [1] * 3, [1] * 3, [0] * 3)
    d5 = oracle.deconv_output_shape(d4, [3] * 3, [2] * 3, [1] * 3, [1] * 3, [0] * 3)
    assert (d4, d5) == ([5, 79, 105], [9, 157, 209])


def test_kat_voxelizer_edges(oracle):
    """KAT (iv): voxel edge, range max excluded, > max_points, > max_voxels (continue, not break)."""
    gen = oracle.VoxelGeneratorV2([0.5, 0.5, 0.5], [0, 0, 0, 2, 2, 1], max_num_points=2, max_voxels=3)
    assert list(gen.grid_size) == [4, 4, 2]
    pts = np.array([
        [0.5, 0.0, 0.0, 1],    # exactly on a voxel edge -> voxel x=1
        [2.0, 0.1, 0.1, 2],    # x == range max -> dropped
        [0.6, 0.1, 0.1, 3],    # same voxel as #0
        [0.7, 0.2, 0.2, 4],    # same voxel, over max_points -> dropped
        [-0.01, 0.1, 0.1, 5],  # below range -> dropped
        [1.9, 1.9, 0.9, 6],    # voxel (z1,y3,x3)
        [0.1, 1.1, 0.6, 7],    # voxel (z1,y2,x0)  -> third voxel, cap reached
        [1.1, 1.1, 0.1, 8],    # would be a 4th voxel -> skipped
        [1.95, 1.95, 0.95, 9],  # existing voxel #1 still accepts points after the cap
    ], np.float32)
    r = gen.generate(pts)
    assert r["voxel_num"] == 3
    np.testing.assert_array_equal(r["coordinates"], [[0, 0, 1], [1, 3, 3], [1, 2, 0]])
    np.testing.assert_array_equal(r["num_points_per_voxel"], [2, 2, 1])
    np.testing.assert_array_equal(r["voxels"][0, :, 3], [1, 3])
    np.testing.assert_array_equal(r["voxels"][1, :, 3], [6, 9])
    np.testing.assert_array_equal(r["voxels"][2, :, 3], [7, 0])
    # the lookup volume is restored: a second call gives the same answer
    r2 = gen.generate(pts)
    np.testing.assert_array_equal(r2["coordinates"], r["coordinates"])


def test_voxelizer_matches_plain_numpy_reference(oracle):
    from btcdet_b200 import synthetic as S
    pts = S.uniform(3000, seed=7)
    gen = oracle.VoxelGeneratorV2(S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 1000)
    r = gen.generate(pts)
    vs, lo = np.array(S.DET_VOXEL_SIZE, np.float32), np.array(S.KITTI_RANGE[:3], np.float32)
    c = np.floor((pts[:, :3] - lo) / vs).astype(np.int64)
    seen, counts = {}, []
    for i in range(pts.shape[0]):
        if np.any(c[i] < 0) or np.any(c[i] >= gen.grid_size):
            continue
        key = (c[i, 2], c[i, 1], c[i, 0])
        if key not in seen:
            if len(seen) >= 1000:
                continue
            seen[key] = len(seen)
            counts.append(0)
        if counts[seen[key]] < 5:
            counts[seen[key]] += 1
    assert r["voxel_num"] == len(seen) == 1000
    np.testing.assert_array_equal(r["coordinates"], np.array(list(seen.keys()), np.int32))
    np.testing.assert_array_equal(r["num_points_per_voxel"], counts)


def test_canonical_pair_order(oracle):
    rng = np.random.default_rng(9)
    idx, _ = _random_sparse(rng, 2, [6, 9, 9], 200, 1)
    for kw in (dict(subm=True), dict(stride=2, padding=1), dict(stride=2, padding=1, transpose=True)):
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(idx, 2, [6, 9, 9], 3, **kw)
        for k in range(27):
            ins = pairs[0, k, :pair_num[k]]
            assert np.all(np.diff(ins) > 0)
            assert np.all(pairs[:, k, pair_num[k]:] == -1)
