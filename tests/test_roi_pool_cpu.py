"""RoI grid pooling on the CPU: the oracle against the reference's own code / known answers, and the
CUDA KERNEL SOURCE (btcdet_b200/csrc/roi_pool_kernels.cuh) executed under the lock-step warp emulation of
tests/host_emul/ against the oracle.  The emulation is test infrastructure: the product has no CPU path."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "host_emul", "roi_pool_emul.cpp")
    out_dir = os.path.join(HERE, "host_emul", "_build")
    os.makedirs(out_dir, exist_ok=True)
    san = os.environ.get("BTC_EMUL_SANITIZE", "")      # "address": AddressSanitizer build (run pytest under LD_PRELOAD=libasan)
    lib = os.path.join(out_dir, "libroi_pool_emul%s.so" % ("_" + san if san else ""))
    deps = [src, os.path.join(HERE, "host_emul", "cuda_emul.h"),
            os.path.join(ROOT, "btcdet_b200", "csrc", "roi_pool_kernels.cuh"), os.path.join(ROOT, "btcdet_b200", "csrc", "common.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-w"] +
                       (["-g", "-fno-omit-frame-pointer", "-fsanitize=" + san, "-DEMUL_THREADS"] if san else []) +
                       ["-I/usr/local/cuda/include", src, "-o", lib], check=True, capture_output=True)
    lib = ctypes.CDLL(lib)
    lib.emul_ball_query_stack.argtypes = [I, I, I, P, P, P, P, P, P, P, I]
    lib.emul_group_points_stack.argtypes = [I, I, I, I, P, P, P, P, P, I]
    lib.emul_group_points_stack_grad.argtypes = [I, I, I, I, P, P, P, P, P, I]
    lib.emul_trilinear_sparse.argtypes = [P, P, I, I, I, P, P, P, L, L, I, I, P, I, P, P, P, I]
    lib.emul_trilinear_sparse_grad.argtypes = [P, P, I, I, P, P, I, I, I, P, P, P, L, L, I, P, I]
    return lib


def _p(a):
    return a.ctypes.data_as(P)


def _scene(rng, counts, qcounts, spread=4.0):
    xyz = (rng.random((sum(counts), 3)) * spread).astype(np.float32)
    new_xyz = (rng.random((sum(qcounts), 3)) * spread).astype(np.float32)
    return xyz, np.array(counts, np.int32), new_xyz, np.array(qcounts, np.int32)


# ---- oracle: known answers ------------------------------------------------------------------------------------------
def test_ball_query_oracle_known_answers(oracle):
    from oracle import roi_pool as R
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0.5, 0, 0], [5, 5, 5], [0.1, 0, 0]], np.float32)
    q = np.array([[0, 0, 0], [9, 9, 9], [5, 5, 5.5]], np.float32)
    idx = R.ball_query_stack(0.6, 4, xyz, [5], q, [3])
    assert idx.tolist() == [[0, 2, 4, 0], [-1, 0, 0, 0], [3, 3, 3, 3]]     # index order; first hit repeated; empty ball
    idx = R.ball_query_stack(0.6, 2, xyz, [5], q, [3])
    assert idx.tolist() == [[0, 2], [-1, 0], [3, 3]]                       # nsample cut
    # d2 < r*r is strict and evaluated in fp32: a point at exactly r is outside
    idx = R.ball_query_stack(1.0, 2, xyz, [5], q[:1], [1])
    assert idx.tolist() == [[0, 2]]
    # scenes: query 1 only sees the points of scene 1, indices are scene-local
    idx = R.ball_query_stack(0.6, 3, xyz, [3, 2], np.array([[0, 0, 0], [0, 0, 0]], np.float32), [1, 1])
    assert idx.tolist() == [[0, 2, 0], [1, 1, 1]]


def test_ball_query_oracle_vs_numpy_bruteforce(oracle):
    from oracle import roi_pool as R
    rng = np.random.default_rng(0)
    xyz, cnt, q, qcnt = _scene(rng, [300, 0, 500], [40, 7, 60])
    for radius, ns in ((0.4, 16), (1.2, 32)):
        idx = R.ball_query_stack(radius, ns, xyz, cnt, q, qcnt)
        pstart = np.concatenate([[0], np.cumsum(cnt)])
        qscene = np.repeat(np.arange(3), qcnt)
        for m in range(q.shape[0]):
            p = xyz[pstart[qscene[m]]:pstart[qscene[m] + 1]].astype(np.float64)
            d2 = ((p - q[m].astype(np.float64)) ** 2).sum(1)
            r2 = float(np.float32(radius) * np.float32(radius))
            if np.any(np.abs(d2 - r2) < 1e-5):
                continue                                   # fp32 rounding may decide these: not a brute-force matter
            hit = np.nonzero(d2 < r2)[0][:ns]
            want = np.zeros(ns, np.int32)
            if len(hit) == 0:
                want[0] = -1
            else:
                want[:] = hit[0]
                want[:len(hit)] = hit
            assert idx[m].tolist() == want.tolist(), (m, radius)


# ---- kernel source under the warp emulation vs the oracle ---------------------------------------------------------
@pytest.mark.parametrize("counts,qcounts,blocks", [([400, 350], [30, 27], 3), ([200, 0, 129], [5, 3, 6], 2), ([64], [1], 1),
                                                   ([33, 31, 1], [2, 2, 2], 5)])
def test_emulated_ball_query_kernel_is_bit_exact(oracle, emul, counts, qcounts, blocks):
    from oracle import roi_pool as R
    rng = np.random.default_rng(len(counts) * 100 + sum(qcounts))
    xyz, cnt, q, qcnt = _scene(rng, counts, qcounts, spread=3.0)
    radii, nsamples = [0.4, 0.8, 1.2, 2.4], [16, 16, 32, 64]
    M = q.shape[0]
    for n_r in (4, 3, 1):
        outs = [np.full((M, nsamples[r]), -77, np.int32) for r in range(n_r)]
        ptrs = (ctypes.c_void_p * n_r)(*[o.ctypes.data for o in outs])
        rad = np.array(radii[:n_r], np.float32)
        nsm = np.array(nsamples[:n_r], np.int32)
        assert emul.emul_ball_query_stack(len(counts), M, n_r, _p(rad), _p(nsm), _p(q), _p(qcnt), _p(xyz), _p(cnt), ptrs, blocks) == 0
        for r in range(n_r):
            want = R.ball_query_stack(radii[r], nsamples[r], xyz, cnt, q, qcnt)
            assert np.array_equal(outs[r], want), (n_r, r)
    assert (R.ball_query_stack(2.4, 64, xyz, cnt, q, qcnt)[:, 0] >= 0).any()


def test_emulated_group_points_and_grad(oracle, emul):
    from oracle import roi_pool as R
    rng = np.random.default_rng(5)
    xyz, cnt, q, qcnt = _scene(rng, [150, 90], [20, 13], spread=2.0)
    idx = R.ball_query_stack(0.8, 16, xyz, cnt, q, qcnt)
    idx[idx[:, 0] == -1] = 0
    feats = rng.standard_normal((xyz.shape[0], 5)).astype(np.float32)
    M, ns = idx.shape
    out = np.full((M, 5, ns), np.nan, np.float32)
    assert emul.emul_group_points_stack(2, M, 5, ns, _p(feats), _p(cnt), _p(idx), _p(qcnt), _p(out), 3) == 0
    assert np.array_equal(out, R.group_points_stack(feats, cnt, idx, qcnt))
    g_out = rng.standard_normal((M, 5, ns)).astype(np.float32)
    g = np.zeros_like(feats)
    assert emul.emul_group_points_stack_grad(2, M, 5, ns, _p(g_out), _p(idx), _p(qcnt), _p(cnt), _p(g), 3) == 0
    want = R.group_points_grad_stack(g_out, idx, qcnt, cnt, feats.shape[0])
    assert np.allclose(g, want, rtol=1e-5, atol=1e-5)               # float atomics: summation order only
    assert np.abs(want).max() > 0.1


# ---- reverse trilinear gather ----------------------------------------------------------------------------------------
def _sparse_source(rng, batch, shape, n, C):
    cells = batch * shape[0] * shape[1] * shape[2]
    flat = rng.choice(cells, size=n, replace=False)
    b, rem = np.divmod(flat, shape[0] * shape[1] * shape[2])
    z, rem = np.divmod(rem, shape[1] * shape[2])
    y, x = np.divmod(rem, shape[2])
    coords = np.stack([b, z, y, x], 1).astype(np.int32)
    feats = rng.standard_normal((n, C)).astype(np.float32)
    feats[rng.random((n, C)) < 0.3] = 0.0            # post-ReLU-like zeros
    feats[:3] = 0.0                                  # a few all-zero rows: active sites that contribute nothing
    return coords, feats


def _targets(rng, batch, shape, per_scene):
    T = batch * per_scene
    zyx = np.stack([rng.uniform(-1.5, shape[0] + 0.5, T), rng.uniform(-1.5, shape[1] + 0.5, T),
                    rng.uniform(-1.5, shape[2] + 0.5, T)], 1).astype(np.float32)
    zyx[::7] = np.floor(zyx[::7])                    # targets exactly on cell centres: four or more zero weights
    zyx[5] = [-3.0, 1.0, 1.0]                        # far outside
    return zyx


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("normalize", [False, True])
def test_trilinear_oracle_is_the_reference_function(normalize):
    """oracle.roi_pool.reverse_trilinear == common_utils.reverse_sparse_trilinear_interpolate_torch, bit for bit (CPU)."""
    from oracle import roi_pool as R
    cu = ref_loader.load_reference_modules()["common_utils"]
    rng = np.random.default_rng(11)
    batch, shape, C = 2, [2, 20, 17], 8
    coords, feats = _sparse_source(rng, batch, shape, 220, C)
    zyx = torch.from_numpy(_targets(rng, batch, shape, 96 * 6))
    b = torch.arange(zyx.shape[0]) // (96 * 6)
    ft, ct = torch.from_numpy(feats), torch.from_numpy(coords)

    class Feat(object):                               # what the reference function touches of a SparseConvTensor
        spatial_shape = shape

        @staticmethod
        def dense():
            return R.dense_volume(ft, ct, batch, shape)

    want = cu.reverse_sparse_trilinear_interpolate_torch(Feat, b, zyx, normalize=normalize)
    got = R.reverse_trilinear(ft, ct, batch, shape, b, zyx, normalize=normalize)
    assert torch.equal(got, want)
    assert int((want.abs().sum(1) > 0).sum()) > 50


@pytest.mark.parametrize("normalize", [False, True])
@pytest.mark.parametrize("batch,shape,C,lshape", [(2, [2, 20, 17], 8, [2, 4, 12]), (1, [3, 9, 8], 40, [1, 2, 3]), (3, [1, 6, 5], 33, [2, 2, 2])])
def test_emulated_trilinear_kernels_are_bit_exact(emul, normalize, batch, shape, C, lshape):
    from oracle import roi_pool as R
    rng = np.random.default_rng(batch * 10 + C)
    Pn = lshape[0] * lshape[1] * lshape[2]
    per_scene = Pn * (2 if Pn > 50 else 5)
    coords, feats = _sparse_source(rng, batch, shape, min(60 * batch, batch * shape[0] * shape[1] * shape[2] // 2), C)
    zyx = _targets(rng, batch, shape, per_scene)
    T = zyx.shape[0]
    want_c, want_f, want_t = R.interpolate_rows(torch.from_numpy(feats), torch.from_numpy(coords), batch, shape,
                                                torch.from_numpy(zyx), per_scene, lshape, normalize=normalize)
    n = want_f.shape[0]
    assert 10 < n < T
    for cap in (T, n, n - 4):
        out_f = np.full((max(cap, 1), C), np.nan, np.float32)
        out_c = np.full((max(cap, 1), 4), -9, np.int32)
        out_t = np.full((max(cap, 1),), -9, np.int64)
        total = emul.emul_trilinear_sparse(_p(feats), _p(coords), coords.shape[0], C, batch, _p(np.array(shape, np.int32)),
                                           _p(zyx), None, T, per_scene, int(normalize), Pn, _p(np.array(lshape, np.int32)),
                                           cap, _p(out_f), _p(out_c), _p(out_t), 4)
        assert total == n                                           # the count is the full count, rows beyond cap dropped
        k = min(cap, n)
        assert np.array_equal(out_f[:k], want_f.numpy()[:k])        # bit-exact rows (== ignores the sign of zero)
        assert np.array_equal(out_c[:k], want_c.numpy()[:k].astype(np.int32))
        assert np.array_equal(out_t[:k], want_t.numpy()[:k])
    # explicit per-target scenes give the same rows
    bt = (np.arange(T) // per_scene).astype(np.int64)
    out_f2 = np.zeros((n, C), np.float32)
    out_c = np.zeros((n, 4), np.int32)
    total = emul.emul_trilinear_sparse(_p(feats), _p(coords), coords.shape[0], C, batch, _p(np.array(shape, np.int32)), _p(zyx),
                                       _p(bt), T, 0, int(normalize), Pn, _p(np.array(lshape, np.int32)), n, _p(out_f2), _p(out_c),
                                       None, 2)
    assert total == n and np.array_equal(out_f2, want_f.numpy())


def test_emulated_trilinear_grad_is_the_adjoint(emul):
    from oracle import roi_pool as R
    rng = np.random.default_rng(3)
    batch, shape, C, lshape = 2, [2, 12, 11], 12, [2, 4, 12]
    Pn, per_scene = 96, 96 * 3
    coords, feats = _sparse_source(rng, batch, shape, 120, C)
    zyx = _targets(rng, batch, shape, per_scene)
    ft = torch.from_numpy(feats).clone().requires_grad_(True)
    _, rows, inds = R.interpolate_rows(ft, torch.from_numpy(coords), batch, shape, torch.from_numpy(zyx), per_scene, lshape)
    g_rows = torch.from_numpy(rng.standard_normal(tuple(rows.shape)).astype(np.float32))
    rows.backward(g_rows)
    want = ft.grad.numpy()
    got = np.zeros_like(feats)
    tgt = inds.numpy().astype(np.int64)
    assert emul.emul_trilinear_sparse_grad(_p(g_rows.numpy()), _p(tgt), len(tgt), C, _p(feats), _p(coords), coords.shape[0], batch,
                                           0, _p(np.array(shape, np.int32)), _p(zyx), None, zyx.shape[0], per_scene, 0, _p(got), 3) == 0
    assert np.allclose(got, want, rtol=1e-5, atol=1e-5) and np.abs(want).max() > 0.1


def test_drop_in_module_surface_and_cpu_rejection():
    """The drop-in exposes the reference extension's entry points (pointnet2_api.cpp:11-23, the three on the RoI path) and
    refuses CPU tensors like the reference's CHECK_INPUT; the fused host functions have no CPU path either."""
    from btcdet_b200 import pointnet2_stack_cuda as m, roi_pool
    for name in ("ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "ball_query_multi"):
        assert callable(getattr(m, name))
    z = torch.zeros(2, 3)
    c = torch.tensor([2], dtype=torch.int32)
    with pytest.raises(RuntimeError):
        m.ball_query_wrapper(1, 2, 0.5, 4, z, c, z, c, torch.zeros(2, 4, dtype=torch.int32))
    zyx = roi_pool.target_indices(torch.tensor([[[35.2, 0.0, -1.0]]]), [0, -40, -3, 70.4, 40, 1], [0.05, 0.05, 0.1], [8, 8, 8])
    assert zyx.shape == (1, 3) and torch.allclose(zyx, torch.tensor([[2.0, 99.5, 87.5]]), atol=1e-4)

    class T(object):
        features = torch.zeros(1, 4)

    with pytest.raises(RuntimeError):
        roi_pool.trilinear_gather_rows(T, zyx, 1, [1, 1, 1])


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_oracle_rows_are_the_reference_conv_head_method():
    """oracle.roi_pool.target_indices + interpolate_rows == ConvHead.create_local_conv_grid /
    interpolate_from_3d_features (conv_head.py:207-223, 505-528) of the reference's own class, constructed from the
    reference's yaml, on the synthetic RoI-head case (CPU): same rows, same coordinates, bit for bit."""
    from btcdet_b200 import synthetic as S
    from oracle import roi_pool as R
    mods = ref_loader.load_roi_head_modules()
    head = ref_loader.build_conv_head(mods, S.DET_VOXEL_SIZE, S.KITTI_RANGE)
    case = S.roi_head_case(batch=2, n_points=6000, n_rois=6, n_occ=100, channels=16)
    rois = torch.from_numpy(case["rois"])
    feats, coords, shape = torch.from_numpy(case["x_features"]), torch.from_numpy(case["x_coords"]), case["x_shape"]
    with ref_loader.cuda_as_cpu():
        grid_pts, _ = head.get_global_grid_points_of_roi(rois, grid_size=head.grid_size, e2e=False, dim_times=head.dim_times)
        sm = head.size_map["x_combine"]
        conv_pts, dense_idx = head.create_local_conv_grid(grid_pts.view(-1, 3), rois, sm["local_grid_size"], sm["dims"],
                                                          sm["scene_times"])

        class Feat(object):
            spatial_shape = shape
            indices = coords

            @staticmethod
            def dense():
                return R.dense_volume(feats, coords, 2, shape)

        want_c, want_f = head.interpolate_from_3d_features(conv_pts, dense_idx, Feat, head.downsample_times_map["x_combine"])
    assert tuple(conv_pts.shape) == (2, 6 * 27 * 96, 3) and tuple(dense_idx.shape) == (2 * 6 * 27, 96, 3)
    zyx = R.target_indices(conv_pts, S.KITTI_RANGE, S.DET_VOXEL_SIZE, [8, 8, 8])
    got_c, got_f, _ = R.interpolate_rows(feats, coords, 2, shape, zyx, conv_pts.shape[1], [2, 4, 12])
    assert want_f.shape[0] > 200
    assert torch.equal(got_f, want_f) and torch.equal(got_c.float(), want_c)
    # the product's host-side index expression is the same torch expression
    from btcdet_b200 import roi_pool
    assert torch.equal(roi_pool.target_indices(conv_pts, S.KITTI_RANGE, S.DET_VOXEL_SIZE, [8, 8, 8]), zyx)


# ---- randomised shapes (hypothesis): the emulated kernels against the oracle ------------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=30, deadline=None, derandomize=True)
@given(st.data())
def test_emulated_ball_query_random_ragged_batches(oracle, emul, data):
    from oracle import roi_pool as R
    B = data.draw(st.integers(1, 4))
    counts = [data.draw(st.integers(0, 90)) for _ in range(B)]
    qcounts = [data.draw(st.integers(0, 9)) for _ in range(B)]
    n_r = data.draw(st.integers(1, 4))
    radii = [data.draw(st.sampled_from([0.3, 0.5, 0.9, 1.7, 4.0])) for _ in range(n_r)]
    nsamples = [data.draw(st.sampled_from([1, 2, 7, 16, 33, 64])) for _ in range(n_r)]
    rng = np.random.default_rng(data.draw(st.integers(0, 10 ** 6)))
    xyz, cnt, q, qcnt = _scene(rng, counts, qcounts, spread=2.0)
    if sum(counts) == 0:
        xyz = np.zeros((1, 3), np.float32)               # a valid pointer; no scene has points
    M = q.shape[0]
    if M == 0:
        return
    outs = [np.full((M, ns), -77, np.int32) for ns in nsamples]
    ptrs = (ctypes.c_void_p * n_r)(*[o.ctypes.data for o in outs])
    assert emul.emul_ball_query_stack(B, M, n_r, _p(np.array(radii, np.float32)), _p(np.array(nsamples, np.int32)), _p(q),
                                      _p(qcnt), _p(xyz), _p(cnt), ptrs, data.draw(st.integers(1, 3))) == 0
    for r in range(n_r):
        assert np.array_equal(outs[r], R.ball_query_stack(radii[r], nsamples[r], xyz, cnt, q, qcnt)), (r, counts, qcounts)


@settings(max_examples=20, deadline=None, derandomize=True)
@given(st.data())
def test_emulated_trilinear_random_shapes(emul, data):
    from oracle import roi_pool as R
    batch = data.draw(st.integers(1, 3))
    shape = [data.draw(st.integers(1, 4)), data.draw(st.integers(1, 9)), data.draw(st.integers(1, 9))]
    C = data.draw(st.sampled_from([1, 3, 31, 32, 33, 70]))
    lshape = [data.draw(st.integers(1, 2)), data.draw(st.integers(1, 3)), data.draw(st.integers(1, 4))]
    normalize = data.draw(st.booleans())
    Pn = lshape[0] * lshape[1] * lshape[2]
    per_scene = Pn * data.draw(st.integers(1, 4))
    rng = np.random.default_rng(data.draw(st.integers(0, 10 ** 6)))
    cells = batch * shape[0] * shape[1] * shape[2]
    n = data.draw(st.integers(0, min(cells, 40)))
    coords, feats = _sparse_source(rng, batch, shape, n, C) if n else (np.zeros((0, 4), np.int32), np.zeros((0, C), np.float32))
    T = batch * per_scene
    zyx = np.stack([rng.uniform(-1.5, shape[0] + 0.5, T), rng.uniform(-1.5, shape[1] + 0.5, T),
                    rng.uniform(-1.5, shape[2] + 0.5, T)], 1).astype(np.float32)
    zyx[::5] = np.floor(zyx[::5])
    want_c, want_f, want_t = R.interpolate_rows(torch.from_numpy(feats), torch.from_numpy(coords), batch, shape,
                                                torch.from_numpy(zyx), per_scene, lshape, normalize=normalize)
    k = want_f.shape[0]
    out_f = np.full((T, C), np.nan, np.float32)
    out_c = np.full((T, 4), -9, np.int32)
    out_t = np.full((T,), -9, np.int64)
    src_f = feats if n else np.zeros((1, C), np.float32)
    src_c = coords if n else np.zeros((1, 4), np.int32)
    total = emul.emul_trilinear_sparse(_p(src_f), _p(src_c), n, C, batch, _p(np.array(shape, np.int32)), _p(zyx), None, T,
                                       per_scene, int(normalize), Pn, _p(np.array(lshape, np.int32)), T, _p(out_f), _p(out_c),
                                       _p(out_t), 2)
    assert total == k
    assert np.array_equal(out_f[:k], want_f.numpy()) and np.array_equal(out_c[:k], want_c.numpy().astype(np.int32))
    assert np.array_equal(out_t[:k], want_t.numpy())


def test_oracle_and_emulated_kernel_against_the_reference_generated_fixture(emul):
    """tests/golden/roi_pool_ref.npz = the reference's ConvHead (built from its yaml) run on the CPU by
    tests/golden/make_roi_pool_golden.py.  Runs without a reference checkout: the oracle AND the kernel source under the
    emulation reproduce the reference's coordinates and rows bit for bit from the reference's own grid points."""
    from btcdet_b200 import synthetic as S
    from oracle import roi_pool as R
    g = np.load(os.path.join(HERE, "golden", "roi_pool_ref.npz"))
    conv_pts = torch.from_numpy(g["conv_grid_points"])
    shape, lshape, stride = g["x_shape"].tolist(), g["local_grid_size"].tolist(), g["stride"].tolist()
    feats, coords = np.ascontiguousarray(g["x_features"]), np.ascontiguousarray(g["x_coords"])
    zyx = R.target_indices(conv_pts, S.KITTI_RANGE, S.DET_VOXEL_SIZE, stride)
    per_scene = conv_pts.shape[1]
    got_c, got_f, _ = R.interpolate_rows(torch.from_numpy(feats), torch.from_numpy(coords), 2, shape, zyx, per_scene, lshape)
    assert got_f.shape[0] == g["out_features"].shape[0] > 500
    assert np.array_equal(got_f.numpy(), g["out_features"]) and np.array_equal(got_c.float().numpy(), g["out_coords"])
    T, C = zyx.shape[0], feats.shape[1]
    n = got_f.shape[0]
    out_f, out_c = np.zeros((n, C), np.float32), np.zeros((n, 4), np.int32)
    zyx_np = np.ascontiguousarray(zyx.numpy())
    total = emul.emul_trilinear_sparse(_p(feats), _p(coords), coords.shape[0], C, 2, _p(np.array(shape, np.int32)), _p(zyx_np),
                                       None, T, per_scene, 0, lshape[0] * lshape[1] * lshape[2],
                                       _p(np.array(lshape, np.int32)), n, _p(out_f), _p(out_c), None, 4)
    assert total == n
    assert np.array_equal(out_f, g["out_features"]) and np.array_equal(out_c.astype(np.float32), g["out_coords"])
