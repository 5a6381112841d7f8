"""The GPU parity tests' own test functions (tests/test_parity_gpu.py) executed on the CPU against the KERNEL SOURCE of
btcdet_b200/csrc/{voxelize,coord_index,rulebook,pool_dense,points_transform,roi_pool}.cu compiled for the host under the
lock-step emulation of tests/host_emul/ (tests/host_emul/build_emul.py: the real `btc_*` C entry points, numpy / CPU-torch
buffers standing in for device memory).  TEST INFRASTRUCTURE: the product has no CPU path — this module swaps the loaded
library, the stream getter and the is-CUDA check of `btcdet_b200.ops` for the duration of a test and puts them back.
Only the small cases run here (one OS thread per CUDA thread); the full sizes run on the B200 (`-m gpu`)."""
import contextlib
import ctypes
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(HERE, "host_emul"))
import ref_loader  # noqa: E402  (cuda_as_cpu: device="cuda" -> "cpu", Tensor.cuda() -> identity)


@pytest.fixture(scope="module")
def emul_lib():
    import build_emul
    from btcdet_b200 import _lib
    lib = ctypes.CDLL(build_emul.build())
    for name, (res, args) in _lib.SIGNATURES.items():
        if hasattr(lib, name):            # the tcgen05 / FFMA / mask files are not part of the emulated build
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


@contextlib.contextmanager
def emulated(lib):
    from btcdet_b200 import _lib, ops
    saved = (_lib._lib, ops._stream, ops._require_cuda, ref_loader.ON_CUDA)
    _lib._lib, ops._stream, ops._require_cuda, ref_loader.ON_CUDA = lib, (lambda: None), (lambda *a, **k: None), False
    try:
        with ref_loader.cuda_as_cpu():
            yield
    finally:
        _lib._lib, ops._stream, ops._require_cuda, ref_loader.ON_CUDA = saved


def test_voxeliser_kernels_bit_exact_under_emulation(oracle, emul_lib):
    import tests.test_parity_gpu as G
    with emulated(emul_lib):
        G.test_voxelize_edge_cases(None, oracle)
        for case in ("config1_uniform2k", "lidar20k", "batch3", "cap_voxels", "occ_grid"):     # the GPU test's full sizes
            G.test_voxelize_bit_exact(None, oracle, case)


def test_rulebook_kernels_bit_exact_under_emulation(oracle, emul_lib):
    import tests.test_parity_gpu as G
    with emulated(emul_lib):
        G.test_rulebook_empty_and_single(None, oracle)
        G.test_rulebook_sort_rows(None, 127, 27)
        G.test_rulebook_sort_rows(None, 2049, 8)


def test_rulebook_pyramids_bit_exact_under_emulation(oracle, emul_lib):
    """The det pyramid's geometries (sub-manifold tables through the hash and through the rank bitmap, k3 s2 p1, k3 s2
    p(0,1,1), k(3,1,1) s(2,1,1), strided output index reused by the next sub-manifold layer) and the occupancy
    backbone's (dilating k3 s1 p1, two transposed convolutions) on small grids: output order, neighbour tables and the
    spconv-format pair lists against the oracle, bit for bit."""
    import tests.test_parity_gpu as G
    from btcdet_b200 import ops
    rng = np.random.default_rng(4)
    with emulated(emul_lib):
        batch, shape = 2, [21, 48, 40]
        flat = rng.choice(batch * int(np.prod(shape)), 450, replace=False)
        coords = np.stack([flat // int(np.prod(shape)), (flat // (shape[1] * shape[2])) % shape[0],
                           (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
        for lvl, (ksize, stride, pad) in enumerate([(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)]):
            c = torch.from_numpy(coords)
            rb = ops.rulebook_subm(c, batch, shape, 3)
            G._check_rulebook(oracle, rb, coords, batch, shape, 3, 1, 1, True, False)
            if lvl == 0:
                rb_bitmap = ops.rulebook_subm(c, batch, shape, 3, index=ops.build_index(c, batch, shape, need_perm=True))
                assert torch.equal(rb_bitmap.nbr_out, rb.nbr_out)
            rb = ops.rulebook_conv(c, batch, shape, ksize, stride, pad)
            coords = G._check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, pad, False, False)
            rb2 = ops.rulebook_subm(rb.out_coords, batch, rb.out_shape, 3, index=rb.out_index)
            G._check_rulebook(oracle, rb2, coords, batch, rb.out_shape, 3, 1, 1, True, False)
            shape = rb.out_shape
        batch, shape = 2, [9, 21, 25]
        flat = rng.choice(batch * int(np.prod(shape)), 300, replace=False)
        coords = np.stack([flat // int(np.prod(shape)), (flat // (shape[1] * shape[2])) % shape[0],
                           (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
        for ksize, stride, pad, tr in [(3, 1, 1, False), (3, 2, 1, False), (3, 2, 1, False), (3, 2, 1, True), (3, 2, 1, True)]:
            rb = ops.rulebook_conv(torch.from_numpy(coords), batch, shape, ksize, stride, pad, transposed=tr)
            coords = G._check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, pad, False, tr)
            shape = rb.out_shape
        assert shape == [9, 21, 25]


def test_pool_dense_revoxelise_kernels_under_emulation(oracle, emul_lib):
    """SparseMaxPool3d forward / backward (zero-init quirk), dense() / its backward, and the sorted re-voxelisation
    (torch.unique(dim=0) + stable slots) on small inputs."""
    import tests.test_parity_gpu as G
    from btcdet_b200 import ops
    rng = np.random.default_rng(9)
    with emulated(emul_lib):
        batch, shape = 2, [9, 24, 20]
        flat = rng.choice(batch * int(np.prod(shape)), 900, replace=False)
        coords = np.stack([flat // int(np.prod(shape)), (flat // (shape[1] * shape[2])) % shape[0],
                           (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
        feat = rng.standard_normal((coords.shape[0], 3)).astype(np.float32)
        outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, batch, shape, 3, 2, 1)
        rb = ops.rulebook_conv(torch.from_numpy(coords), batch, shape, 3, 2, 1)
        f = torch.from_numpy(feat).requires_grad_()
        out = ops.SparseMaxPoolFunction.apply(f, rb)
        np.testing.assert_array_equal(out.detach().numpy(), oracle.indice_maxpool(feat, pairs, pair_num, outids.shape[0]))
        out.sum().backward()
        assert np.all(f.grad.numpy() >= 0) and f.grad.sum() > 0
        fd = torch.from_numpy(feat).requires_grad_()
        d = ops.ToDenseFunction.apply(fd, torch.from_numpy(coords), batch, shape)
        np.testing.assert_array_equal(d.detach().numpy(), oracle.dense(feat, coords, shape, batch))
        wgt = torch.randn_like(d)
        (d * wgt).sum().backward()
        li = torch.from_numpy(coords).long()
        np.testing.assert_array_equal(fd.grad.numpy(), wgt[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]].numpy())
        # re-voxelisation
        g = torch.Generator().manual_seed(0)
        base = torch.stack([torch.randint(0, 2, (300,), generator=g), torch.randint(0, 9, (300,), generator=g),
                            torch.randint(0, 24, (300,), generator=g), torch.randint(0, 20, (300,), generator=g)], 1)
        pc = torch.cat([base, base[:120], base[:30]], 0)
        pc = pc[torch.randperm(pc.shape[0], generator=g)]
        pf = torch.randn(pc.shape[0], 6, generator=g)
        vox, cnt, vc = ops.revoxelize_sorted(pc, pf, 2, shape)
        u, inv, counts = torch.unique(pc, dim=0, sorted=True, return_inverse=True, return_counts=True)
        assert torch.equal(vc, u) and torch.equal(cnt, counts) and vox.shape == (u.shape[0], int(counts.max()), 6)
        ref, seen = torch.zeros_like(vox), {}
        for i in range(pc.shape[0]):
            v = int(inv[i])
            ref[v, seen.get(v, 0)] = pf[i]
            seen[v] = seen.get(v, 0) + 1
        assert torch.equal(vox, ref)


def test_transform_conv_golden_and_iou_kernels_under_emulation(oracle, emul_lib):
    """a2 coordinate transform (numpy's op order; angles to the GPU test's 4 ulp), BASELINE configs[0] against the
    committed golden fixture (voxelise -> SubM rulebook -> fp32 FFMA gather-GEMM), a strided FFMA convolution with bias /
    folded affine / ReLU, its backward against autograd, and the rotated BEV IoU against the float64 oracle."""
    import tests.test_iou3d_gpu as GI
    import tests.test_parity_gpu as G
    import tests.test_points_transform_gpu as GP
    with emulated(emul_lib):
        GP.test_points_to_cylinder_and_sphere(None, 0, 20000)
        G.test_config1_golden(None, oracle)
        G.test_sparse_conv_forward_within_tolerance(None, oracle, 6, 16, "conv")
    # rotated BEV IoU / overlap straight through the C entry point (tolerances of tests/test_iou3d_gpu.py)
    from oracle import iou3d
    a, b = GI.random_boxes(60, 3), GI.random_boxes(50, 4)
    a[:, 0] += 40.0
    b[:, 0] += 40.0
    for mode, tol in ((0, 2e-4), (1, 2e-3)):
        out = np.zeros((60, 50), np.float32)
        assert emul_lib.btc_boxes_bev(a.ctypes.data, 60, b.ctypes.data, 50, mode, out.ctypes.data, None) == 0
        want = iou3d.boxes_bev(a, b, bool(mode))
        assert np.abs(out - want).max() < tol and (want > 0).sum() > 50
    # NMS: 64 x 64 mask blocks + the greedy scan by one warp, keep list and count in "device" memory
    bx = GI._gap_boxes(150, 7, 0.3)
    want = iou3d.greedy_nms(iou3d.boxes_bev(bx, bx), 0.3)
    keep, num = np.full(len(bx), -1, np.int64), np.zeros(1, np.int32)
    ws = np.zeros(int(emul_lib.btc_nms_workspace_bytes(len(bx))) + 64, np.uint8)
    assert emul_lib.btc_nms(bx.ctypes.data, len(bx), 0.3, 0, keep.ctypes.data, num.ctypes.data, ws.ctypes.data, ws.size, None) == 0
    assert 5 < len(want) < len(bx) and int(num[0]) == len(want) and keep[:len(want)].tolist() == want.tolist()


# ---- randomised geometries (hypothesis): rulebook kernels against the oracle -----------------------------------------
from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=16, deadline=None, derandomize=True)
@given(st.data())
def test_rulebooks_random_geometries_under_emulation(oracle, emul_lib, data):
    """Kernel sizes 1..3, strides 1..3, paddings 0..2, dilation with stride 1, regular / transposed / sub-manifold, on
    small random grids and batches: output coordinates (ascending flat key), both neighbour tables and the spconv-format
    pair lists equal the oracle's (the straight-line fast paths and the general path are both reached)."""
    import tests.test_parity_gpu as G
    from btcdet_b200 import ops
    batch = data.draw(st.integers(1, 3))
    shape = [data.draw(st.integers(2, 9)), data.draw(st.integers(3, 14)), data.draw(st.integers(3, 14))]
    cells = batch * shape[0] * shape[1] * shape[2]
    n = data.draw(st.integers(1, min(cells, 120)))
    rng = np.random.default_rng(data.draw(st.integers(0, 10 ** 6)))
    flat = rng.choice(cells, n, replace=False)
    if data.draw(st.booleans()):
        flat = np.sort(flat)
    coords = np.stack([flat // (shape[0] * shape[1] * shape[2]), (flat // (shape[1] * shape[2])) % shape[0],
                       (flat // shape[2]) % shape[1], flat % shape[2]], 1).astype(np.int32)
    kind = data.draw(st.sampled_from(["conv", "transposed", "subm"]))
    with emulated(emul_lib):
        c = torch.from_numpy(coords)
        if kind == "subm":
            ksize = [data.draw(st.sampled_from([1, 3])) for _ in range(3)]
            rb = ops.rulebook_subm(c, batch, shape, ksize)
            G._check_rulebook(oracle, rb, coords, batch, shape, ksize, 1, [k // 2 for k in ksize], True, False)
            return
        ksize = [data.draw(st.integers(1, 3)) for _ in range(3)]
        stride = [data.draw(st.integers(1, 3)) for _ in range(3)]
        pad = [data.draw(st.integers(0, min(2, k - 1))) for k in ksize]
        tr = kind == "transposed"
        out_shape = (ops.deconv_output_shape(shape, ksize, stride, pad, [1, 1, 1], [0, 0, 0]) if tr
                     else ops.conv_output_shape(shape, ksize, stride, pad, [1, 1, 1]))
        if min(out_shape) <= 0 or batch * out_shape[0] * out_shape[1] * out_shape[2] > 60000:
            return
        rb = ops.rulebook_conv(c, batch, shape, ksize, stride, pad, transposed=tr)
        G._check_rulebook(oracle, rb, coords, batch, shape, ksize, stride, pad, False, tr)


@settings(max_examples=12, deadline=None, derandomize=True)
@given(st.data())
def test_voxeliser_random_scenes_under_emulation(oracle, emul_lib, data):
    """Random small scenes (empty scenes, points outside the range, both caps biting): voxel order, coordinates, counts
    and the point rows equal the sequential oracle's, bit for bit."""
    import tests.test_parity_gpu as G
    n_scenes = data.draw(st.integers(1, 3))
    rng = np.random.default_rng(data.draw(st.integers(0, 10 ** 6)))
    vs = [data.draw(st.sampled_from([0.25, 0.5, 1.0])) for _ in range(3)]
    rg = [0.0, -2.0, -1.0, 4.0, 2.0, 1.0]
    mp, mv = data.draw(st.integers(1, 4)), data.draw(st.sampled_from([3, 20, 400]))
    scenes = []
    for _ in range(n_scenes):
        n = data.draw(st.integers(0, 300))
        p = rng.uniform([-0.5, -2.5, -1.2, 0.0], [4.5, 2.5, 1.2, 1.0], (n, 4)).astype(np.float32)
        if n and data.draw(st.booleans()):
            p[: n // 2, :3] = np.round(p[: n // 2, :3] * 4) / 4          # points exactly on voxel edges
        scenes.append(p)
    with emulated(emul_lib):
        ov, oc, on = oracle.voxelize_batch(scenes, vs, rg, mp, mv)
        gv, gc, gn, gmean, nv = G._gpu_voxelize(scenes, vs, rg, mp, mv, want_mean=True)
    assert gc.shape == oc.shape
    np.testing.assert_array_equal(gc, oc)
    np.testing.assert_array_equal(gn, on)
    np.testing.assert_array_equal(gv, ov)


def test_mask_kernels_under_emulation(oracle, emul_lib):
    """Occupancy / occlusion masks (rows a5-a8, a12) and the box-driven targets (a9-a11) on the host: the border-clamp known
    answer, the fixture generated by the reference's own OccTargets3D on the CPU (integer-only masks exact, the occluded
    set within the bound the GPU test uses for that cross-device fixture — the host's libm is one more libm), exact loss-map
    algebra, the edge cases, and the box targets against the CPU oracle."""
    import tests.test_box_masks_gpu as GB
    import tests.test_occ_gpu as GO
    with emulated(emul_lib):
        GO.test_occ_masks_empty_and_edge(None)
        GO.test_occ_masks_vs_reference_fixture(None, oracle)
        GB.test_end_to_end_targets_close_to_reference_fixture(None, oracle)
        GB.test_loss_maps_exact(None, oracle)
        GB.test_box_targets_edge_cases(None, oracle)
        GB.test_box_targets_match_oracle_on_device(None, oracle, [11], 6000, False, False)


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_sync_free_proposal_layer_equals_the_reference_method_under_emulation(emul_lib):
    """btcdet_b200.proposal.proposal_layer against `RoIHeadTemplate.proposal_layer` (roi_head_template.py:46-101) of the
    reference's own ConvHead instance, both on the emulated library: identical rois / scores / labels, with tied scores,
    a scene with fewer kept boxes than NMS_POST_MAXSIZE and both NMS types."""
    import tests.test_iou3d_gpu as GI
    from btcdet_b200 import iou3d_nms_cuda as iou, proposal, synthetic as S
    mods = ref_loader.load_roi_head_modules()
    head = ref_loader.build_conv_head(mods, S.DET_VOXEL_SIZE, S.KITTI_RANGE)
    rng = np.random.default_rng(12)
    boxes = np.stack([GI._gap_boxes(260, 20 + b, 0.7)[:200] for b in range(2)])           # clustered: NMS has work
    scores = np.round(rng.uniform(0, 1, (2, 200, 1)), 2).astype(np.float32)               # two decimals: many ties
    saved = (iou._check_boxes, iou._stream)
    iou._check_boxes, iou._stream = (lambda *a: None), (lambda: None)
    try:
        with emulated(emul_lib):
            for nms_type, pre, post, thresh in (("nms_gpu", 150, 40, 0.7), ("nms_gpu", 4096, 300, 0.1), ("nms_normal_gpu", 100, 16, 0.5)):
                cfg = ref_loader.Cfg({"NMS_TYPE": nms_type, "MULTI_CLASSES_NMS": False, "NMS_PRE_MAXSIZE": pre,
                                      "NMS_POST_MAXSIZE": post, "NMS_THRESH": thresh})
                mk = lambda: {"batch_size": 2, "batch_box_preds": torch.from_numpy(boxes.copy()),   # noqa: E731
                              "batch_cls_preds": torch.from_numpy(scores.copy())}
                want = head.proposal_layer(mk(), nms_config=cfg)
                got = proposal.proposal_layer(mk(), cfg)
                for k in ("rois", "roi_scores", "roi_labels"):
                    assert torch.equal(got[k], want[k]), (nms_type, k)
                assert got["has_class_labels"] == want["has_class_labels"]
                kept = int((want["roi_scores"] > 0).sum())
                assert 10 < kept <= 2 * post
    finally:
        iou._check_boxes, iou._stream = saved


def test_proposal_layer_against_the_oracle_nms_under_emulation(emul_lib):
    import tests.test_zz_proposal_gpu as GZ
    from btcdet_b200 import iou3d_nms_cuda as iou
    saved = (iou._check_boxes, iou._stream)
    iou._check_boxes, iou._stream = (lambda *a: None), (lambda: None)
    try:
        with emulated(emul_lib):
            GZ.test_proposal_layer_against_the_oracle_nms(None)
    finally:
        iou._check_boxes, iou._stream = saved


def test_ffma_conv_backward_under_emulation(oracle, emul_lib):
    """dX / dW / db of the fp32 kernels (FFMA gather-GEMM over the mirrored table, slab-split outer products + atomics,
    column sums) against autograd through the oracle formulation, on a small scene."""
    import tests.test_parity_gpu as G
    from btcdet_b200 import ops
    rng = np.random.default_rng(0)
    batch, shape = 1, [11, 40, 36]
    flat = rng.choice(int(np.prod(shape)), 500, replace=False)
    coords = np.stack([np.zeros_like(flat), flat // (shape[1] * shape[2]), (flat // shape[2]) % shape[1], flat % shape[2]],
                      1).astype(np.int32)
    with emulated(emul_lib):
        for kind, cin, cout in [("subm", 16, 32), ("conv", 6, 20)]:
            if kind == "subm":
                outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, subm=True)
                rb = ops.rulebook_subm(torch.from_numpy(coords), 1, shape, 3)
            else:
                outids, pairs, pair_num, _ = oracle.get_indice_pairs(coords, 1, shape, 3, 2, 1)
                rb = ops.rulebook_conv(torch.from_numpy(coords), 1, shape, 3, 2, 1)
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
            f2, w2, b2 = (t.clone().requires_grad_() for t in (feat, w, b))
            y = ops.SparseConvFunction.apply(f2, w2, b2, rb, 1)
            assert G.rel_err(y.detach().numpy(), (out + b1).detach().numpy()) < G.REL_TOL
            y.backward(go)
            assert G.rel_err(f2.grad.numpy(), f1.grad.numpy()) < G.REL_TOL
            assert G.rel_err(w2.grad.numpy(), w1.grad.numpy()) < G.REL_TOL
            assert G.rel_err(b2.grad.numpy(), b1.grad.numpy()) < G.REL_TOL



@pytest.mark.skipif(not os.environ.get("BTC_EMUL_FULL"), reason="minutes of CPU time: set BTC_EMUL_FULL=1")
def test_full_size_rulebook_pyramids_under_emulation(oracle, emul_lib):
    """The GPU tests' full-size rulebook cases (20k-point scene on the KITTI grid [41,1600,1408], both backbones' pyramids,
    max-pool / dense, re-voxelisation) on the emulated library: ~5 min."""
    import tests.test_parity_gpu as G
    with emulated(emul_lib):
        G.test_rulebooks_det_pyramid_bit_exact(None, oracle, 1, False)
        G.test_rulebooks_occ_pyramid_with_transposed_bit_exact(None, oracle)
        G.test_maxpool_and_dense(None, oracle)
        G.test_revoxelize_sorted_matches_torch_unique(None)


def test_trilinear_entry_points_under_emulation(emul_lib):
    """btc_trilinear_sparse_flag / _emit / _grad through the emulated library (workspace layout, the device-wide flag scan,
    the count, capacity clamp) against the torch oracle."""
    from oracle import roi_pool as R
    import tests.test_roi_pool_cpu as TR
    rng = np.random.default_rng(21)
    batch, shape, C, lshape, per_scene = 2, [2, 20, 17], 8, [2, 4, 12], 96 * 4
    coords, feats = TR._sparse_source(rng, batch, shape, 150, C)
    zyx = TR._targets(rng, batch, shape, per_scene)
    T = zyx.shape[0]
    want_c, want_f, want_t = R.interpolate_rows(torch.from_numpy(feats), torch.from_numpy(coords), batch, shape,
                                                torch.from_numpy(zyx), per_scene, lshape)
    n = want_f.shape[0]
    shp, lsh = (ctypes.c_int * 3)(*shape), (ctypes.c_int * 3)(*lshape)
    ws_bytes = int(emul_lib.btc_trilinear_sparse_workspace_bytes(T, batch, shp))
    ws = np.zeros(ws_bytes + 256, np.uint8)
    wp = (ws.ctypes.data + 255) & ~255
    count = np.full(1, -1, np.int32)
    assert emul_lib.btc_trilinear_sparse_flag(feats.ctypes.data, coords.ctypes.data, coords.shape[0], None, C, batch, shp,
                                              zyx.ctypes.data, None, T, per_scene, 0, count.ctypes.data, wp, ws_bytes, None) == 0
    assert int(count[0]) == n > 20
    for cap in (n, n - 7):
        out_f, out_c = np.zeros((cap, C), np.float32), np.zeros((cap, 4), np.int32)
        out_t = np.full(cap, -1, np.int64)
        assert emul_lib.btc_trilinear_sparse_emit(feats.ctypes.data, C, batch, shp, zyx.ctypes.data, None, T, per_scene, 0, 96, lsh,
                                                  cap, out_f.ctypes.data, out_c.ctypes.data, out_t.ctypes.data, wp, ws_bytes, None) == 0
        assert np.array_equal(out_f, want_f.numpy()[:cap]) and np.array_equal(out_c, want_c.numpy()[:cap].astype(np.int32))
        assert np.array_equal(out_t, want_t.numpy()[:cap])
    g_rows = rng.standard_normal((n, C)).astype(np.float32)
    grad = np.zeros_like(feats)
    assert emul_lib.btc_trilinear_sparse_grad(g_rows.ctypes.data, out_t.ctypes.data if cap == n else want_t.numpy().ctypes.data, n,
                                              None, C, batch, shp, zyx.ctypes.data, None, T, per_scene, 0, grad.ctypes.data, wp,
                                              ws_bytes, None) == 0
    ft = torch.from_numpy(feats).clone().requires_grad_(True)
    _, rows, _ = R.interpolate_rows(ft, torch.from_numpy(coords), batch, shape, torch.from_numpy(zyx), per_scene, lshape)
    rows.backward(torch.from_numpy(g_rows))
    assert np.allclose(grad, ft.grad.numpy(), rtol=1e-5, atol=1e-5)


# ---- the spconv shim's host logic on the emulated library ------------------------------------------------------------
@contextlib.contextmanager
def shim_on_host(lib):
    """As `emulated`, plus: tensors report is_cuda (the shim and the ops refuse CPU tensors — there is no CPU path — so the
    test makes the host buffers look like device buffers for the duration of the block)."""
    torch.Tensor.is_cuda = property(lambda self: True)
    try:
        with emulated(lib):
            yield
    finally:
        del torch.Tensor.is_cuda


def test_spconv_shim_backbone_on_the_emulated_library(oracle, emul_lib):
    """`backbones.VoxelBackBone8x` through the `spconv` package (SparseSequential, SubM / strided layers sharing indice
    keys, BatchNorm + ReLU in between) on two small scenes: every level's indices exact, features within 1e-4 of the
    oracle network.  (The emulated build has no tcgen05 tile: every layer takes the fp32 FFMA tile.)"""
    import tests.test_backbone_gpu as GB
    from btcdet_b200 import backbones, synthetic as S
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    scenes = [S.lidar_like(1000, seed=100 + b) for b in range(2)]
    v, c, npts = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    mean = (v.sum(1) / np.maximum(npts, 1)[:, None]).astype(np.float32)
    ref = GB._oracle_forward(model, mean, c, 2)
    with shim_on_host(emul_lib), torch.no_grad():
        out = model({"voxel_features": torch.from_numpy(mean), "voxel_coords": torch.from_numpy(c), "batch_size": 2})
    got = dict(out["multi_scale_3d_features"], out=out["encoded_spconv_tensor"])
    for name, r in ref.items():
        g = got[name]
        assert list(g.spatial_shape) == r.spatial_shape, name
        np.testing.assert_array_equal(g.indices.numpy(), r.indices, err_msg=name)
        assert GB.rel_err(g.features.numpy(), r.features) < GB.REL_TOL, name
    assert got["out"].spatial_shape == [2, 200, 176] and got["out"].features.shape[0] > 50


@pytest.mark.skipif(not os.environ.get("BTC_EMUL_FULL"), reason="~2 min of CPU time: set BTC_EMUL_FULL=1")
def test_reference_topologies_on_the_emulated_library(oracle, emul_lib):
    """The dataflow of the reference's two backbones through the shim (tests/models_mirror.py): VoxelBackBone8xOcc —
    SparseMaxPool3d side channel, sparse_cat, cached 'spconv3' / 'spconv4' rulebooks, dense() + gather — and
    VoxelBackBoneDeconv + the occupancy head — dilating SparseConv3d, two SparseConvTranspose3d, SubM heads, dense()."""
    import spconv
    import tests.test_backbone_gpu as GB
    from btcdet_b200 import synthetic as S
    from oracle.occ_masks import OccGeometry
    from tests import models_mirror, oracle_net
    batch = 2
    model = GB._randomize(models_mirror.DetBackboneMirror(6, 4), 3)
    scenes = [S.lidar_like(900, seed=300 + b) for b in range(batch)]
    v, coords, npts = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    rng = np.random.default_rng(0)
    feats = np.concatenate([(v.sum(1) / np.maximum(npts, 1)[:, None]), rng.uniform(0, 1, (v.shape[0], 2))], 1).astype(np.float32)
    occ_feats = np.abs(rng.standard_normal((v.shape[0], 2))).astype(np.float32)
    ref = model.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, model.sparse_shape, batch), occ_feats)
    with shim_on_host(emul_lib), torch.no_grad():
        x = spconv.SparseConvTensor(torch.from_numpy(feats), torch.from_numpy(coords), model.sparse_shape, batch)
        got = model.run(models_mirror.ShimBackend(), x, torch.from_numpy(occ_feats))
    for name in ("x_conv2", "x_conv3", "out", "x_combine"):
        np.testing.assert_array_equal(got[name].indices.numpy(), ref[name].indices, err_msg=name)
        assert GB.rel_err(got[name].features.numpy(), ref[name].features) < GB.REL_TOL, name
    assert got["x_combine"].features.shape[1] == 128 and list(got["out"].spatial_shape) == [2, 200, 176]
    # occupancy backbone + head
    geo = OccGeometry()
    model = GB._randomize(models_mirror.OccBackboneMirror(4), 5)
    gen = oracle.VoxelGeneratorV2(geo.voxel_size, geo.point_cloud_range, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"])
    fs, cs = [], []
    for b in range(batch):
        pts = S.lidar_like(500, seed=400 + b)
        cyl = np.stack([np.linalg.norm(pts[:, :2], axis=1), np.arctan2(-pts[:, 1], pts[:, 0]) * 180. / np.pi, pts[:, 2],
                        pts[:, 3]], -1).astype(np.float32)
        r = gen.generate(cyl)
        fs.append(r["voxels"].sum(1) / np.maximum(r["num_points_per_voxel"], 1)[:, None])
        cs.append(np.pad(r["coordinates"], ((0, 0), (1, 0)), constant_values=b))
    feats, coords = np.concatenate(fs).astype(np.float32), np.concatenate(cs).astype(np.int32)
    feats = feats / np.array([70.0, 40.0, 3.0, 1.0], np.float32)
    ref = model.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, model.sparse_shape, batch))
    with shim_on_host(emul_lib), torch.no_grad():
        x = spconv.SparseConvTensor(torch.from_numpy(feats), torch.from_numpy(coords), model.sparse_shape, batch)
        got = model.run(models_mirror.ShimBackend(), x)
        dense_cls = got["cls"].dense()
    for name in ("encoded", "cls", "res"):
        np.testing.assert_array_equal(got[name].indices.numpy(), ref[name].indices, err_msg=name)
        assert GB.rel_err(got[name].features.numpy(), ref[name].features) < GB.REL_TOL, name
    assert list(got["encoded"].spatial_shape) == [9, 157, 209]
    np.testing.assert_array_equal(dense_cls.numpy(), oracle.dense(got["cls"].features.numpy(), ref["cls"].indices, [9, 157, 209], batch))


def test_bench_roi_pool_leg_runs_on_the_emulated_library(emul_lib, monkeypatch):
    """tools/bench_legs.py::roi_pool_leg end to end (both stacked ball queries, the fused gather, the three mini-grid
    convolutions through the shim, dense()) at a tiny size: the leg's host code, which only ever runs inside bench.py on a
    GPU box, checked where the CPU suite runs."""
    import time
    from btcdet_b200 import pointnet2_stack_cuda as p2, roi_pool
    from tools import bench_legs

    class Ev(object):
        def record(self):
            self.t = time.time()

        def elapsed_time(self, other):
            return (other.t - self.t) * 1e3

    monkeypatch.setattr(bench_legs, "_events", lambda: (Ev(), Ev()))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(p2, "_stream", lambda: None)
    monkeypatch.setattr(roi_pool, "_stream", lambda: None)
    with shim_on_host(emul_lib):
        out = bench_legs.roi_pool_leg("cpu", steps=1, warmup=0, batch=2, n_rois=2)
    assert out["queries"] == 2 * 2 * 27 and out["targets"] == 2 * 2 * 27 * 96
    assert out["out_shape"] == [108, 128, 1, 1, 1] and out["gathered_rows"] > 0 and out["value"] > 0
