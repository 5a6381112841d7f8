"""SURVEY 8(f) N1 on the GPU: the RoI grid pooling ops (csrc/roi_pool_kernels.cuh) through the C ABI, against
  * the C / torch oracle (oracle/roi_pool.py; pinned to the reference's own code on the CPU, tests/test_roi_pool_cpu.py),
  * the REFERENCE'S OWN CUDA kernels compiled from the checkout (oracle/_ref/libpointnet2_ref.so), same inputs, same GPU,
  * the reference's `ConvHead.roi_conv_pool` run unchanged on the GPU (staged sources, O3) with and without the fused
    replacements (`btcdet_b200.roi_pool.patch_conv_head`).
Integer outputs and the interpolated rows are bit-exact; only float atomics (the two backward scatters) get a tolerance."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

RADII, NSAMPLES = [0.4, 0.8, 1.2, 2.4], [16, 16, 32, 64]     # btcdet_kitti_car.yaml:277-280 (raw_points)


def _queries(case, per_roi=27, seed=0):
    """Query points around the RoIs (stand-ins for the 3x3x3 grid points), stacked by scene."""
    rng = np.random.default_rng(seed)
    rois = case["rois"]
    q = np.repeat(rois[:, :, None, :3], per_roi, axis=2) + rng.normal(0.0, 0.7, rois.shape[:2] + (per_roi, 3))
    return q.reshape(-1, 3).astype(np.float32), np.full(rois.shape[0], rois.shape[1] * per_roi, np.int32)


def _scene_points(case):
    pts = case["points"]
    cnt = np.bincount(pts[:, 0].astype(np.int64), minlength=case["batch_size"]).astype(np.int32)
    return np.ascontiguousarray(pts[:, 1:4]), cnt


def _run_multi(xyz, cnt, q, qcnt, radii=RADII, nsamples=NSAMPLES):
    from btcdet_b200 import pointnet2_stack_cuda as ext
    dev = "cuda"
    t = [torch.from_numpy(a).to(dev) for a in (xyz, cnt, q, qcnt)]
    outs = [torch.full((q.shape[0], ns), -77, dtype=torch.int32, device=dev) for ns in nsamples]
    assert ext.ball_query_multi(radii, nsamples, t[2], t[3], t[0], t[1], outs) == 1
    return outs, t


def test_ball_query_is_bit_exact_with_oracle_and_reference_kernels(cuda, oracle):
    from btcdet_b200 import pointnet2_stack_cuda as ext, synthetic as S
    from oracle import roi_pool as R
    case = S.roi_head_case(batch=2, n_points=20000, n_rois=40)
    xyz, cnt = _scene_points(case)
    q, qcnt = _queries(case)
    q, qcnt = q[1:-1], qcnt - 1                                  # 1079 + 1079 queries: a group of four straddles the scenes
    outs, (xyz_t, cnt_t, q_t, qcnt_t) = _run_multi(xyz, cnt, q, qcnt)
    ref = R.RefPointnet2() if R.RefPointnet2.available() else None
    filled = 0
    for r, (radius, ns) in enumerate(zip(RADII, NSAMPLES)):
        want = R.ball_query_stack(radius, ns, xyz, cnt, q, qcnt)
        assert np.array_equal(outs[r].cpu().numpy(), want), radius
        one = torch.zeros((q.shape[0], ns), dtype=torch.int32, device="cuda")          # the reference's calling convention
        assert ext.ball_query_wrapper(2, q.shape[0], radius, ns, q_t, qcnt_t, xyz_t, cnt_t, one) == 1
        assert torch.equal(one, outs[r])
        if ref is not None:
            assert torch.equal(ref.ball_query(radius, ns, xyz_t, cnt_t, q_t, qcnt_t), outs[r]), radius
        filled += int((want[:, 0] >= 0).sum())
    assert filled > 2000
    # ragged: an empty scene in the middle, single queries
    xyz2 = np.concatenate([xyz[:cnt[0]], xyz[cnt[0]:cnt[0] + 777]])
    cnt2 = np.array([cnt[0], 0, 777], np.int32)
    q2, qcnt2 = q[:11], np.array([5, 3, 3], np.int32)
    outs2, t2 = _run_multi(xyz2, cnt2, q2, qcnt2, RADII[:3], NSAMPLES[:3])
    for r in range(3):
        assert np.array_equal(outs2[r].cpu().numpy(), R.ball_query_stack(RADII[r], NSAMPLES[r], xyz2, cnt2, q2, qcnt2))
        if ref is not None:
            assert torch.equal(ref.ball_query(RADII[r], NSAMPLES[r], t2[0], t2[1], t2[2], t2[3]), outs2[r])
    assert (outs2[0][5:8] == torch.tensor([-1] + [0] * 15, dtype=torch.int32, device="cuda")).all()    # empty scene


def test_group_points_and_grad_against_reference_kernels(cuda, oracle):
    from btcdet_b200 import pointnet2_stack_cuda as ext, synthetic as S
    from oracle import roi_pool as R
    case = S.roi_head_case(batch=2, n_points=8000, n_rois=16)
    xyz, cnt = _scene_points(case)
    q, qcnt = _queries(case)
    outs, (xyz_t, cnt_t, q_t, qcnt_t) = _run_multi(xyz, cnt, q, qcnt)
    idx = outs[2]
    idx[idx[:, 0] == -1] = 0
    M, ns = idx.shape
    feats = torch.randn((xyz.shape[0], 5), device="cuda")
    out = torch.full((M, 5, ns), float("nan"), device="cuda")
    assert ext.group_points_wrapper(2, M, 5, ns, feats, cnt_t, idx, qcnt_t, out) == 1
    want = R.group_points_stack(feats.cpu().numpy(), cnt, idx.cpu().numpy(), qcnt)
    assert np.array_equal(out.cpu().numpy(), want)
    g_out = torch.randn((M, 5, ns), device="cuda")
    g = torch.zeros_like(feats)
    assert ext.group_points_grad_wrapper(2, M, 5, feats.shape[0], ns, g_out, idx, qcnt_t, cnt_t, g) == 1
    want_g = R.group_points_grad_stack(g_out.cpu().numpy(), idx.cpu().numpy(), qcnt, cnt, feats.shape[0])
    assert np.allclose(g.cpu().numpy(), want_g, rtol=1e-4, atol=1e-4)
    if R.RefPointnet2.available():
        ref = R.RefPointnet2()
        assert torch.equal(ref.group_points(feats, cnt_t, idx, qcnt_t), out)
        assert torch.allclose(ref.group_points_grad(g_out, idx, qcnt_t, cnt_t, feats.shape[0]), g, rtol=1e-4, atol=1e-4)


def _sparse(case, channels=None):
    import spconv
    f = torch.from_numpy(case["x_features"]).cuda()
    if channels:
        f = f[:, :channels].contiguous()
    return spconv.SparseConvTensor(f, torch.from_numpy(case["x_coords"]).cuda(), case["x_shape"], case["batch_size"])


def _grid_targets(case, n_local=27, seed=1):
    """Targets like the head's: 96-cell mini grids (2 x 4 x 12 cells of 0.8 x 0.4 x 0.4 m) around the RoIs, as
    fractional (z, y, x) indices of the stride-8 feature level."""
    from btcdet_b200 import roi_pool, synthetic as S
    rng = np.random.default_rng(seed)
    rois = case["rois"]
    B, N = rois.shape[:2]
    centre = np.repeat(rois[:, :, None, :3], n_local, axis=2) + rng.normal(0.0, 0.6, (B, N, n_local, 3))
    lz, ly, lx = np.meshgrid(np.arange(2), np.arange(4), np.arange(12), indexing="ij")
    cell = np.stack([(lx.ravel() + 0.5) * 0.4 - 2.4, (ly.ravel() + 0.5) * 0.4 - 0.8, (lz.ravel() + 0.5) * 0.8 - 0.8], axis=1)
    pts = (centre[:, :, :, None, :] + cell[None, None, None]).reshape(B, -1, 3).astype(np.float32)
    pts = torch.from_numpy(pts).cuda()
    return pts, roi_pool.target_indices(pts, S.KITTI_RANGE, S.DET_VOXEL_SIZE, [8, 8, 8])


@pytest.mark.parametrize("normalize", [False, True])
def test_trilinear_rows_are_bit_exact_with_the_torch_expression(cuda, normalize):
    from btcdet_b200 import roi_pool, synthetic as S
    from oracle import roi_pool as R
    case = S.roi_head_case(batch=2, n_points=20000, n_rois=24)
    sp = _sparse(case)
    pts, zyx = _grid_targets(case)
    per_scene = pts.shape[1]
    want_c, want_f, want_t = R.interpolate_rows(sp.features, sp.indices, 2, case["x_shape"], zyx, per_scene, [2, 4, 12],
                                                normalize=normalize)
    coords, rows, tgt = roi_pool.trilinear_gather_rows(sp, zyx, per_scene, [2, 4, 12], normalize=normalize, want_target=True)
    n = want_f.shape[0]
    assert n > 3000 and rows.shape[0] == n
    assert torch.equal(rows, want_f) and torch.equal(coords.long(), want_c) and torch.equal(tgt, want_t)
    # static form: capacity-sized outputs, device count, rows beyond the capacity dropped
    for cap in (n + 100, n - 50):
        c2, r2, cnt = roi_pool.trilinear_gather_rows(sp, zyx, per_scene, [2, 4, 12], normalize=normalize, out_cap=cap)
        k = min(cap, n)
        assert int(cnt.item()) == n and torch.equal(r2[:k], want_f[:k]) and torch.equal(c2[:k].long(), want_c[:k])
        assert float(r2[k:].abs().sum()) == 0.0
    # explicit per-target scene indices
    bt = torch.arange(zyx.shape[0], device="cuda") // per_scene
    c3, r3 = roi_pool.trilinear_gather_rows(sp, zyx, 1, [2, 4, 12], normalize=normalize, b_target=bt)
    assert torch.equal(r3, want_f) and torch.equal(c3, coords)


def test_trilinear_rows_backward_is_the_autograd_gradient(cuda):
    from btcdet_b200 import roi_pool, synthetic as S
    from oracle import roi_pool as R
    case = S.roi_head_case(batch=2, n_points=8000, n_rois=8, channels=32)
    sp = _sparse(case)
    pts, zyx = _grid_targets(case)
    f_ref = sp.features.clone().requires_grad_(True)
    _, want_f, _ = R.interpolate_rows(f_ref, sp.indices, 2, case["x_shape"], zyx, pts.shape[1], [2, 4, 12])
    g = torch.randn_like(want_f)
    want_f.backward(g)
    sp.features = sp.features.clone().requires_grad_(True)
    _, rows = roi_pool.trilinear_gather_rows(sp, zyx, pts.shape[1], [2, 4, 12])
    rows.backward(g)
    assert torch.equal(rows.detach(), want_f.detach())
    assert float(f_ref.grad.abs().max()) > 0.1
    assert torch.allclose(sp.features.grad, f_ref.grad, rtol=1e-4, atol=1e-4)


def _batch_dict(case):
    import spconv
    d = {"batch_size": case["batch_size"], "rois": torch.from_numpy(case["rois"]).cuda(),
         "points": torch.from_numpy(case["points"]).cuda(), "occ_pnts": torch.from_numpy(case["occ_pnts"]).cuda(),
         "added_occ_b_ind": torch.from_numpy(case["added_occ_b_ind"]).cuda()}
    d["multi_scale_3d_features"] = {"x_combine": spconv.SparseConvTensor(
        torch.from_numpy(case["x_features"]).cuda(), torch.from_numpy(case["x_coords"]).cuda(), case["x_shape"],
        case["batch_size"])}
    return d


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not staged (oracle/stage_reference.py)")
def test_reference_conv_head_pooling_with_and_without_the_fused_ops(cuda):
    """`ConvHead.roi_conv_pool` (conv_head.py:247-379) of the reference's own class, built from the reference's yaml, on
    the GPU: (1) unchanged — its pointnet2 extension is this repo's drop-in, its spconv the shim; (2) with
    interpolate_from_3d_features and the StackSAModuleMSG forwards replaced by the fused ops.  Identical outputs."""
    from btcdet_b200 import roi_pool, synthetic as S
    mods = ref_loader.load_roi_head_modules(device="cuda")
    torch.manual_seed(0)
    head = ref_loader.build_conv_head(mods, S.DET_VOXEL_SIZE, S.KITTI_RANGE).cuda().eval()
    case = S.roi_head_case(batch=2, n_points=20000, n_rois=24, n_occ=2000)
    with torch.no_grad():
        want, _ = head.roi_conv_pool(_batch_dict(case))
        roi_pool.patch_conv_head(head)
        got, _ = head.roi_conv_pool(_batch_dict(case))
    n_grid, c_out = 27, 16 * 4 + 16 * 3 + 128                     # raw points 4 scales, occ points 3, x_combine convs
    assert tuple(want.shape) == (2 * 24, n_grid * c_out, 1)
    conv_part = want.view(2 * 24, c_out, n_grid)[:, -128:]        # channel-major: the x_combine convolutions' 128 channels
    assert float(want.abs().max()) > 0.1 and float((conv_part != 0).float().mean()) > 0.01
    assert torch.equal(got, want)
