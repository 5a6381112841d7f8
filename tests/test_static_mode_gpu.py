"""Static (capacity) mode of the spconv shim and the one-graph config-3 chain (btcdet_b200.chain.PlannedHotPath).

Static mode = SparseConvTensor.n_dev: capacity-sized tensors, every row count on the device, no host read, eval-mode
BatchNorm + ReLU folded into the convolution epilogues.  It must reproduce the exact-shape eager shim (which the other
tests pin against the oracle and the reference's own modules): coordinates / counts exactly, features to fp32 rounding of
the folded affine.  The planned chain is captured once and replayed on two different batches.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _live(t, n):
    return t[:n].detach().cpu().numpy()


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _det_inputs(seeds, cap=None):
    from btcdet_b200 import ops, synthetic as S
    scenes = [S.lidar_like(20000, seed=s) for s in seeds]
    pts, offs = S.batch_points(scenes)
    v, c, n, mean, nv = ops.voxelize(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(), S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                                     S.DET_MAX_POINTS, S.DET_MAX_VOXELS["test"], want_mean=True)
    m = int(nv[-1].item())
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = torch.cat([mean[:m], torch.rand((m, 2), device="cuda", generator=g)], dim=1).contiguous()
    occ = torch.rand((m, 2), device="cuda", generator=g)
    return feats, occ, c[:m].contiguous(), m


def test_static_det_backbone_equals_exact(cuda):
    from btcdet_b200 import backbones, ops
    torch.manual_seed(0)
    net = backbones.randomize_bn_(backbones.DetBackboneOcc(6, 4)).cuda().eval()
    feats, occ, coords, m = _det_inputs([3, 4])
    with torch.no_grad():
        ref = net({"voxel_features": feats, "occ_voxel_features": occ, "voxel_coords": coords, "batch_size": 2})
    cap = int(1.5 * m) + 321          # strided levels get the input level's capacity (ops.STATIC_GROWTH = 1.0)
    pad = lambda t: torch.cat([t, torch.zeros((cap - m,) + t.shape[1:], dtype=t.dtype, device=t.device)])   # noqa: E731
    n_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
    with torch.no_grad(), ops.static_checks() as chk:
        got = net({"voxel_features": pad(feats), "occ_voxel_features": pad(occ), "voxel_coords": pad(coords), "batch_size": 2,
                   "voxel_n_dev": n_dev})
    chk.verify()
    for key in ("encoded_spconv_tensor",):
        a, b = ref[key], got[key]
        n = a.features.shape[0]
        assert int(b.n_dev.item()) == n
        assert np.array_equal(_live(b.indices, n), a.indices.cpu().numpy())
        assert _rel(_live(b.features, n), a.features.cpu().numpy()) < 2e-5
    a, b = ref["multi_scale_3d_features"]["x_combine"], got["multi_scale_3d_features"]["x_combine"]
    n = a.features.shape[0]
    assert int(b.n_dev.item()) == n
    assert np.array_equal(_live(b.indices, n), a.indices.cpu().numpy())
    assert _rel(_live(b.features, n), a.features.cpu().numpy()) < 2e-5
    torch.testing.assert_close(got["encoded_spconv_tensor"].dense(), ref["encoded_spconv_tensor"].dense(), rtol=1e-4, atol=1e-5)


def test_static_capacity_overflow_is_reported(cuda):
    from btcdet_b200 import ops
    feats, occ, coords, m = _det_inputs([5])
    n_dev = torch.tensor([m], dtype=torch.int32, device="cuda")
    with ops.static_checks() as chk:
        rb = ops.rulebook_conv(coords, 1, [41, 1600, 1408], 3, 2, 1, out_cap=1000, n_dev=n_dev)
    assert rb.nbr_out.shape[0] == 1000
    with pytest.raises(Exception, match="capacity exceeded"):
        chk.verify()


def _chain_model(seed=0):
    from btcdet_b200 import backbones, chain
    torch.manual_seed(seed)
    m = chain.BtcHotPath()
    backbones.randomize_bn_(m, seed)
    return m.cuda().eval()


def test_planned_chain_one_graph_matches_eager_chain(cuda):
    from btcdet_b200 import chain
    model = _chain_model()
    batches = [chain.synthetic_batch([21, 22], n_points=20000, with_rot=True, mode="test"),
               chain.synthetic_batch([23, 24], n_points=17000, with_rot=True, mode="test")]
    chain.calibrate_occ_head_bias(model, batches[0], 0.03)
    # The folded BatchNorm rounds differently from torch's eval-mode kernel (~1e-7 relative), so a probability that sits on
    # the occupancy threshold could fall on either side of it: put the threshold in the middle of the widest gap between
    # the eager probabilities near 0.3 (the selection is then the same set and everything behind it compares exactly).
    near = []
    for bd in batches:
        with torch.no_grad():
            pr = model({k: (v.clone() if torch.is_tensor(v) else v) for k, v in bd.items()})["batch_pred_occ_prob"]
        near.append(pr[(pr > 0.29) & (pr < 0.31)].flatten())
    near = torch.sort(torch.cat(near + [torch.tensor([0.29, 0.31], device="cuda")])).values
    gaps = near[1:] - near[:-1]
    k = int(torch.argmax(gaps))
    assert float(gaps[k]) > 2e-6
    model.occ_thresh = float((near[k] + near[k + 1]) / 2)
    plan = chain.PlannedHotPath(model, 2, occ_vox_cap=2 * 20000, det_vox_cap=2 * 40000).capture()
    assert plan.graph is not None
    for bd in batches + batches[:1]:                     # replay on different batches, then on the first one again
        ref_in = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in bd.items()}
        with torch.no_grad():
            ref = model(ref_in)
        out = plan(bd)
        plan.verify()
        torch.cuda.synchronize()
        assert torch.equal(out["general_cls_loss_mask"], ref["general_cls_loss_mask"])
        torch.testing.assert_close(out["batch_pred_occ_prob"], ref["batch_pred_occ_prob"], rtol=1e-4, atol=1e-6)
        m = ref["voxel_coords"].shape[0]
        assert int(out["voxel_n_dev"].item()) == m
        assert np.array_equal(_live(out["voxel_coords"], m), ref["voxel_coords"].cpu().numpy())
        assert np.array_equal(_live(out["voxel_num_points"], m), ref["voxel_num_points"].cpu().numpy())
        p = ref["voxels"].shape[1]
        # raw points are copies; pseudo points carry head outputs (residuals, probabilities) through the folded BatchNorm
        torch.testing.assert_close(out["voxels"][:m, :p], ref["voxels"], rtol=1e-5, atol=1e-5)
        assert float(out["voxels"][:m, p:].abs().max()) == 0.0
        enc_r, enc = ref["encoded_spconv_tensor"], out["encoded_spconv_tensor"]
        n = enc_r.features.shape[0]
        assert int(out["encoded_n_dev"].item()) == n
        assert np.array_equal(_live(enc.indices, n), enc_r.indices.cpu().numpy())
        assert _rel(_live(enc.features, n), enc_r.features.cpu().numpy()) < 1e-4
        assert _rel(out["spatial_features"].cpu().numpy(), ref["spatial_features"].cpu().numpy()) < 1e-4


def test_fused_occ_head_probability_equals_dense_softmax(cuda):
    """btc_occ_head_prob against the reference's op sequence dense() -> softmax(dim=1)[:, -1] * mask
    (occ_head_3D.py:46-49) on random sparse logits, including cells without an active site (softmax of zeros = 0.5)."""
    import spconv
    from btcdet_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    B, nz, ny, nx = 2, 9, 157, 209
    cells = B * nz * ny * nx
    pick = torch.randperm(cells, device="cuda", generator=g)[:40000].sort().values
    b, rem = pick // (nz * ny * nx), pick % (nz * ny * nx)
    coords = torch.stack([b, rem // (ny * nx), (rem // nx) % ny, rem % nx], dim=1).int()
    logits = torch.randn((coords.shape[0], 2), device="cuda", generator=g) * 3
    mask = (torch.rand((B, nz, ny, nx), device="cuda", generator=g) > 0.4).to(torch.uint8)
    dense = spconv.SparseConvTensor(logits, coords, [nz, ny, nx], B).dense()
    want = torch.softmax(dense, dim=1)[:, -1] * mask
    got = ops.occ_head_prob(logits, coords, B, (nx, ny, nz), mask)
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-7)
    assert float((got - want).abs().max()) <= 6e-8          # one ulp of 0.5: same op order as torch's softmax
    torch.testing.assert_close(ops.occ_head_prob(logits, coords, B, (nx, ny, nz), None), torch.softmax(dense, dim=1)[:, -1],
                               rtol=1e-6, atol=1e-7)
