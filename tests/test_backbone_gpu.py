"""BASELINE.json configs[1]: VoxelBackBone8x forward on a synthetic 20k-point KITTI-range cloud,
voxel [0.05,0.05,0.1], C_in=4 — every level's indices exact, features within 1e-4 relative of the
oracle (eager spconv-shim path and the sync-free planned engine)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def _inputs(oracle, batch, n=20000, seed=100):
    from btcdet_b200 import synthetic as S
    scenes = [S.lidar_like(n, seed=seed + b) for b in range(batch)]
    v, c, npts = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    mean = (v.sum(1) / np.maximum(npts, 1)[:, None]).astype(np.float32)
    return scenes, mean, c


def _oracle_forward(model, mean, coords, batch):
    from tests import oracle_net
    x = oracle_net.to_oracle_tensor(mean, coords, model.sparse_shape, batch)
    levels = {}
    x = oracle_net.run(model.conv_input, x)
    for name in ("conv1", "conv2", "conv3", "conv4"):
        x = oracle_net.run(getattr(model, name), x)
        levels["x_" + name] = x
    levels["out"] = oracle_net.run(model.conv_out, x)
    return levels


@pytest.mark.parametrize("batch", [1, 2])
def test_voxelbackbone8x_eager_matches_oracle(cuda, oracle, batch):
    from btcdet_b200 import backbones
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    scenes, mean, coords = _inputs(oracle, batch)
    ref = _oracle_forward(model, mean, coords, batch)
    model = model.cuda()
    with torch.no_grad():
        out = model({"voxel_features": torch.from_numpy(mean).cuda(), "voxel_coords": torch.from_numpy(coords).cuda(),
                     "batch_size": batch})
    got = dict(out["multi_scale_3d_features"], out=out["encoded_spconv_tensor"])
    for name, r in ref.items():
        g = got[name]
        assert list(g.spatial_shape) == r.spatial_shape, name
        np.testing.assert_array_equal(g.indices.cpu().numpy(), r.indices, err_msg=name)
        assert rel_err(g.features.cpu().numpy(), r.features) < REL_TOL, name
    assert got["out"].spatial_shape == [2, 200, 176]


def test_training_step_runs_and_matches_dense_autograd_direction(cuda, oracle):
    """fwd + bwd through the shim in train mode (BatchNorm batch statistics): finite grads everywhere."""
    from btcdet_b200 import backbones
    torch.manual_seed(0)
    model = backbones.VoxelBackBone8x(4).cuda().train()
    _, mean, coords = _inputs(oracle, 1, n=6000)
    out = model({"voxel_features": torch.from_numpy(mean).cuda(), "voxel_coords": torch.from_numpy(coords).cuda(),
                 "batch_size": 1})
    loss = out["encoded_spconv_tensor"].dense().square().mean()
    loss.backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    assert sum(float(p.grad.abs().sum()) for p in model.parameters()) > 0


@pytest.mark.parametrize("batch,use_graph,algo,sort", [(1, True, 1, False), (3, True, 0, False), (2, False, 0, False),
                                                        (1, True, 2, False), (2, True, 0, True)])
def test_planned_engine_matches_oracle(cuda, oracle, batch, use_graph, algo, sort):
    """points -> GPU voxelize -> MeanVFE -> folded-BN backbone under one CUDA graph, no host sync."""
    from btcdet_b200 import backbones, engine, synthetic as S
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, batch, batch * 20000, S.DET_VOXEL_SIZE,
                               S.KITTI_RANGE, max_points=5, max_voxels=16000, algo=algo, use_graph=use_graph,
                               sort_rows=sort).capture()
    for rep in range(2):  # replay the same graph on different inputs
        scenes, mean, coords = _inputs(oracle, batch, seed=200 + 10 * rep)
        ref = _oracle_forward(model, mean, coords, batch)["out"]
        pts, offs = S.batch_points(scenes)
        feat, out_coords, n_dev = plan.forward(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda())
        counts = plan.read_counts()
        n = int(n_dev.item())
        assert counts[0] == coords.shape[0] and counts[-1] == n == ref.indices.shape[0]
        np.testing.assert_array_equal(out_coords[:n].cpu().numpy(), ref.indices)
        assert rel_err(feat[:n].cpu().numpy(), ref.features) < REL_TOL


def test_planned_engine_reports_capacity_overflow(cuda):
    from btcdet_b200 import _lib, backbones, engine, synthetic as S
    model = backbones.VoxelBackBone8x(4).eval()
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, 1, 20000, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_voxels=16000, level_growth=0.05, algo=1, use_graph=False).capture()
    pts, offs = S.batch_points([S.uniform(20000, seed=1)])   # uniform clouds dilate under strided convs
    plan.forward(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda())
    with pytest.raises(_lib.BtcError):
        plan.read_counts()


def _randomize(model, seed):
    from btcdet_b200 import backbones
    torch.manual_seed(seed)
    for m in model.modules():
        if hasattr(m, "reset_parameters") and not isinstance(m, torch.nn.BatchNorm1d):
            m.reset_parameters()
    return backbones.randomize_bn_(model, seed).eval()


def test_reference_det_backbone_topology_matches_oracle(cuda, oracle):
    """VoxelBackBone8xOcc dataflow (spconv_backbone.py:936-1019): SubM/SparseConv/SparseMaxPool3d, sparse_cat across
    tensors built by different ops, cached 'spconv3'/'spconv4' rulebooks reused by down2/down3, dense() + gather."""
    import spconv
    from btcdet_b200 import synthetic as S
    from tests import models_mirror, oracle_net
    batch = 2
    model = _randomize(models_mirror.DetBackboneMirror(6, 4), 3)
    scenes = [S.lidar_like(12000, seed=300 + b) for b in range(batch)]
    v, coords, npts = oracle.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    rng = np.random.default_rng(0)
    feats = np.concatenate([(v.sum(1) / np.maximum(npts, 1)[:, None]), rng.uniform(0, 1, (v.shape[0], 2))], 1).astype(np.float32)
    occ_feats = np.abs(rng.standard_normal((v.shape[0], 2))).astype(np.float32)
    ref = model.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, model.sparse_shape, batch),
                    occ_feats)
    model = model.cuda()
    with torch.no_grad():
        x = spconv.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda(), model.sparse_shape, batch)
        got = model.run(models_mirror.ShimBackend(), x, torch.from_numpy(occ_feats).cuda())
    for name in ("x_conv2", "x_conv3", "out", "x_combine"):
        np.testing.assert_array_equal(got[name].indices.cpu().numpy(), ref[name].indices, err_msg=name)
        assert rel_err(got[name].features.cpu().numpy(), ref[name].features) < REL_TOL, name
    assert got["x_combine"].features.shape[1] == 128 and list(got["out"].spatial_shape) == [2, 200, 176]


def test_reference_occ_backbone_topology_matches_oracle(cuda, oracle):
    """VoxelBackBoneDeconv + OccHead3D convs (spconv_backbone.py:138-203, occ_head_3D.py:41-52): dilating
    SparseConv3d, two SparseConvTranspose3d, SubM heads with/without bias, dense() of the head outputs."""
    import spconv
    from btcdet_b200 import synthetic as S
    from tests import models_mirror, oracle_net
    from oracle.occ_masks import OccGeometry
    geo = OccGeometry()
    batch = 2
    model = _randomize(models_mirror.OccBackboneMirror(4), 5)
    gen = oracle.VoxelGeneratorV2(geo.voxel_size, geo.point_cloud_range, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"])
    fs, cs = [], []
    for b in range(batch):
        pts = S.lidar_like(6000, seed=400 + b)
        cyl = np.stack([np.linalg.norm(pts[:, :2], axis=1), np.arctan2(-pts[:, 1], pts[:, 0]) * 180. / np.pi, pts[:, 2],
                        pts[:, 3]], -1).astype(np.float32)
        r = gen.generate(cyl)
        fs.append(r["voxels"].sum(1) / np.maximum(r["num_points_per_voxel"], 1)[:, None])
        cs.append(np.pad(r["coordinates"], ((0, 0), (1, 0)), constant_values=b))
    feats, coords = np.concatenate(fs).astype(np.float32), np.concatenate(cs).astype(np.int32)
    feats = feats / np.array([70.0, 40.0, 3.0, 1.0], np.float32)      # keep activations O(1)
    ref = model.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, model.sparse_shape, batch))
    model = model.cuda()
    with torch.no_grad():
        x = spconv.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda(), model.sparse_shape, batch)
        got = model.run(models_mirror.ShimBackend(), x)
        dense_cls = got["cls"].dense()
    for name in ("encoded", "cls", "res"):
        np.testing.assert_array_equal(got[name].indices.cpu().numpy(), ref[name].indices, err_msg=name)
        assert rel_err(got[name].features.cpu().numpy(), ref[name].features) < REL_TOL, name
    assert list(got["encoded"].spatial_shape) == [9, 157, 209] and got["encoded"].indices.shape[0] > 5 * coords.shape[0]
    np.testing.assert_array_equal(dense_cls.cpu().numpy(), oracle.dense(got["cls"].features.cpu().numpy(),
                                                                         ref["cls"].indices, [9, 157, 209], batch))


def test_pipelined_submit_retrieve_matches_forward(cuda, oracle):
    """engine.submit/retrieve (2-deep pipeline, results staged compactly and copied out on a side stream) returns
    exactly what the synchronous forward computes, in submission order."""
    from btcdet_b200 import backbones, engine, synthetic as S
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, 2, 2 * 20000, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=5, max_voxels=16000).capture().enable_pipeline()
    batches = []
    for i in range(4):
        pts, offs = S.batch_points([S.lidar_like(20000, seed=500 + 2 * i), S.lidar_like(15000, seed=501 + 2 * i)])
        batches.append((torch.from_numpy(pts).pin_memory(), torch.from_numpy(offs).pin_memory()))
    want = []
    for p, o in batches:
        feat, coords, n_dev = plan.forward(p, o)
        n = int(n_dev.item())
        want.append((feat[:n].cpu().clone(), coords[:n].cpu().clone()))
    got = []
    for i, (p, o) in enumerate(batches):
        plan.submit(p, o)
        if i >= 1:
            f, c, ev = plan.retrieve()
            ev.synchronize()
            got.append((f.clone(), c.clone()))
    f, c, ev = plan.retrieve()
    ev.synchronize()
    got.append((f.clone(), c.clone()))
    assert len(got) == len(want)
    for (gf, gc), (wf, wc) in zip(got, want):
        assert torch.equal(gc, wc) and torch.equal(gf, wf)
    # the same pipeline with the result rows left on the device (retrieve(to_host=False)): views of the slot's device
    # staging buffers, identical rows, only the counts cross to the host
    got_dev = []
    for i, (p, o) in enumerate(batches):
        plan.submit(p, o)
        if i >= 1:
            f, c, ev = plan.retrieve(to_host=False)
            assert f.is_cuda and c.is_cuda
            got_dev.append((f.cpu(), c.cpu()))
    f, c, ev = plan.retrieve(to_host=False)
    got_dev.append((f.cpu(), c.cpu()))
    for (gf, gc), (wf, wc) in zip(got_dev, want):
        assert torch.equal(gc, wc) and torch.equal(gf, wf)
