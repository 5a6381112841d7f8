"""BASELINE config 3: the composed hot path (btcdet_b200.chain.BtcHotPath = modules 1-8 of BtcNet.forward,
btcdet/models/detectors/btcnet.py:32-56) on B = 2 synthetic scenes.

(a) stage-wise against the oracle: every stage's oracle (torch restatement on the same device for the mask / injection
    stages — bit-exact rules as in their own tests; the C oracle on the CPU for the sparse layers) is fed THE CHAIN'S OWN
    input of that stage, so a mismatch is attributable to one stage; indices exact, features 1e-4, masks exact.
(b) against the REFERENCE'S OWN MODULES chained the way BtcNet chains them (OccTargets3D -> MeanVFE ->
    VoxelBackBoneDeconv -> OccHead convs -> PassOccVox -> OccVFE -> VoxelBackBone8xOcc -> HeightCompression), every torch
    line of theirs executed on the GPU and every spconv layer on this repo's shim; needs the staged reference sources.
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

REL = 1e-4


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def _model(seed=0):
    from btcdet_b200 import backbones, chain
    torch.manual_seed(seed)
    m = chain.BtcHotPath()
    for mod in m.modules():
        if hasattr(mod, "reset_parameters") and not isinstance(mod, torch.nn.BatchNorm1d):
            mod.reset_parameters()
    backbones.randomize_bn_(m, seed)
    return m.cuda().eval()


def test_chain_stagewise_against_oracle(cuda, oracle):
    from btcdet_b200 import chain, ops, synthetic as S
    from oracle import box_masks, occ_inject, occ_masks
    from tests import models_mirror, oracle_net
    geo = occ_masks.OccGeometry()
    model = _model()
    bd = chain.synthetic_batch([11, 12], n_points=20000, with_rot=True, mode="test")
    chain.calibrate_occ_head_bias(model, bd, 0.03)      # an untrained head would pass nothing (or everything)
    inp = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in bd.items()}
    with torch.no_grad():
        out = model(bd)
    B = 2
    # ---- stage 1: masks (same-device torch restatement; bit-exact rules of test_occ_gpu / test_box_masks_gpu)
    occ = occ_masks.occ_targets(inp["voxels"], inp["voxel_coords"], inp["voxel_num_points"], B, geo, rot_z=inp["rot_z"])
    box = box_masks.box_targets(occ["valid_coords"], occ["valid_feats"], inp["gt_boxes"], inp["gt_boxes_num"], inp["box_mirr_flag"],
                                B, geo, rot_z=inp["rot_z"])
    maps = box_masks.loss_maps(occ, box)
    for k in ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask"):
        assert torch.equal(out[k].bool(), occ[k].bool()), k
    for k in ("occ_fore_cls_mask", "occ_mirr_cls_mask", "pos_mask", "general_reg_loss_mask"):
        assert torch.equal(out[k].bool(), maps[k].bool()), k
    assert torch.equal(out["forebox_label"], box["forebox_label"])
    assert torch.equal(out["general_cls_loss_mask_float"], maps["general_cls_loss_mask_float"])
    # ---- stage 2: absolute coordinates + MeanVFE
    vabs = torch.cat([occ_masks.cylinder_uvd2absxyz(inp["voxels"][..., 0], inp["voxels"][..., 1], inp["voxels"][..., 2]),
                      inp["voxels"][..., 3:]], dim=-1)
    mean = vabs.sum(1) / torch.clamp_min(inp["voxel_num_points"].view(-1, 1).float(), 1.0)
    # ---- stages 3-4: occupancy backbone + head on the CPU oracle, from the chain's own VFE features
    mir = models_mirror.OccBackboneMirror(4).eval()
    sd = dict(model.occ_backbone.state_dict())
    sd.update(model.occ_head.state_dict())
    mir.load_state_dict({k: v.cpu() for k, v in sd.items()})
    feats0 = (vabs.sum(1) / torch.clamp_min(inp["voxel_num_points"].view(-1, 1).float(), 1.0))
    torch.testing.assert_close(feats0, mean, rtol=0, atol=0)
    got_vfe = chain.occ_abs_mean_vfe(inp["voxels"], inp["voxel_num_points"])[1]
    torch.testing.assert_close(got_vfe, mean, rtol=1e-6, atol=1e-5)
    ref3 = mir.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(got_vfe.cpu().numpy(), inp["voxel_coords"].cpu().numpy(),
                                                                              mir.sparse_shape, B))
    enc = out["encoded_spconv_tensor_occ"] if "encoded_spconv_tensor_occ" in out else None
    logits = oracle.dense(ref3["cls"].features, ref3["cls"].indices, [9, 157, 209], B)
    assert rel_err(out["pred_occ_logit"].cpu().numpy(), logits) < REL
    res_dense = oracle.dense(ref3["res"].features, ref3["res"].indices, [9, 157, 209], B)
    assert rel_err(out["pred_sem_residuals"].cpu().numpy(), res_dense) < REL
    prob = torch.softmax(out["pred_occ_logit"], dim=1)[:, -1] * out["general_cls_loss_mask"]
    assert torch.equal(prob, out["batch_pred_occ_prob"])
    n_above = int((prob > 0.3).sum())
    assert 200 < n_above < 80000, n_above
    # ---- stages 5-6: injection + OccVFE (same-device restatement, from the chain's own probabilities)
    inj = occ_inject.pass_occ_vox(out["batch_pred_occ_prob"], out["pred_sem_residuals"], inp["det_voxels"],
                                  inp["det_voxel_num_points"], inp["det_voxel_coords"], geo, S.DET_VOXEL_SIZE, [1408, 1600, 40],
                                  S.KITTI_RANGE, thresh=0.3, max_points=40000, rot_z=inp["rot_z"])
    assert torch.equal(out["voxel_coords"].long(), inj["voxel_coords"].long())
    assert torch.equal(out["voxel_num_points"].long(), inj["voxel_num_points"].long())
    assert torch.equal(out["added_occ_xyz"], inj["occ_xyz"])
    rf, ro = occ_inject.occ_vfe(out["voxels"], out["voxel_num_points"], 4)
    torch.testing.assert_close(out["voxel_features"], rf, rtol=1e-6, atol=1e-6)
    assert torch.equal(out["occ_voxel_features"], ro)
    # ---- stage 7-8: detection backbone on the CPU oracle, from the chain's own voxel features
    dm = models_mirror.DetBackboneMirror(6, 4).eval()
    dm.load_state_dict({k: v.cpu() for k, v in model.det_backbone.state_dict().items()})
    ref7 = dm.run(models_mirror.OracleBackend(),
                  oracle_net.to_oracle_tensor(out["voxel_features"].cpu().numpy(), out["voxel_coords"].cpu().numpy(), dm.sparse_shape, B),
                  out["occ_voxel_features"].cpu().numpy())
    enc = out["encoded_spconv_tensor"]
    np.testing.assert_array_equal(enc.indices.cpu().numpy(), ref7["out"].indices)
    assert rel_err(enc.features.cpu().numpy(), ref7["out"].features) < REL
    xc = out["multi_scale_3d_features"]["x_combine"]
    np.testing.assert_array_equal(xc.indices.cpu().numpy(), ref7["x_combine"].indices)
    assert rel_err(xc.features.cpu().numpy(), ref7["x_combine"].features) < REL
    sf = oracle.dense(ref7["out"].features, ref7["out"].indices, [2, 200, 176], B).reshape(B, 256, 200, 176)
    assert out["spatial_features"].shape == (B, 256, 200, 176)
    assert rel_err(out["spatial_features"].cpu().numpy(), sf) < REL


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not staged (oracle/stage_reference.py)")
def test_chain_against_the_reference_modules_on_cuda(cuda, oracle):
    import make_occ_golden
    import test_occ_inject_cpu as TI
    from btcdet_b200 import chain, synthetic as S
    from oracle import occ_masks
    geo = occ_masks.OccGeometry()
    mods = ref_loader.load_reference_modules("cuda")
    Cfg = ref_loader.Cfg
    model = _model(1)
    bd = chain.synthetic_batch([21, 22], n_points=20000, with_rot=True, mode="test")
    chain.calibrate_occ_head_bias(model, bd, 0.03)
    inp = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in bd.items()}
    with torch.no_grad():
        out = model(bd)
    # ---- the reference's modules, configured as in btcdet_kitti_car.yaml, sharing the same weights
    sb = mods["spconv_backbone"]
    occ_bb = sb.VoxelBackBoneDeconv(Cfg(), input_channels=4, grid_size=[209, 157, 9]).cuda().eval()
    occ_bb.load_state_dict(model.occ_backbone.state_dict())
    cfg = Cfg(OCC_CONV_TYPE=['identity', 'maxpool'], OCC_CONV_EXECUTE=[False, True],
              OUT_FEAT_TYPE=['None', 'None', 'None', 'None', 'big_bev_combine'])
    det_bb = sb.VoxelBackBone8xOcc(cfg, input_channels=6, grid_size=np.array([1408, 1600, 40]),
                                   original_num_rawpoint_features=4).cuda().eval()
    det_bb.load_state_dict(model.det_backbone.state_dict())
    nx, ny, nz = geo.grid_size
    vs = torch.tensor(geo.voxel_size, dtype=torch.float32, device="cuda")
    centers = mods["coords_utils"].get_all_voxel_centers_zyx(1, torch.tensor([nx, ny, nz], dtype=torch.int32, device="cuda"),
                                                             geo.point_cloud_range[:3], vs)[0]
    centers = mods["coords_utils"].uvd2absxyz(centers[2], centers[1], centers[0], "cylinder", dim=-1)
    vc = {"all_voxel_centers": centers, "all_voxel_centers_2d": torch.mean(centers[:, :, :, :2], dim=0).view(-1, 2)}
    data_cfg = Cfg.wrap(make_occ_golden.data_cfg(geo))
    tgt = mods["occ_targets_3d"].OccTargets3D(Cfg.wrap(make_occ_golden.MODEL_OCC_CFG), voxel_size=geo.voxel_size,
                                              point_cloud_range=geo.point_cloud_range, data_cfg=data_cfg,
                                              grid_size=geo.grid_size, num_class=1, voxel_centers=vc).cuda()
    pov = mods["pass_occ_vox"].PassOccVox(Cfg.wrap(TI.MODEL_CFG), data_cfg, S.KITTI_RANGE, geo.voxel_size, geo.grid_size,
                                          S.DET_VOXEL_SIZE, [1408, 1600, 40], "test", vc)
    vfe = mods["occ_vfe"].OccVFE(Cfg(), 6, Cfg.wrap({"POINT_FEATURE_ENCODING": {"used_feature_list": ["x", "y", "z", "intensity"]}}),
                                 maxprob=True)
    with torch.no_grad():
        r = dict(inp)
        r["voxel_coords"] = r["voxel_coords"].float()              # load_data_to_gpu hands every array over as float32
        r["voxel_num_points"] = r["voxel_num_points"].float()
        r["det_voxel_coords"] = r["det_voxel_coords"].float()
        r["det_voxel_num_points"] = r["det_voxel_num_points"].float()
        r = tgt(r)                                                   # 1 OccTargets3D.forward
        v = r["voxels"]                                              # 2 MeanVFE.forward (mean_vfe.py:27-44, maxprob False)
        r["voxel_features"] = (v.sum(dim=1) / torch.clamp_min(r["voxel_num_points"].view(-1, 1), min=1.0).type_as(v)).contiguous()
        r = occ_bb(r)                                                # 3 VoxelBackBoneDeconv.forward
        enc = r["encoded_spconv_tensor"]                             # 4 OccHead3D.forward (:41-52) on the shared head convs
        logits = model.occ_head.conv_cls(enc).dense()
        r["batch_pred_occ_prob"] = torch.softmax(logits, dim=1)[:, -1:, ...][:, -1, ...] * r["general_cls_loss_mask"]
        r["pred_sem_residuals"] = model.occ_head.conv_res(enc).dense()
        r["use_occ_prob"] = [True] * 2
        r = pov(r)                                                   # 5 PassOccVox.forward
        r = vfe(r)                                                   # 6 OccVFE.forward
        r = det_bb(r)                                                # 7 VoxelBackBone8xOcc.forward
        d = r["encoded_spconv_tensor"].dense()                       # 8 HeightCompression.forward
        n, c, dd, h, w = d.shape
        ref_sf = d.view(n, c * dd, h, w)
    for k in ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask", "pos_mask", "forebox_label"):
        assert torch.equal(out[k].to(r[k].dtype), r[k]), k
    assert rel_err(out["batch_pred_occ_prob"].cpu().numpy(), r["batch_pred_occ_prob"].cpu().numpy()) < REL
    # the two chains threshold probabilities that agree to ~1e-6: a cell within that distance of 0.3 may be kept by one only
    same = out["voxel_coords"].shape == r["voxel_coords"].shape and torch.equal(out["voxel_coords"].long(), r["voxel_coords"].long())
    near = int(((r["batch_pred_occ_prob"] - 0.3).abs() < 1e-5).sum())
    assert same or near > 0, "re-voxelised coordinate sets differ without any probability at the threshold"
    if same:
        assert torch.equal(out["voxel_num_points"].long(), r["voxel_num_points"].long())
        assert rel_err(out["voxel_features"].cpu().numpy(), r["voxel_features"].cpu().numpy()) < REL
        a, b = out["encoded_spconv_tensor"], r["encoded_spconv_tensor"]
        assert torch.equal(a.indices, b.indices)
        assert rel_err(a.features.cpu().numpy(), b.features.cpu().numpy()) < REL
        xa, xb = out["multi_scale_3d_features"]["x_combine"], r["multi_scale_3d_features"]["x_combine"]
        assert torch.equal(xa.indices, xb.indices) and rel_err(xa.features.cpu().numpy(), xb.features.cpu().numpy()) < REL
        assert rel_err(out["spatial_features"].cpu().numpy(), ref_sf.cpu().numpy()) < REL
