"""O3: execute the REFERENCE'S OWN torch code on the B200 and compare this library against it.

Needs the staged reference sources (`python oracle/stage_reference.py` in the build container; they travel to the GPU box
in the git-ignored oracle/_ref/).  Sections:
  masks     OccTargets3D.forward (occ_targets_3d.py:18-93) on CUDA  vs  ops.occ_training_targets  (rows a5-a12)
  inverse   the 4x4 box transforms and their torch.inverse on CUDA (what a9-a11 must reproduce bit for bit)
  inject    PassOccVox.forward + OccVFE.forward on CUDA             vs  ops.pass_occ_vox / ops.occ_vfe (rows a17-a20)
  backbone  VoxelBackBoneDeconv.forward / VoxelBackBone8xOcc.forward (spconv_backbone.py:138-203, 936-1019) THROUGH THE
            SHIM on CUDA  vs  the test mirror on the shim (bit-equal) and the CPU oracle (1e-4)
Writes gpurun_out/o3/report.json and CUDA-generated fixtures gpurun_out/o3/*_cuda.npz (copied to tests/golden/ by hand).
"""
import json
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out", "o3")

MASK_CASES = [("a", [3, 4], 6000, True, True), ("b", [11], 6000, False, False), ("c", [21, 22, 23], 20000, True, True)]
U8 = ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask", "fore_voxelwise_mask", "pos_mask",
      "occ_fore_cls_mask", "occ_mirr_cls_mask", "occ_bm_cls_mask", "general_reg_loss_mask", "forebox_label")
F32 = ("general_cls_loss_mask_float", "general_reg_loss_mask_float", "res_mtrx")


def _mask_case(seeds, n, with_rot, with_bm):
    import make_occ_golden as G
    import test_box_masks_cpu as TB
    inp, geo = G.make_inputs(seeds, n_points=n, with_rot=with_rot)
    if with_rot:
        inp["rot_z"] = np.array([7.5, -11.25, 3.0, -2.0][:len(seeds)], np.float32)
    if with_bm:
        inp["bm_points"] = TB.bm_points_for(inp, 5)
    return inp, geo


def section_masks(report):
    import make_occ_golden as G
    from btcdet_b200 import ops
    for tag, seeds, n, with_rot, with_bm in MASK_CASES:
        inp, geo = _mask_case(seeds, n, with_rot, with_bm)
        ref = G.run_reference({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in inp.items()}, geo,
                              device="cuda", keep_all=True)
        gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                         geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
        t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
        got = ops.occ_training_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], len(seeds), t["gt_boxes"],
                                       inp["gt_boxes_num"], gf, gi, box_mirr_flag=t["box_mirr_flag"],
                                       bm_points=t.get("bm_points"), rot_z=t.get("rot_z"))
        rep = {}
        for k in U8:
            if k in ref and got.get(k) is not None:
                r, g = ref[k], got[k].cpu().numpy()
                rep[k] = {"ref_set": int((r != 0).sum()), "got_set": int((g != 0).sum()),
                          "diff": int((r.astype(np.int32) != g.astype(np.int32)).sum())}
        for k in F32:
            if k in ref and got.get(k) is not None:
                r, g = ref[k], got[k].cpu().numpy()
                d = np.abs(r - g)
                rep[k] = {"ref_nonzero": int((r != 0).sum()), "cells_differ": int((r != g).sum()),
                          "max_abs": float(d.max()), "cells_over_1e-5": int((d > 1e-5).sum())}
        report["masks_" + tag] = rep
        packed = {}
        for k in U8:
            if k in ref:
                packed["ref_" + k] = np.packbits(ref[k] != 0) if k != "forebox_label" else ref[k].astype(np.int8)
        for k in F32:
            if k in ref:
                nz = np.flatnonzero(ref[k])
                packed["ref_" + k + "_idx"] = nz.astype(np.int32)
                packed["ref_" + k + "_val"] = ref[k].reshape(-1)[nz].astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "occ_masks_cuda_%s.npz" % tag), shape=np.array(ref["voxelwise_mask"].shape),
                            seeds=np.array(seeds), n_points=n, with_rot=with_rot, with_bm=with_bm, **packed)


def section_inverse(report):
    """Dump T and torch.inverse(T) (CUDA) for a batch of box transforms, plus box-frame coordinates of probe points, so
    the LU op order can be reproduced offline."""
    from oracle import box_masks
    rng = np.random.default_rng(7)
    m = 256
    boxes = np.zeros((m, 8), np.float32)
    boxes[:, 0] = rng.uniform(0, 70, m)
    boxes[:, 1] = rng.uniform(-40, 40, m)
    boxes[:, 2] = rng.uniform(-2.5, 0.5, m)
    boxes[:, 3:6] = rng.uniform(0.5, 5, (m, 3))
    boxes[:, 6] = rng.uniform(-np.pi, np.pi, m)
    boxes[:, 7] = 1
    b = torch.from_numpy(boxes).cuda()
    rot = box_masks.yaw_rotation(b[:, 6])
    T = box_masks.transform4(rot, b[:, :3])
    inv_all = torch.inverse(T)
    inv_12 = torch.inverse(T[:12])
    inv_1 = torch.stack([torch.inverse(T[i]) for i in range(16)])
    pts = torch.from_numpy(rng.uniform(-1, 1, (512, 3)).astype(np.float32) * 40 + np.array([35, 0, -1], np.float32)).cuda()
    q = torch.einsum("nj,mij->nmi", pts, inv_12[:, :3, :3]) + inv_12[:, :3, 3]
    # 2-D transforms of torch_points_in_box_2d_mask (point_box_utils.py:332-365), the mirrored points' way back
    # (einsum "nmj,mij->nmi", :295-301) and rotatez (torch.matmul, :241-250): the remaining GEMM shapes of rows a9-a11
    c2, s2 = torch.cos(b[:, 6]), torch.sin(b[:, 6])
    rot2 = torch.stack([torch.stack([c2, -1.0 * s2], dim=-1), torch.stack([s2, c2], dim=-1)], dim=-2)
    t2 = torch.cat([rot2, b[:, :2].unsqueeze(-1)], dim=-1)
    last = torch.cat([torch.zeros_like(b[:, :2]), torch.ones_like(b[:, 0:1])], dim=-1)
    T2 = torch.cat([t2, last.unsqueeze(-2)], dim=-2)
    inv2_all = torch.inverse(T2)
    q2 = torch.einsum("nj,mij->nmi", pts[:, :2].contiguous(), inv2_all[:12, :2, :2]) + inv2_all[:12, :2, 2]
    qm = q.clone()
    qm[:, :, 1] = -qm[:, :, 1]
    back = torch.einsum("nmj,mij->nmi", qm, rot[:12]) + b[:12, :3]
    yaw = torch.tensor(7.5, device="cuda") * np.pi / 180.
    cy_, sy_ = torch.cos(yaw), torch.sin(yaw)
    r3 = torch.transpose(torch.stack([torch.stack([cy_, -1.0 * sy_, torch.zeros_like(yaw)], dim=-1),
                                      torch.stack([sy_, cy_, torch.zeros_like(yaw)], dim=-1),
                                      torch.stack([torch.zeros_like(yaw), torch.zeros_like(yaw), torch.ones_like(yaw)], dim=-1)],
                                     dim=-2), 0, 1)
    r2 = torch.transpose(torch.stack([torch.stack([cy_, -1.0 * sy_], dim=-1), torch.stack([sy_, cy_], dim=-1)], dim=-2), 0, 1)
    rot3_pts = torch.matmul(pts, r3)
    rot2_pts = torch.matmul(pts[:, :2].contiguous(), r2)
    np.savez_compressed(os.path.join(OUT, "box_inverse_cuda.npz"), boxes=boxes, T=T.cpu().numpy(), inv_all=inv_all.cpu().numpy(),
                        inv_12=inv_12.cpu().numpy(), inv_1=inv_1.cpu().numpy(), pts=pts.cpu().numpy(), q=q.cpu().numpy(),
                        cos=torch.cos(b[:, 6]).cpu().numpy(), sin=torch.sin(b[:, 6]).cpu().numpy(),
                        T2=T2.cpu().numpy(), inv2_all=inv2_all.cpu().numpy(), q2=q2.cpu().numpy(), back=back.cpu().numpy(),
                        rot_cos=cy_.cpu().numpy(), rot_sin=sy_.cpu().numpy(), rot3_pts=rot3_pts.cpu().numpy(),
                        rot2_pts=rot2_pts.cpu().numpy())
    report["inverse"] = {"batched_eq_first12": bool(torch.equal(inv_all[:12], inv_12)),
                         "batched_eq_single": bool(torch.equal(inv_all[:16], inv_1)),
                         "cpu_eq_cuda": bool(torch.equal(torch.inverse(T.cpu()), inv_all.cpu()))}


def section_inject(report):
    import ref_loader
    import make_occ_golden
    import test_occ_inject_cpu as TI
    from btcdet_b200 import ops, synthetic as S
    mods = ref_loader.load_reference_modules("cuda")
    for seed, with_rot, is_train, dense in [(1, True, False, 0.004), (2, False, False, 0.004), (3, True, True, 0.05)]:
        case, geo = TI.make_case(seed, dense=dense, with_rot=with_rot)
        case = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}
        data_cfg = ref_loader.Cfg.wrap(make_occ_golden.data_cfg(geo))
        mod = mods["pass_occ_vox"].PassOccVox(ref_loader.Cfg.wrap(TI.MODEL_CFG), data_cfg, S.KITTI_RANGE, geo.voxel_size,
                                              geo.grid_size, S.DET_VOXEL_SIZE, [1408, 1600, 40], "train",
                                              {"all_voxel_centers": torch.zeros(1, device="cuda")})
        bd = {"voxels": torch.zeros(3, 12, 4, device="cuda"), "voxel_num_points": torch.ones(3, device="cuda"),
              "voxel_coords": torch.zeros(3, 4, device="cuda"), "batch_size": case["batch"],
              "batch_pred_occ_prob": case["probs"].clone(), "pred_sem_residuals": case["res"].clone(),
              "points": case["points"], "det_voxels": case["det_voxels"].clone(), "det_voxel_coords": case["det_voxel_coords"],
              "det_voxel_num_points": case["det_voxel_num_points"], "is_train": is_train, "use_occ_prob": [True] * case["batch"]}
        if case["rot_z"] is not None:
            bd["rot_z"] = case["rot_z"]
        ref = mod(bd)
        vfe = mods["occ_vfe"].OccVFE(ref_loader.Cfg(), 6, ref_loader.Cfg.wrap(
            {"POINT_FEATURE_ENCODING": {"used_feature_list": ["x", "y", "z", "intensity"]}}), maxprob=True)
        ref = vfe(ref)
        vox, cnt, vc, sel = ops.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                             case["det_voxel_coords"], case["batch"], 0.3, 2048 if is_train else 40000,
                                             geo.voxel_size, geo.point_cloud_range[:3], S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                                             [1408, 1600, 40], rot_z=case["rot_z"])
        feats, occ = ops.occ_vfe(vox, cnt, 4)
        rep = {"coords_equal": bool(torch.equal(vc.long(), ref["voxel_coords"].long())),
               "counts_equal": bool(torch.equal(cnt.long(), ref["voxel_num_points"].long())),
               "n_voxels": int(vc.shape[0])}
        if not is_train:
            rep["occ_xyz_equal"] = bool(torch.equal(sel["occ_xyz"], ref["added_occ_xyz"]))
            rep["occ_xyz_maxdiff"] = float((sel["occ_xyz"] - ref["added_occ_xyz"]).abs().max()) if sel["occ_xyz"].shape == ref["added_occ_xyz"].shape else -1.0
            rep["occ_b_equal"] = bool(torch.equal(sel["occ_coords"][:, 0].long(), ref["added_occ_b_ind"].long()))
        if rep["coords_equal"]:
            rep["voxel_features_maxdiff"] = float((feats - ref["voxel_features"]).abs().max())
            rep["voxel_features_equal"] = bool(torch.equal(feats, ref["voxel_features"]))
            rep["occ_voxel_features_equal"] = bool(torch.equal(occ, ref["occ_voxel_features"]))
        report["inject_seed%d" % seed] = rep
        if not is_train:
            np.savez_compressed(os.path.join(OUT, "occ_inject_cuda_%d.npz" % seed), seed=seed, with_rot=with_rot, dense=dense,
                                voxel_coords=ref["voxel_coords"].cpu().numpy().astype(np.int32),
                                voxel_num_points=ref["voxel_num_points"].cpu().numpy().astype(np.int32),
                                added_occ_xyz=ref["added_occ_xyz"].cpu().numpy(),
                                voxel_features=ref["voxel_features"].cpu().numpy(),
                                occ_voxel_features=ref["occ_voxel_features"].cpu().numpy())


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def section_backbone(report):
    import ref_loader
    import spconv
    from btcdet_b200 import backbones, synthetic as S
    from oracle import oracle as O
    from oracle.occ_masks import OccGeometry
    from tests import models_mirror, oracle_net
    O.build()
    mods = ref_loader.load_reference_modules("cuda")
    sb = mods["spconv_backbone"]
    Cfg = ref_loader.Cfg

    def randomize(model, seed):
        torch.manual_seed(seed)
        for m in model.modules():
            if hasattr(m, "reset_parameters") and not isinstance(m, torch.nn.BatchNorm1d):
                m.reset_parameters()
        return backbones.randomize_bn_(model, seed).eval()

    # ---- det backbone: the reference class's own forward on the shim
    batch = 2
    cfg = Cfg(OCC_CONV_TYPE=['identity', 'maxpool'], OCC_CONV_EXECUTE=[False, True],
              OUT_FEAT_TYPE=['None', 'None', 'None', 'None', 'big_bev_combine'])
    ref_det = randomize(sb.VoxelBackBone8xOcc(cfg, input_channels=6, grid_size=np.array([1408, 1600, 40]),
                                              original_num_rawpoint_features=4), 3)
    mir = models_mirror.DetBackboneMirror(6, 4).eval()
    mir.load_state_dict(ref_det.state_dict())
    scenes = [S.lidar_like(12000, seed=300 + b) for b in range(batch)]
    v, coords, npts = O.voxelize_batch(scenes, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000)
    rng = np.random.default_rng(0)
    feats = np.concatenate([(v.sum(1) / np.maximum(npts, 1)[:, None]), rng.uniform(0, 1, (v.shape[0], 2))], 1).astype(np.float32)
    occ_feats = np.abs(rng.standard_normal((v.shape[0], 2))).astype(np.float32)
    want = mir.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, mir.sparse_shape, batch), occ_feats)
    ref_det, mir = ref_det.cuda(), mir.cuda()
    with torch.no_grad():
        bd = ref_det({"voxel_features": torch.from_numpy(feats).cuda(), "voxel_coords": torch.from_numpy(coords).cuda().float(),
                      "occ_voxel_features": torch.from_numpy(occ_feats).cuda(), "batch_size": batch})
        x = spconv.SparseConvTensor(torch.from_numpy(feats).cuda(), torch.from_numpy(coords).cuda(), mir.sparse_shape, batch)
        m_out = mir.run(models_mirror.ShimBackend(), x, torch.from_numpy(occ_feats).cuda())
    r_out = {"out": bd["encoded_spconv_tensor"], "x_combine": bd["multi_scale_3d_features"]["x_combine"]}
    rep = {}
    for k in ("out", "x_combine"):
        rep[k] = {"rows": int(r_out[k].features.shape[0]),
                  "indices_eq_mirror": bool(torch.equal(r_out[k].indices, m_out[k].indices)),
                  "features_eq_mirror": bool(torch.equal(r_out[k].features, m_out[k].features)),
                  "indices_eq_oracle": bool(np.array_equal(r_out[k].indices.cpu().numpy(), want[k].indices)),
                  "rel_err_vs_oracle": _rel(r_out[k].features.cpu().numpy(), want[k].features)}
    report["backbone_det_reference_forward"] = rep
    # ---- occ backbone
    geo = OccGeometry()
    ref_occ = randomize(sb.VoxelBackBoneDeconv(Cfg(), input_channels=4, grid_size=[209, 157, 9]), 5)
    mir = models_mirror.OccBackboneMirror(4).eval()
    mir.load_state_dict(ref_occ.state_dict(), strict=False)
    gen = O.VoxelGeneratorV2(geo.voxel_size, geo.point_cloud_range, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["train"])
    fs, cs = [], []
    for b in range(batch):
        pts = S.lidar_like(6000, seed=400 + b)
        cyl = np.stack([np.linalg.norm(pts[:, :2], axis=1), np.arctan2(-pts[:, 1], pts[:, 0]) * 180. / np.pi, pts[:, 2],
                        pts[:, 3]], -1).astype(np.float32)
        r = gen.generate(cyl)
        fs.append(r["voxels"].sum(1) / np.maximum(r["num_points_per_voxel"], 1)[:, None])
        cs.append(np.pad(r["coordinates"], ((0, 0), (1, 0)), constant_values=b))
    feats, coords = np.concatenate(fs).astype(np.float32), np.concatenate(cs).astype(np.int32)
    feats = feats / np.array([70.0, 40.0, 3.0, 1.0], np.float32)
    want = mir.run(models_mirror.OracleBackend(), oracle_net.to_oracle_tensor(feats, coords, mir.sparse_shape, batch))
    ref_occ = ref_occ.cuda()
    with torch.no_grad():
        bd = ref_occ({"voxel_features": torch.from_numpy(feats).cuda(), "voxel_coords": torch.from_numpy(coords).cuda().float(),
                      "batch_size": batch})
    enc = bd["encoded_spconv_tensor"]
    report["backbone_occ_reference_forward"] = {
        "rows": int(enc.features.shape[0]),
        "indices_eq_oracle": bool(np.array_equal(enc.indices.cpu().numpy(), want["encoded"].indices)),
        "rel_err_vs_oracle": _rel(enc.features.cpu().numpy(), want["encoded"].features)}


def main():
    os.makedirs(OUT, exist_ok=True)
    import ref_loader
    report = {"reference_root": ref_loader.REF, "device": torch.cuda.get_device_name(0), "torch": torch.__version__}
    only = sys.argv[1:]
    for name, fn in (("masks", section_masks), ("inverse", section_inverse), ("inject", section_inject),
                     ("backbone", section_backbone)):
        if only and name not in only:
            continue
        try:
            fn(report)
        except Exception:
            report[name + "_error"] = traceback.format_exc()
    with open(os.path.join(OUT, "report.json"), "w") as fh:
        json.dump(report, fh, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
