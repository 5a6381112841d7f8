"""GPU parity of the box-driven occupancy targets (btc_occ_box_targets / btc_occ_loss_maps).

Checker: oracle/box_masks.py (pinned bit-for-bit to the reference's create_voxel_res_label on CPU by
tests/test_box_masks_cpu.py) executed with torch on the SAME device, i.e. through the same torch.inverse (CUDA batched LU)
and einsum / matmul (K = 3 GEMM) the reference itself runs.  The kernels reproduce both operation by operation
(csrc/box_masks.cu lu_inverse + FMA chains; pinned to B200 dumps in tests/test_box_inverse_cpu.py), so
  * point labels, fore / mirrored / best-match masks and the forebox label are BIT-EXACT — including points on a box face
    and mirrored points on a bin edge (tests/test_reference_on_gpu.py shows the same against OccTargets3D.forward itself);
  * the per-cell mean residuals agree to 2e-5: the reference accumulates them with float atomics (scatter_add_, order
    dependent), the kernel in order-independent 2^-24 fixed point;
  * the loss-map algebra (a12) is exact on identical inputs."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
MARGIN = 1e-4


def _bm_points(inp, seed):
    rng = np.random.default_rng(seed)
    rows = []
    for b in range(inp["gt_boxes"].shape[0]):
        for bx in inp["gt_boxes"][b][:6]:
            u = rng.uniform(-0.6, 0.6, (60, 3)) * bx[3:6]
            c, s = np.cos(bx[6]), np.sin(bx[6])
            xy = np.stack([c * u[:, 0] - s * u[:, 1], s * u[:, 0] + c * u[:, 1]], 1) + bx[:2]
            rows.append(np.concatenate([np.full((60, 1), b), xy, u[:, 2:3] + bx[2]], 1))
    return np.concatenate(rows).astype(np.float32)


def _case(seeds, n_points, with_rot, with_bm, boxes_num=None):
    import make_occ_golden
    inp, geo = make_occ_golden.make_inputs(seeds, n_points=n_points, with_rot=with_rot)
    if with_rot:
        inp["rot_z"] = np.array([7.5, -11.25, 3.0, -2.0][:len(seeds)], np.float32)
    if with_bm:
        inp["bm_points"] = _bm_points(inp, 5)
    if boxes_num is not None:
        inp["gt_boxes_num"] = boxes_num
    return inp, geo


def _run(inp, geo, num_class=1):
    from btcdet_b200 import ops
    from oracle import box_masks, occ_masks
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
    B = inp["batch_size"]
    rot = t.get("rot_z")
    occ = ops.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, gf, gi, rot_z=rot)
    got = ops.occ_box_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, t["gt_boxes"], inp["gt_boxes_num"], gf, gi,
                              box_mirr_flag=t["box_mirr_flag"], bm_points=t.get("bm_points"), rot_z=rot, num_class=num_class,
                              want_point_label=True)
    assert int(got["status"].item()) == 0
    ref_occ = occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, geo, rot_z=rot)
    want = box_masks.box_targets(ref_occ["valid_coords"], ref_occ["valid_feats"], t["gt_boxes"], inp["gt_boxes_num"],
                                 t["box_mirr_flag"], B, geo, rot_z=rot, num_class=num_class, bm_points=t.get("bm_points"))
    # per valid point: smallest distance to a face of any box of its scene (in the reference's own box frame)
    vc, vf = ref_occ["valid_coords"], ref_occ["valid_feats"]
    margin = torch.full((vc.shape[0],), 1e9, device="cuda")
    boxes = t["gt_boxes"].clone()
    for b in range(B):
        sel = torch.nonzero(vc[:, 0] == b)[:, 0]
        nb = int(inp["gt_boxes_num"][b])
        if sel.numel() == 0 or nb == 0:
            continue
        q, _ = box_masks.box_frame(vf[sel, :3], boxes[b, :nb])
        d = (q.abs() - boxes[b, :nb, 3:6] * 0.5).abs().amin(dim=(1, 2))
        margin[sel] = d
    return t, occ, got, ref_occ, want, margin


def _frac_diff(a, b):
    a, b = a.bool(), b.bool()
    return int((a != b).sum()), max(int(b.sum()), 1)


@pytest.mark.parametrize("seeds,n,with_rot,with_bm", [([3, 4], 6000, True, True), ([11], 6000, False, False),
                                                      ([21, 22, 23], 20000, True, True)])
def test_box_targets_match_oracle_on_device(cuda, oracle, seeds, n, with_rot, with_bm):
    inp, geo = _case(seeds, n, with_rot, with_bm)
    t, occ, got, ref_occ, want, margin = _run(inp, geo)
    # a9 point labels: exact for every point, the ones within rounding distance of a box face included
    mask = ref_occ["voxel_point_mask"]
    assert torch.equal(got["point_label"][mask], want["point_label"])
    assert int((want["point_label"] > 0).sum()) > 30
    assert torch.equal(got["fore_voxelwise_mask"], want["fore_voxelwise_mask"])
    torch.testing.assert_close(got["fore_res_mtrx"], want["fore_res_mtrx"], rtol=0, atol=2e-5)
    assert int(want["fore_voxelwise_mask"].sum()) > 50
    # mirrored / template cells: the same cells; residuals up to the accumulation order
    for mk, rk in (("mirr_fore_voxelwise_mask", "mirr_res_mtrx"),) + ((("bm_voxelwise_mask", "bm_res_mtrx"),) if with_bm else ()):
        assert torch.equal(got[mk], want[mk]), mk
        assert int(want[mk].sum()) > 10
        torch.testing.assert_close(got[rk], want[rk], rtol=0, atol=2e-5)
    # a11 forebox label (2-D pre-filter on the z-mean centres + 3-D test): exact
    assert torch.equal(got["forebox_label"], want["forebox_label"])
    assert int((want["forebox_label"] > 0).sum()) > 100


def test_loss_maps_exact(cuda, oracle):
    """a12 on identical inputs (the kernel's own intermediate volumes): pure mask algebra -> bit-exact."""
    from btcdet_b200 import ops
    from oracle import box_masks
    inp, geo = _case([3, 4], 6000, True, True)
    t, occ, got, ref_occ, want, margin = _run(inp, geo)
    for weights in (None, {"occ_mirr_res_weight": 0.05, "occ_bm_res_weight": 0.02, "occ_mirr_cls_weight": 0.7,
                           "occ_bm_cls_weight": 0.4, "occ_neg_cls_weight": 0.9}):
        maps = ops.occ_loss_maps(occ, got, weights=weights, box_weight=0.2)
        w = dict(ops.DEFAULT_LOSS_WEIGHTS)
        w.update(weights or {})
        ow = {"fore_cls": w["occ_fore_cls_weight"], "mirr_cls": w["occ_mirr_cls_weight"], "bm_cls": w["occ_bm_cls_weight"],
              "neg_cls": w["occ_neg_cls_weight"], "fore_res": w["occ_fore_res_weight"], "mirr_res": w["occ_mirr_res_weight"],
              "bm_res": w["occ_bm_res_weight"], "box_weight": 0.2}
        box = {k: got[k] for k in ("fore_voxelwise_mask", "mirr_fore_voxelwise_mask", "bm_voxelwise_mask", "fore_res_mtrx",
                                   "mirr_res_mtrx", "bm_res_mtrx", "forebox_label")}
        ref = box_masks.loss_maps({"voxelwise_mask": occ["voxelwise_mask"], "general_cls_loss_mask": occ["general_cls_loss_mask"]},
                                  box, ow)
        for k in ("occ_fore_cls_mask", "occ_mirr_cls_mask", "occ_bm_cls_mask", "pos_mask", "general_reg_loss_mask"):
            assert torch.equal(maps[k].bool(), ref[k].bool()), k
        assert torch.equal(maps["general_cls_loss_mask_float"], ref["general_cls_loss_mask_float"])
        assert torch.equal(maps["general_reg_loss_mask_float"], ref["general_reg_loss_mask_float"])
        assert torch.equal(maps["res_mtrx"], ref["res_mtrx"])
        vm = occ["voxelwise_mask"]
        mirr_m = got["mirr_fore_voxelwise_mask"] * (1 - vm)
        bm_m = got["bm_voxelwise_mask"] * (1 - vm) * (1 - mirr_m)
        assert int(maps["pos_all_num"].item()) == int((got["fore_voxelwise_mask"] | mirr_m | bm_m).sum())
        assert int(maps["occ_mirr_cls_mask"].sum()) > 5 and float(maps["res_mtrx"].abs().sum()) > 0


def test_end_to_end_targets_close_to_reference_fixture(cuda, oracle):
    """ops.occ_training_targets vs the fixture produced by the reference's own code on CPU (other device, LU inverse):
    fore / pos masks agree up to the documented bin-edge sensitivity of the occlusion mask they are ANDed with."""
    from btcdet_b200 import ops
    import make_occ_golden
    g = np.load(os.path.join(HERE, "golden", "occ_masks.npz"))
    inp, geo = make_occ_golden.make_inputs([int(s) for s in g["seeds"]], n_points=int(g["n_points"]), with_rot=True)
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
    out = ops.occ_training_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], 2, t["gt_boxes"], inp["gt_boxes_num"],
                                   gf, gi, box_mirr_flag=t["box_mirr_flag"], rot_z=t["rot_z"])
    shape = tuple(g["shape"])
    for k, tol in (("fore_voxelwise_mask", 0.01), ("pos_mask", 0.15)):
        want = np.unpackbits(g["ref_" + k])[:int(np.prod(shape))].reshape(shape).astype(bool)
        diff = int((out[k].cpu().numpy().astype(bool) != want).sum())
        assert want.sum() > 50 and diff <= max(2, int(tol * want.sum())), (k, diff, int(want.sum()))


def test_box_targets_edge_cases(cuda, oracle):
    """A scene without boxes, no mirror flags, multi-class labels, BOX_WEIGHT == 1 (no forebox), no points."""
    from btcdet_b200 import ops
    inp, geo = _case([5, 6], 4000, False, False, boxes_num=[0, 7])
    inp["box_mirr_flag"][1, ::2] = 0.0
    inp["gt_boxes"][1, :, 7] = np.arange(inp["gt_boxes"].shape[1]) % 3 + 1
    t, occ, got, ref_occ, want, margin = _run(inp, geo, num_class=3)
    assert int(got["fore_voxelwise_mask"][0].sum()) == 0 and int(got["forebox_label"][0].abs().sum()) == 0
    assert int(got["mirr_fore_voxelwise_mask"][0].sum()) == 0
    mask = ref_occ["voxel_point_mask"]
    assert torch.equal(got["point_label"][mask], want["point_label"])
    assert int(want["point_label"].max()) >= 2                                   # multi-class labels survive
    assert torch.equal(got["mirr_fore_voxelwise_mask"], want["mirr_fore_voxelwise_mask"])
    assert torch.equal(got["forebox_label"], want["forebox_label"])
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    out = ops.occ_box_targets(t["voxels"][:0], t["voxel_coords"][:0], t["voxel_num_points"][:0], 2, t["gt_boxes"],
                              inp["gt_boxes_num"], gf, gi, want_forebox=False)
    assert out["forebox_label"] is None and int(out["fore_voxelwise_mask"].sum()) == 0
    maps = ops.occ_loss_maps(occ, got, box_weight=1.0)
    assert maps["forebox_label"] is None
    neg = occ["general_cls_loss_mask"].bool() & ~maps["pos_mask"].bool()
    assert torch.equal(maps["general_cls_loss_mask_float"][neg], torch.ones_like(maps["general_cls_loss_mask_float"][neg]))
