"""O3: the REFERENCE'S OWN code executed on the B200 against this library.

The reference's Python sources are staged, unmodified, under the git-ignored oracle/_ref/reference_py/ by
`oracle/stage_reference.py` (run by __graft_entry__.build() where /root/reference exists) and travel to the GPU box with
the snapshot; these tests skip when the staged copy is absent.  What runs here is not a mirror and not a restatement:
  * `VoxelBackBone8xOcc.forward` and `VoxelBackBoneDeconv.forward` (spconv_backbone.py:936-1019, 138-203) are called
    unchanged; every spconv layer they touch is this repo's shim + CUDA library;
  * `OccTargets3D.forward`, `PassOccVox.forward`, `OccVFE.forward` are called unchanged on CUDA tensors and compared with
    the fused kernels (same device, same CUDA libm, same ATen kernels, same torch.inverse).
The comparison logic lives in tools/o3_reference_cuda.py (which also writes the *_cuda.npz fixtures)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not staged (oracle/stage_reference.py)")


@needs_ref
def test_reference_backbone_forwards_run_on_the_shim_and_match_the_oracle(cuda, oracle):
    import o3_reference_cuda as O3
    rep = {}
    O3.section_backbone(rep)
    det = rep["backbone_det_reference_forward"]
    for k in ("out", "x_combine"):
        assert det[k]["indices_eq_oracle"] and det[k]["indices_eq_mirror"], (k, det[k])
        assert det[k]["features_eq_mirror"], k                   # the test mirror IS the reference dataflow, bit for bit
        assert det[k]["rel_err_vs_oracle"] < 1e-4, (k, det[k])
        assert det[k]["rows"] > 1000
    occ = rep["backbone_occ_reference_forward"]
    assert occ["indices_eq_oracle"] and occ["rel_err_vs_oracle"] < 1e-4 and occ["rows"] > 100000


@needs_ref
def test_reference_pass_occ_vox_and_occ_vfe_on_cuda(cuda, oracle):
    import o3_reference_cuda as O3
    os.makedirs(O3.OUT, exist_ok=True)
    rep = {}
    O3.section_inject(rep)
    for seed in (1, 2, 3):
        r = rep["inject_seed%d" % seed]
        assert r["coords_equal"] and r["counts_equal"], (seed, r)              # a19: bit-exact voxel set, order, counts
        assert r["occ_voxel_features_equal"], (seed, r)                         # a20: max over slots, exact
        assert r["voxel_features_maxdiff"] <= 1e-6, (seed, r)                   # a20: means (summation order only)
        if seed != 3:                                                           # no top-k branch: a17-a18 bit-exact
            assert r["occ_xyz_equal"] and r["occ_b_equal"], (seed, r)


@needs_ref
def test_reference_occ_targets_on_cuda(cuda, oracle):
    """rows a5-a12 against OccTargets3D.forward run unchanged on the same GPU."""
    import o3_reference_cuda as O3
    os.makedirs(O3.OUT, exist_ok=True)
    rep = {}
    O3.section_masks(rep)
    for tag in ("a", "b", "c"):
        r = rep["masks_" + tag]
        # every mask of rows a5-a12, the forebox label included: bit-exact against the reference on the same device
        for k in O3.U8:
            assert r[k]["diff"] == 0, (tag, k, r[k])
        assert r["fore_voxelwise_mask"]["ref_set"] > 30 and r["forebox_label"]["ref_set"] > 1000
        assert r["general_cls_loss_mask_float"]["cells_differ"] == 0 and r["general_reg_loss_mask_float"]["cells_differ"] == 0
        # mean residuals: the reference's scatter_add_ is order dependent (float atomics); agreement to 2e-5
        assert r["res_mtrx"]["max_abs"] <= 2e-5, (tag, r["res_mtrx"])


def test_occ_targets_vs_cuda_generated_reference_fixture(cuda, oracle):
    """tests/golden/occ_masks_cuda.npz = OccTargets3D.forward of the reference run unchanged on a B200
    (tools/o3_reference_cuda.py, case "a": seeds 3 / 4, 6000 points, rot_z, template points).  Runs without the staged
    reference: every mask identical, res_mtrx to 2e-5."""
    import o3_reference_cuda as O3
    from btcdet_b200 import ops
    g = np.load(os.path.join(HERE, "golden", "occ_masks_cuda.npz"))
    inp, geo = O3._mask_case([int(v) for v in g["seeds"]], int(g["n_points"]), bool(g["with_rot"]), bool(g["with_bm"]))
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
    got = ops.occ_training_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], inp["batch_size"], t["gt_boxes"],
                                   inp["gt_boxes_num"], gf, gi, box_mirr_flag=t["box_mirr_flag"], bm_points=t.get("bm_points"),
                                   rot_z=t.get("rot_z"))
    shape = tuple(int(v) for v in g["shape"])
    n = int(np.prod(shape))
    for k in O3.U8:
        if "ref_" + k not in g.files or got.get(k) is None:
            continue
        if k == "forebox_label":
            np.testing.assert_array_equal(got[k].cpu().numpy(), g["ref_" + k], err_msg=k)
        else:
            want = np.unpackbits(g["ref_" + k])[:n].reshape(shape).astype(bool)
            np.testing.assert_array_equal(got[k].cpu().numpy() != 0, want, err_msg=k)
    for k, tol in (("general_cls_loss_mask_float", 0.0), ("general_reg_loss_mask_float", 0.0), ("res_mtrx", 2e-5)):
        want = np.zeros(int(np.prod(got[k].shape)), np.float32)
        want[g["ref_" + k + "_idx"]] = g["ref_" + k + "_val"]
        assert float(np.abs(got[k].cpu().numpy().reshape(-1) - want).max()) <= tol, k
