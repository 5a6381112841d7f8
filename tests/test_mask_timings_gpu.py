"""Timing of the mask / target / injection kernels next to the torch restatement of the
reference's code run ON THE SAME GPU ("time stock-torch reference code for the mask path on the GPU as the
reference GPU number").  Config-3 shape: 2 scenes x 20 k points, kitti_car occupancy geometry.

The numbers are written to gpurun_out/mask_path_timings.json (when writable) and summarised under profiles/; there is
deliberately no speed assertion (timing asserts are flaky); parity of the same calls is asserted in test_occ_gpu.py,
test_box_masks_gpu.py and test_occ_inject_gpu.py."""
import json
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))
sys.path.insert(0, HERE)


def _time_us(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(out))


def test_mask_path_timings(cuda, oracle):
    import make_occ_golden
    import test_box_masks_gpu as TB
    import test_occ_inject_cpu as TI
    from btcdet_b200 import ops, synthetic as S
    from oracle import box_masks, occ_inject, occ_masks

    inp, geo = TB._case([21, 22], 20000, True, True)
    B = inp["batch_size"]
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
    rot = t.get("rot_z")
    cells = B * 9 * 157 * 209
    rows = []

    def add(name, ours_us, torch_us, alg_bytes):
        rows.append({"stage": name, "ours_us": round(ours_us, 1), "torch_restatement_us": round(torch_us, 1),
                     "speedup": round(torch_us / ours_us, 1), "alg_MB": round(alg_bytes / 1e6, 3),
                     "GBps": round(alg_bytes / (ours_us * 1e-6) / 1e9, 1)})
        assert ours_us > 0 and torch_us > 0

    # a5-a8 + general_cls_loss_mask: 16 B per valid point in, one byte per cell per emitted mask (4 masks)
    n_valid = int(t["voxel_num_points"].sum().item())
    ours = _time_us(lambda: ops.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, gf, gi, rot_z=rot))
    ref = _time_us(lambda: occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, geo, rot_z=rot),
                   reps=5, warm=1)
    add("occ_targets (a5-a8: valid points, dilation, spherical occlusion, filter)", ours, ref, 16 * n_valid + 4 * cells)

    # a9-a11: box-driven masks and residuals (3 masks + forebox label + 3 residual volumes of 12 B per cell)
    ref_occ = occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, geo, rot_z=rot)
    ours = _time_us(lambda: ops.occ_box_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], B, t["gt_boxes"],
                                                inp["gt_boxes_num"], gf, gi, box_mirr_flag=t["box_mirr_flag"],
                                                bm_points=t.get("bm_points"), rot_z=rot, num_class=1))
    ref = _time_us(lambda: box_masks.box_targets(ref_occ["valid_coords"], ref_occ["valid_feats"], t["gt_boxes"],
                                                 inp["gt_boxes_num"], t["box_mirr_flag"], B, geo, rot_z=rot, num_class=1,
                                                 bm_points=t.get("bm_points")), reps=5, warm=1)
    add("occ_box_targets (a9-a11: fore / mirror / best-match masks, mean residuals, forebox label)", ours, ref,
        16 * n_valid + (4 + 36) * cells)

    # a17-a20: occupancy-point selection + re-voxelisation + OccVFE
    case, geo_i = TI.make_case(3, dense=0.05, with_rot=True)
    case = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in case.items()}
    det_grid = [1408, 1600, 40]

    def ours_inject():
        vox, cnt, vc, sel = ops.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                             case["det_voxel_coords"], case["batch"], 0.3, 40000, geo_i.voxel_size,
                                             geo_i.point_cloud_range[:3], S.DET_VOXEL_SIZE, S.KITTI_RANGE, det_grid,
                                             rot_z=case["rot_z"])
        return ops.occ_vfe(vox, cnt, 4)

    def ref_inject():
        r = occ_inject.pass_occ_vox(case["probs"], case["res"], case["det_voxels"], case["det_voxel_num_points"],
                                    case["det_voxel_coords"], geo_i, S.DET_VOXEL_SIZE, det_grid, S.KITTI_RANGE, thresh=0.3,
                                    max_points=40000, rot_z=case["rot_z"])
        return r

    n_det = int(case["det_voxel_num_points"].sum().item())
    add("pass_occ_vox + occ_vfe (a17-a20: select, pseudo points, sorted re-voxelisation, VFE)", _time_us(ours_inject),
        _time_us(ref_inject, reps=5, warm=1), 16 * int(case["probs"].numel()) + 2 * 24 * n_det)

    rec = {"config": "config-3 shape: %d scenes x 20000 points, occ grid [9,157,209]; CUDA events, median of 10 (warm)" % B,
           "gpu": torch.cuda.get_device_name(0), "stages": rows}
    print(json.dumps(rec))
    out_dir = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "mask_path_timings.json"), "w") as fh:
            json.dump(rec, fh)
    except OSError:
        pass
