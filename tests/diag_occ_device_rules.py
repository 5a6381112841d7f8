"""Diagnostic (not collected by pytest): where the occlusion masks of the CUDA kernel, the torch restatement on the
GPU and the torch restatement on the CPU differ, and why (torch's tensor / python_scalar rule per device, libm ulps on
bin edges).  Lives under tests/ because it uses the oracle, which only test infrastructure may import.
  python tests/diag_occ_device_rules.py        # on a B200"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_occ_golden
from btcdet_b200 import ops
from oracle import occ_masks
inp, geo = make_occ_golden.make_inputs([3, 4], n_points=6000, with_rot=True)
gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern, geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
got = ops.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], 2, gf, gi, rot_z=t["rot_z"], want_sphere=True)
strict = occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], 2, geo, rot_z=t["rot_z"])
c = {k: torch.from_numpy(v) for k, v in inp.items() if isinstance(v, np.ndarray)}
cpu = occ_masks.occ_targets(c["voxels"], c["voxel_coords"], c["voxel_num_points"], 2, geo, rot_z=c["rot_z"])
occ_masks.CUDA_SCALAR_DIV = True
cpu2 = occ_masks.occ_targets(c["voxels"], c["voxel_coords"], c["voxel_num_points"], 2, geo, rot_z=c["rot_z"])
occ_masks.CUDA_SCALAR_DIV = False
for k in ["occ_voxelwise_mask", "general_cls_loss_mask"]:
    print(k, "kernel vs cpu-oracle with CUDA scalar-division rule:", int((got[k].cpu().bool() != cpu2[k].bool()).sum()))
for k in ["voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask"]:
    g = got[k].cpu().bool(); s = strict[k].cpu().bool(); cc = cpu[k].bool()
    print(k, "sum", int(g.sum()), int(s.sum()), "diff gpu-kernel vs gpu-oracle", int((g != s).sum()), "gpu-oracle vs cpu-oracle", int((s != cc).sum()), "kernel vs cpu", int((g != cc).sum()))
sm = got["sphere_map"].cpu()
so = strict["sphere_map"].cpu().clone()
# oracle's sphere map has bin 0 rewritten; compare bins >= 1 only
print("sphere bins>=1 diff", int((sm[..., 1:] != so[..., 1:]).sum()), int(sm[..., 1:].sum()), int(so[..., 1:].sum()))
d = (got["occ_voxelwise_mask"].cpu().bool() != strict["occ_voxelwise_mask"].cpu().bool()).nonzero()
print(d[:20])
