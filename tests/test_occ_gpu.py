"""GPU parity of the fused occupancy / occlusion mask kernels (btc_occ_targets).

Strict reference: oracle/occ_masks.py executed on the SAME device (same CUDA libm, same fp32 op order, same
CUDA rule for tensor / python_scalar) — bit-exact, every mask.
Cross-device: the reference algorithm back-projects sphere-bin lower corners that land exactly on cylinder-bin
edges, so results are one-ulp sensitive.  The CPU oracle (pinned bit-for-bit to the reference's code on CPU,
tests/test_occ_oracle_cpu.py) differs from a CUDA run of the same torch code in ~10 % of the occluded cells
because torch divides by python scalars differently per device; with oracle.occ_masks.CUDA_SCALAR_DIV the CPU run
follows the CUDA rule and only libm-ulp edge cases remain (~5 % of the occluded cells on the fixture scene, bounded
at 10 % below; tests/diag_occ_device_rules.py prints the breakdown)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
MASKS = ["voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask"]


def _run(inp, geo, device="cuda"):
    from btcdet_b200 import ops
    from oracle import occ_masks
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if isinstance(v, np.ndarray)}
    got = ops.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], inp["batch_size"], gf, gi,
                          rot_z=t.get("rot_z"))
    strict = occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], inp["batch_size"], geo,
                                   rot_z=t.get("rot_z"))
    c = {k: torch.from_numpy(v) for k, v in inp.items() if isinstance(v, np.ndarray)}
    occ_masks.CUDA_SCALAR_DIV = True
    try:
        cpu = occ_masks.occ_targets(c["voxels"], c["voxel_coords"], c["voxel_num_points"], inp["batch_size"], geo,
                                    rot_z=c.get("rot_z"))
    finally:
        occ_masks.CUDA_SCALAR_DIV = False
    return got, strict, cpu


@pytest.mark.parametrize("seeds,with_rot,n", [([3, 4], True, 6000), ([11], False, 6000), ([21, 22, 23], True, 20000)])
def test_occ_masks_bit_exact_on_device(cuda, oracle, seeds, with_rot, n):
    import make_occ_golden
    inp, geo = make_occ_golden.make_inputs(seeds, n_points=n, with_rot=with_rot)
    if with_rot and len(seeds) == 3:
        inp["rot_z"] = np.array([7.5, -11.25, 3.0], np.float32)
    got, strict, cpu = _run(inp, geo)
    for k in MASKS:
        g = got[k].cpu().numpy().astype(bool)
        assert np.array_equal(g, strict[k].cpu().numpy().astype(bool)), k          # same device: bit-exact
        diff = int((g != cpu[k].numpy().astype(bool)).sum())
        exact = k in ("voxelwise_mask", "vcc_mask")                                 # integer-only paths
        assert diff <= (0 if exact else int(0.10 * max(int(g.sum()), 1))), (k, diff)   # other libm: bin-edge ulps only
    assert int(got["occ_voxelwise_mask"].sum()) > 1000 and int(got["vcc_mask"].sum()) > 1000


def test_occ_masks_vs_reference_fixture(cuda, oracle):
    """Fixture = the reference's own code on CPU: the integer-only masks agree exactly; the occluded set agrees up
    to the documented device-dependent scalar-division rule (see module docstring)."""
    import make_occ_golden
    g = np.load(os.path.join(HERE, "golden", "occ_masks.npz"))
    inp, geo = make_occ_golden.make_inputs([int(s) for s in g["seeds"]], n_points=int(g["n_points"]), with_rot=True)
    got, _, _ = _run(inp, geo)
    shape = tuple(g["shape"])
    for k in MASKS:
        want = np.unpackbits(g["ref_" + k])[:int(np.prod(shape))].reshape(shape).astype(bool)
        diff = int((got[k].cpu().numpy().astype(bool) != want).sum())
        if k in ("voxelwise_mask", "vcc_mask"):
            assert diff == 0, (k, diff)
        else:
            assert diff <= int(0.15 * want.sum()), (k, diff)


def test_occ_masks_empty_and_edge(cuda):
    from btcdet_b200 import ops
    from oracle import occ_masks
    geo = occ_masks.OccGeometry()
    gf, gi = ops.occ_geometry_arrays(geo.voxel_size, geo.point_cloud_range, geo.support_sphere_range, geo.dist_kern,
                                     geo.half_x, geo.empt_sur_thresh, geo.det_point_cloud_range)
    vox = torch.zeros(1, 12, 4, device="cuda")
    vox[0, 0, :3] = torch.tensor([2.3, -40.5, -2.5])
    out = ops.occ_targets(vox, torch.tensor([[0, 0, 0, 0]], device="cuda"), torch.tensor([1], device="cuda"), 1, gf, gi)
    vcc = out["vcc_mask"][0]
    assert int(vcc.sum()) == 75 and bool(vcc[:3, :5, :5].all())            # border clamp KAT (v)
    empty = ops.occ_targets(torch.zeros(0, 12, 4, device="cuda"), torch.zeros(0, 4, dtype=torch.int32, device="cuda"),
                            torch.zeros(0, dtype=torch.int32, device="cuda"), 2, gf, gi)
    assert all(int(v.sum()) == 0 for v in empty.values())
