"""Pins oracle/occ_masks.py: bit-exact against the reference's OWN code executed on
CPU in this container (skipped where /root/reference is absent) and against the committed golden fixture that
was generated from that reference run (tests/golden/make_occ_golden.py)."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_occ_golden  # noqa: E402
import ref_loader  # noqa: E402

MASKS = ["voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask"]


def _oracle(inp, geo):
    from oracle import occ_masks
    t = {k: torch.from_numpy(v) for k, v in inp.items() if isinstance(v, np.ndarray)}
    return occ_masks.occ_targets(t["voxels"], t["voxel_coords"], t["voxel_num_points"], inp["batch_size"], geo,
                                 rot_z=t.get("rot_z"))


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
@pytest.mark.parametrize("seeds,with_rot", [([3, 4], True), ([11], False)])
def test_oracle_equals_reference_code_on_cpu(seeds, with_rot):
    inp, geo = make_occ_golden.make_inputs(seeds, with_rot=with_rot)
    ref = make_occ_golden.run_reference(inp, geo)
    got = _oracle(inp, geo)
    for k in MASKS:
        np.testing.assert_array_equal(got[k].numpy().astype(bool), ref[k].astype(bool), err_msg=k)
    # the mask algebra of prepare_cls_loss_map with the reference's own box masks plugged in
    fore = torch.from_numpy(ref["fore_voxelwise_mask"])
    g = got["general_cls_loss_mask"]
    np.testing.assert_array_equal((fore & g).numpy(), ref["occ_fore_cls_mask"])


def test_oracle_equals_committed_golden_fixture():
    g = np.load(os.path.join(HERE, "golden", "occ_masks.npz"))
    inp, geo = make_occ_golden.make_inputs([int(s) for s in g["seeds"]], n_points=int(g["n_points"]), with_rot=True)
    np.testing.assert_array_equal(inp["rot_z"], g["rot_z"])
    got = _oracle(inp, geo)
    shape = tuple(g["shape"])
    for k in MASKS:
        want = np.unpackbits(g["ref_" + k])[:int(np.prod(shape))].reshape(shape).astype(bool)
        np.testing.assert_array_equal(got[k].numpy().astype(bool), want, err_msg=k)


def test_kat_border_clamp_and_one_point_occlusion():
    """KATs (v) and (vi) of SURVEY §8c."""
    from oracle import occ_masks
    geo = occ_masks.OccGeometry()
    nx, ny, nz = geo.grid_size
    assert (nx, ny, nz) == (209, 157, 9) and geo.sphere_grid_size == [214, 157, 49] and geo.concede_x == 2
    # one voxel in the (z=0, y=0, x=0) corner: the 5x9x5 window is clamped ONTO the border, x in [0, 4]
    vox = torch.zeros(1, 12, 4)
    vox[0, 0, :3] = torch.tensor([2.3, -40.5, -2.5])     # (rho, phi, z) inside cell (0,0,0)
    out = occ_masks.occ_targets(vox, torch.tensor([[0., 0, 0, 0]]), torch.tensor([1.]), 1, geo)
    vcc = out["vcc_mask"][0]
    assert int(vcc.sum()) == 3 * 5 * 5 and bool(vcc[:3, :5, :5].all())
    # a single return at ~20 m straight ahead occludes every range bin at or behind it in its (el, az) column
    vox[0, 0, :3] = torch.tensor([20.0, 0.1, -1.0])
    cell = torch.tensor([[0., 4, 78, 55]])
    out = occ_masks.occ_targets(vox, cell, torch.tensor([1.]), 1, geo)
    sm = out["sphere_map"][0]
    el, az, r = [int(v) for v in torch.nonzero(sm)[0]]
    assert int(sm.sum()) == 1 and r == int((np.float32(np.sqrt(401.0)) - np.float32(2.24)) / np.float32(0.32))
