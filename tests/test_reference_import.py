"""The reference's own model files import and BUILD unchanged on the drop-in `spconv` package
(north_star: "so btcdet/models builds unchanged").  Runs only where /root/reference exists (this
container); the GPU box has no reference checkout."""
import importlib.util
import os
import sys
import types

import pytest

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class _Cfg(dict):
    """Stand-in for easydict.EasyDict (not installed here): attribute access, AttributeError when absent."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def test_reference_backbones_build_on_the_shim():
    import spconv  # the shim, found first on sys.path via conftest
    assert spconv.__version__.endswith("btcdet_b200")
    sb = _load("ref_spconv_backbone", "btcdet/models/backbones_3d/spconv_backbone.py")
    occ_bb = sb.VoxelBackBoneDeconv(_Cfg(), input_channels=4, grid_size=[209, 157, 9])
    assert occ_bb.sparse_shape == [9, 157, 209]
    assert isinstance(occ_bb.deconv4[0][0], spconv.SparseConvTranspose3d)
    cfg = _Cfg(OCC_CONV_TYPE=['identity', 'maxpool'], OCC_CONV_EXECUTE=[False, True],
               OUT_FEAT_TYPE=['None', 'None', 'None', 'None', 'big_bev_combine'])
    import numpy as np
    det_bb = sb.VoxelBackBone8xOcc(cfg, input_channels=6, grid_size=np.array([1408, 1600, 40]),
                                   original_num_rawpoint_features=4)
    assert list(det_bb.sparse_shape) == [41, 1600, 1408]
    assert isinstance(det_bb.occ_conv2[0][0], spconv.SparseMaxPool3d)
    assert det_bb.conv1[0].weight.shape == (3, 3, 3, 6, 16)       # spconv-1.2.1 weight layout
    keys = det_bb.state_dict().keys()
    assert "conv2.0.0.weight" in keys and "down_combine.1.1.running_mean" in keys
    fix = sb.fixSparseConv3d(2, 2, 3, stride=2, padding=1, defaultvalue=1.0 / 27)
    assert float(fix.weight.min()) == float(fix.weight.max()) == pytest.approx(1.0 / 27)
    block = sb.SparseBasicBlock(16, 16, norm_fn=lambda c: __import__("torch").nn.BatchNorm1d(c), indice_key="r")
    assert isinstance(block, spconv.SparseModule)


def test_mirrors_have_the_reference_state_dict_layout():
    """tests/models_mirror.py (used by the -m gpu tests, where no reference exists) is checked here against the
    real classes: identical parameter / buffer names and shapes."""
    import numpy as np
    from tests import models_mirror
    sb = _load("ref_spconv_backbone", "btcdet/models/backbones_3d/spconv_backbone.py")
    ref_occ = sb.VoxelBackBoneDeconv(_Cfg(), input_channels=4, grid_size=[209, 157, 9])
    mir_occ = models_mirror.OccBackboneMirror(4, (209, 157, 9))
    ref_sd = {k: tuple(v.shape) for k, v in ref_occ.state_dict().items()}
    mir_sd = {k: tuple(v.shape) for k, v in mir_occ.state_dict().items() if not k.startswith(("conv_cls", "conv_res"))}
    assert ref_sd == mir_sd and mir_occ.sparse_shape == ref_occ.sparse_shape
    cfg = _Cfg(OCC_CONV_TYPE=['identity', 'maxpool'], OCC_CONV_EXECUTE=[False, True],
               OUT_FEAT_TYPE=['None', 'None', 'None', 'None', 'big_bev_combine'])
    ref_det = sb.VoxelBackBone8xOcc(cfg, input_channels=6, grid_size=np.array([1408, 1600, 40]),
                                    original_num_rawpoint_features=4)
    mir_det = models_mirror.DetBackboneMirror(6, 4, (1408, 1600, 40))
    assert {k: tuple(v.shape) for k, v in ref_det.state_dict().items()} == \
        {k: tuple(v.shape) for k, v in mir_det.state_dict().items()}
    assert list(ref_det.sparse_shape) == mir_det.sparse_shape
    # the real classes' parameters load into the mirror unchanged (same names, same layouts)
    mir_det.load_state_dict(ref_det.state_dict())
