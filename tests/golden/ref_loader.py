"""Loads pieces of the reference (/root/reference, read-only) in THIS container so that their outputs
can pin the oracle and generate golden fixtures (the GPU box has no reference checkout).

The reference's model code cannot be imported as a package here: `btcdet/models/__init__.py` pulls in
detectors that need compiled CUDA extensions (THC-era, unbuildable — SURVEY §8c), and the mask code
hard-codes device="cuda".  This loader
  * registers empty namespace packages for `btcdet`, `btcdet.utils`, `btcdet.models`, ... and executes
    only the requested source files under their real dotted names (relative imports then resolve);
  * stubs `btcdet.ops.roiaware_pool3d.roiaware_pool3d_utils` (imported, never called on these paths);
  * while the reference code runs, rewrites device="cuda" to "cpu" in torch factory calls.
Nothing is copied: the files are executed where they lie.
"""
import contextlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
_STAGED = os.path.join(_ROOT, "oracle", "_ref", "reference_py")   # oracle/stage_reference.py (GPU box: O3 runs)


def _find_ref():
    for cand in (os.environ.get("BTC_REFERENCE_ROOT"), "/root/reference", _STAGED):
        if cand and os.path.isdir(os.path.join(cand, "btcdet")):
            return cand
    return "/root/reference"


REF = _find_ref()
ON_CUDA = False        # set by load_reference_modules(device="cuda"): the reference code then runs unchanged on the GPU


def available():
    return os.path.isdir(os.path.join(REF, "btcdet"))


def _ns(name, path):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(name, relpath):
    if name in sys.modules and getattr(sys.modules[name], "__file__", None):
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod


def load_reference_modules(device="cpu"):
    """Returns a dict of the reference modules on the mask / re-voxelisation path.  device="cuda": no rewriting at
    all — the reference's constructors allocate on the GPU exactly as they do in the reference's own runs (O3)."""
    global ON_CUDA
    assert available(), "reference checkout not present"
    ON_CUDA = str(device).startswith("cuda")
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)   # the drop-in `spconv` shim
    if not hasattr(np, "int"):
        np.int = int               # removed numpy alias used at occ_targets_template.py:51
    _ns("btcdet", os.path.join(REF, "btcdet"))
    _ns("btcdet.utils", os.path.join(REF, "btcdet/utils"))
    _ns("btcdet.ops", os.path.join(REF, "btcdet/ops"))
    _ns("btcdet.ops.roiaware_pool3d", os.path.join(REF, "btcdet/ops/roiaware_pool3d"))
    stub = types.ModuleType("btcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
    sys.modules[stub.__name__] = stub
    sys.modules["btcdet.ops.roiaware_pool3d"].roiaware_pool3d_utils = stub
    _ns("btcdet.models", os.path.join(REF, "btcdet/models"))
    _ns("btcdet.models.backbones_3d", os.path.join(REF, "btcdet/models/backbones_3d"))
    _ns("btcdet.models.occ_pnt", os.path.join(REF, "btcdet/models/occ_pnt"))
    _ns("btcdet.models.occ_pnt.occ_training_targets", os.path.join(REF, "btcdet/models/occ_pnt/occ_training_targets"))
    vis = types.ModuleType("btcdet.utils.vis_occ_utils")      # skimage-based drawing helpers: imported, never called here
    sys.modules[vis.__name__] = vis
    sys.modules["btcdet.utils"].vis_occ_utils = vis
    _ns("btcdet.models.backbones_3d.vfe", os.path.join(REF, "btcdet/models/backbones_3d/vfe"))
    mods = {}
    mods["common_utils"] = _load("btcdet.utils.common_utils", "btcdet/utils/common_utils.py")
    mods["coords_utils"] = _load("btcdet.utils.coords_utils", "btcdet/utils/coords_utils.py")
    mods["point_box_utils"] = _load("btcdet.utils.point_box_utils", "btcdet/utils/point_box_utils.py")
    with cuda_as_cpu():
        mods["spconv_backbone"] = _load("btcdet.models.backbones_3d.spconv_backbone",
                                        "btcdet/models/backbones_3d/spconv_backbone.py")
        mods["occ_targets_template"] = _load("btcdet.models.occ_pnt.occ_training_targets.occ_targets_template",
                                             "btcdet/models/occ_pnt/occ_training_targets/occ_targets_template.py")
        mods["occ_targets_3d"] = _load("btcdet.models.occ_pnt.occ_training_targets.occ_targets_3d",
                                       "btcdet/models/occ_pnt/occ_training_targets/occ_targets_3d.py")
        mods["add_occ_template"] = _load("btcdet.models.occ_pnt.add_occ_template", "btcdet/models/occ_pnt/add_occ_template.py")
        mods["pass_occ_vox"] = _load("btcdet.models.occ_pnt.pass_occ_vox", "btcdet/models/occ_pnt/pass_occ_vox.py")
        mods["vfe_template"] = _load("btcdet.models.backbones_3d.vfe.vfe_template", "btcdet/models/backbones_3d/vfe/vfe_template.py")
        mods["occ_vfe"] = _load("btcdet.models.backbones_3d.vfe.occ_vfe", "btcdet/models/backbones_3d/vfe/occ_vfe.py")
    return mods


_FACTORIES = ["zeros", "ones", "empty", "full", "tensor", "as_tensor", "arange", "zeros_like", "ones_like", "rand",
              "randint", "linspace", "eye", "meshgrid"]


@contextlib.contextmanager
def cuda_as_cpu():
    """Rewrite device="cuda" to "cpu" in torch factory calls while reference code executes (no-op in O3 mode)."""
    if ON_CUDA:
        yield
        return
    saved = {}

    def wrap(fn):
        def inner(*a, **kw):
            dev = kw.get("device")
            if dev is not None and str(dev).startswith("cuda"):
                kw["device"] = "cpu"
            return fn(*a, **kw)
        return inner

    for name in _FACTORIES:
        saved[name] = getattr(torch, name)
        setattr(torch, name, wrap(saved[name]))
    tensor_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **kw: self        # `.cuda()` in constructors (loss_utils.py:196)
    try:
        yield
    finally:
        torch.Tensor.cuda = tensor_cuda
        for name, fn in saved.items():
            setattr(torch, name, fn)


class Cfg(dict):
    """easydict.EasyDict stand-in (not installed here)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    @staticmethod
    def wrap(d):
        if isinstance(d, dict):
            return Cfg({k: Cfg.wrap(v) for k, v in d.items()})
        return d


def load_roi_head_modules(device="cpu"):
    """The reference's RoI head (`btcdet/models/roi_heads/conv_head.py`) with its import chain: the
    compiled extensions it reaches are replaced by this repo's drop-ins (`btcdet_b200.pointnet2_stack_cuda`,
    `btcdet_b200.iou3d_nms_cuda`), everything else is the reference's own Python executed where it lies.
    Returns {"conv_head": module, "ConvHead": class, "cfg": the ROI_HEAD section of btcdet_kitti_car.yaml (Cfg)}."""
    import yaml
    mods = load_reference_modules(device)
    from btcdet_b200 import iou3d_nms_cuda, pointnet2_stack_cuda
    _ns("btcdet.ops.iou3d_nms", os.path.join(REF, "btcdet/ops/iou3d_nms"))
    sys.modules["btcdet.ops.iou3d_nms.iou3d_nms_cuda"] = iou3d_nms_cuda
    sys.modules["btcdet.ops.iou3d_nms"].iou3d_nms_cuda = iou3d_nms_cuda
    _ns("btcdet.ops.pointnet2", os.path.join(REF, "btcdet/ops/pointnet2"))
    _ns("btcdet.ops.pointnet2.pointnet2_stack", os.path.join(REF, "btcdet/ops/pointnet2/pointnet2_stack"))
    sys.modules["btcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda"] = pointnet2_stack_cuda
    sys.modules["btcdet.ops.pointnet2.pointnet2_stack"].pointnet2_stack_cuda = pointnet2_stack_cuda
    _ns("btcdet.models.roi_heads", os.path.join(REF, "btcdet/models/roi_heads"))
    _ns("btcdet.models.roi_heads.target_assigner", os.path.join(REF, "btcdet/models/roi_heads/target_assigner"))
    _ns("btcdet.models.model_utils", os.path.join(REF, "btcdet/models/model_utils"))
    with cuda_as_cpu():
        _load("btcdet.ops.iou3d_nms.iou3d_nms_utils", "btcdet/ops/iou3d_nms/iou3d_nms_utils.py")
        _load("btcdet.ops.pointnet2.pointnet2_stack.pointnet2_utils", "btcdet/ops/pointnet2/pointnet2_stack/pointnet2_utils.py")
        _load("btcdet.ops.pointnet2.pointnet2_stack.pointnet2_modules", "btcdet/ops/pointnet2/pointnet2_stack/pointnet2_modules.py")
        _load("btcdet.utils.box_utils", "btcdet/utils/box_utils.py")
        _load("btcdet.utils.box_coder_utils", "btcdet/utils/box_coder_utils.py")
        _load("btcdet.utils.loss_utils", "btcdet/utils/loss_utils.py")
        _load("btcdet.models.model_utils.model_nms_utils", "btcdet/models/model_utils/model_nms_utils.py")
        _load("btcdet.models.roi_heads.target_assigner.proposal_target_layer",
              "btcdet/models/roi_heads/target_assigner/proposal_target_layer.py")
        _load("btcdet.models.roi_heads.roi_head_template", "btcdet/models/roi_heads/roi_head_template.py")
        ch = _load("btcdet.models.roi_heads.conv_head", "btcdet/models/roi_heads/conv_head.py")
    with open(os.path.join(REF, "tools/cfgs/model_configs/btcdet_kitti_car.yaml")) as fh:
        cfg = Cfg.wrap(yaml.safe_load(fh))
    mods.update({"conv_head": ch, "ConvHead": ch.ConvHead, "cfg": cfg.MODEL.ROI_HEAD})
    return mods


def build_conv_head(mods, voxel_size, point_cloud_range, num_rawpoint_features=4):
    """ConvHead exactly as `Detector3DTemplate.build_roi_head` constructs it (detector3d_template.py:344-358)."""
    import copy
    with cuda_as_cpu():
        return mods["ConvHead"](input_channels=128, model_cfg=copy.deepcopy(mods["cfg"]), num_class=1,
                                det_voxel_size=voxel_size, point_cloud_range=point_cloud_range,
                                num_rawpoint_features=num_rawpoint_features, pre_conv_num_bev_features=256)
