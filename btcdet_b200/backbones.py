"""Sparse 3-D backbones assembled from the drop-in `spconv` layers.

`VoxelBackBone8x` is the topology BASELINE.json's configs[1] names.  The reference itself only
ships its occupancy-aware variants (btcdet/models/backbones_3d/spconv_backbone.py:630
VoxelBackBone8xOcc, :91 VoxelBackBoneDeconv, which import and run unchanged on the shim — see
tests/test_reference_import.py); this class restates the plain OpenPCDet-0.3.0 pyramid those
variants extend with the same block factory conventions
(post_act_block, spconv_backbone.py:7-43): SubM 4->16 'subm1' x2; SparseConv 16->32 s2 'spconv2' +
2 SubM 'subm2'; SparseConv 32->64 s2 'spconv3' + 2 SubM 'subm3'; SparseConv 64->64 s2 p(0,1,1)
'spconv4' + 2 SubM 'subm4'; SparseConv 64->128 k(3,1,1) s(2,1,1) 'spconv_down2'; every conv is
followed by BatchNorm1d(eps=1e-3, momentum=0.01) + ReLU.
"""
from functools import partial

import torch
import torch.nn as nn

import spconv


def conv_bn_relu(in_ch, out_ch, kernel_size, indice_key, stride=1, padding=0, conv_type="subm", norm_fn=None):
    if conv_type == "subm":
        conv = spconv.SubMConv3d(in_ch, out_ch, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == "spconv":
        conv = spconv.SparseConv3d(in_ch, out_ch, kernel_size, stride=stride, padding=padding, bias=False,
                                   indice_key=indice_key)
    else:
        raise NotImplementedError(conv_type)
    return spconv.SparseSequential(conv, norm_fn(out_ch), nn.ReLU())


class VoxelBackBone8x(nn.Module):
    def __init__(self, input_channels=4, grid_size=(1408, 1600, 40)):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = [grid_size[2] + 1, grid_size[1], grid_size[0]]
        blk = partial(conv_bn_relu, norm_fn=norm_fn)
        self.conv_input = blk(input_channels, 16, 3, "subm1", padding=1)
        self.conv1 = spconv.SparseSequential(blk(16, 16, 3, "subm1", padding=1))
        self.conv2 = spconv.SparseSequential(
            blk(16, 32, 3, "spconv2", stride=2, padding=1, conv_type="spconv"),
            blk(32, 32, 3, "subm2", padding=1), blk(32, 32, 3, "subm2", padding=1))
        self.conv3 = spconv.SparseSequential(
            blk(32, 64, 3, "spconv3", stride=2, padding=1, conv_type="spconv"),
            blk(64, 64, 3, "subm3", padding=1), blk(64, 64, 3, "subm3", padding=1))
        self.conv4 = spconv.SparseSequential(
            blk(64, 64, 3, "spconv4", stride=2, padding=(0, 1, 1), conv_type="spconv"),
            blk(64, 64, 3, "subm4", padding=1), blk(64, 64, 3, "subm4", padding=1))
        self.conv_out = blk(64, 128, (3, 1, 1), "spconv_down2", stride=(2, 1, 1), padding=0, conv_type="spconv")
        self.num_point_features = 128

    def forward(self, batch_dict):
        """batch_dict keys as in the reference: voxel_features [M,C], voxel_coords [M,4] (b,z,y,x), batch_size."""
        x = spconv.SparseConvTensor(features=batch_dict["voxel_features"], indices=batch_dict["voxel_coords"].int(),
                                    spatial_shape=self.sparse_shape, batch_size=batch_dict["batch_size"])
        x = self.conv_input(x)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        batch_dict.update({
            "encoded_spconv_tensor": out,
            "encoded_spconv_tensor_stride": 8,
            "multi_scale_3d_features": {"x_conv1": x_conv1, "x_conv2": x_conv2, "x_conv3": x_conv3,
                                        "x_conv4": x_conv4},
        })
        return batch_dict

    def layer_specs(self):
        """Flat (conv module, bn module) list in execution order — what engine.BackbonePlan compiles."""
        specs = []
        for seq in (self.conv_input, self.conv1, self.conv2, self.conv3, self.conv4, self.conv_out):
            for conv, bn in _iter_conv_bn(seq):
                specs.append((conv, bn))
        return specs


def _iter_conv_bn(module):
    mods = list(module.children())
    if mods and isinstance(mods[0], spconv.SparseConvolution):
        yield mods[0], (mods[1] if len(mods) > 1 and isinstance(mods[1], nn.BatchNorm1d) else None)
        return
    for m in mods:
        if isinstance(m, spconv.SparseSequential):
            yield from _iter_conv_bn(m)


def randomize_bn_(model, seed=0):
    """Random-init BatchNorm statistics/affine so that eval-mode BN is a non-trivial affine map."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm1d):
            with torch.no_grad():
                m.weight.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) * 0.5 + 0.75)
    return model
