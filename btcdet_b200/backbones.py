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


def _post_act_block(cin, cout, k, indice_key, stride=1, padding=0, conv_type="subm"):
    """post_act_block (spconv_backbone.py:7-43): conv (no bias) + BatchNorm1d(eps 1e-3, momentum 0.01) + ReLU."""
    norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
    if conv_type == "subm":
        conv = spconv.SubMConv3d(cin, cout, k, bias=False, indice_key=indice_key)
    elif conv_type == "spconv":
        conv = spconv.SparseConv3d(cin, cout, k, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    elif conv_type == "spdeconv":
        conv = spconv.SparseConvTranspose3d(cin, cout, k, stride=stride, padding=padding, bias=False, indice_key=indice_key)
    else:
        raise NotImplementedError(conv_type)
    return spconv.SparseSequential(conv, norm(cout), nn.ReLU())


class OccBackbone(nn.Module):
    """The reference's occupancy backbone VoxelBackBoneDeconv (spconv_backbone.py:91-203) with identical attribute
    names, layer table and state-dict layout: conv1 SparseConv3d 4->16 (dilating, p1); conv2 SparseConv3d 16->32 s2 +
    SubM 32; conv3 SparseConv3d 32->64 s2 + SubM 64; deconv4 SparseConvTranspose3d 64->32 s2 + SubM 32; deconv5
    SparseConvTranspose3d 32->32 s2 + SubM 32 on the cylindrical grid [9,157,209].  (The reference class itself runs
    unchanged on the shim — tests/test_reference_on_gpu.py; this copy exists because bench.py / the GPU box must not
    depend on a reference checkout.)"""

    def __init__(self, input_channels=4, grid_size=(209, 157, 9)):
        super().__init__()
        self.sparse_shape = list(grid_size[::-1])
        c = [16, 32, 64]
        blk = _post_act_block
        self.conv1 = spconv.SparseSequential(blk(input_channels, c[0], 3, "spconv1", padding=1, conv_type="spconv"))
        self.conv2 = spconv.SparseSequential(blk(c[0], c[1], 3, "spconv2", stride=2, padding=1, conv_type="spconv"),
                                             blk(c[1], c[1], 3, "subm2", padding=1))
        self.conv3 = spconv.SparseSequential(blk(c[1], c[2], 3, "spconv3", stride=2, padding=1, conv_type="spconv"),
                                             blk(c[2], c[2], 3, "subm3", padding=1))
        self.deconv4 = spconv.SparseSequential(blk(c[2], c[1], 3, "spconv4", stride=2, padding=1, conv_type="spdeconv"),
                                               blk(c[1], c[1], 3, "subm4", padding=1))
        self.deconv5 = spconv.SparseSequential(blk(c[1], c[1], 3, "spconv5", stride=2, padding=1, conv_type="spdeconv"),
                                               blk(c[1], c[1], 3, "subm5", padding=1))
        self.num_point_features = c[1]

    def forward(self, batch_dict):
        x = spconv.SparseConvTensor(features=batch_dict["voxel_features"], indices=batch_dict["voxel_coords"].int(),
                                    spatial_shape=self.sparse_shape, batch_size=batch_dict["batch_size"],
                                    n_dev=batch_dict.get("voxel_n_dev"))      # static mode when the count is on the device
        x = self.deconv5(self.deconv4(self.conv3(self.conv2(self.conv1(x)))))
        batch_dict.update({"encoded_spconv_tensor": x, "encoded_spconv_tensor_stride": 1})
        return batch_dict


class OccHead(nn.Module):
    """OccHead3D convolutions (occ_head_3D.py:25-31, softmax, class agnostic, REG): SubMConv3d 32->2 with bias
    ('cls_ind') and SubMConv3d 32->3 without ('res_ind'); forward = :41-52 (dense, softmax[:, -1] * mask)."""

    def __init__(self, input_channels=32, num_class=1, res_num_dim=3):
        super().__init__()
        self.conv_cls = spconv.SparseSequential(spconv.SubMConv3d(input_channels, num_class + 1, 3, padding=1, bias=True,
                                                                  indice_key="cls_ind"))
        self.conv_res = spconv.SparseSequential(spconv.SubMConv3d(input_channels, num_class * res_num_dim, 3, padding=1,
                                                                  bias=False, indice_key="res_ind"))

    def forward(self, batch_dict):
        enc = batch_dict["encoded_spconv_tensor"]
        cls = self.conv_cls(enc)
        if batch_dict.get("fused_occ_head", False) and not torch.is_grad_enabled():
            # inference: probability volume straight from the sparse rows (btc_occ_head_prob) — no dense
            # logits, no softmax / product passes over [B, 2, 9, 157, 209]
            from . import ops
            nz, ny, nx = [int(v) for v in cls.spatial_shape]
            batch_dict["batch_pred_occ_prob"] = ops.occ_head_prob(cls.features, cls.indices, batch_dict["batch_size"], (nx, ny, nz),
                                                                  batch_dict["general_cls_loss_mask"], n_dev=cls.n_dev)
        else:
            logits = cls.dense()
            prob = torch.softmax(logits, dim=1)[:, -1:, ...]
            batch_dict["pred_occ_logit"] = logits
            batch_dict["batch_pred_occ_prob"] = prob[:, -1, ...] * batch_dict["general_cls_loss_mask"]
        batch_dict["pred_sem_residuals"] = self.conv_res(enc).dense()
        return batch_dict


class DetBackboneOcc(nn.Module):
    """The reference's detection backbone VoxelBackBone8xOcc (spconv_backbone.py:630-1019) in the shipped kitti_car
    configuration (OCC_CONV_TYPE ['identity','maxpool'], OCC_CONV_EXECUTE [False, True], OUT_FEAT_TYPE
    [None x4, 'big_bev_combine']; yaml :175-179), same attribute names / state-dict layout: SubM 6->16 + SubM 16 'subm1';
    SparseConv 16->32 s2 'spconv2' with the max-pooled occupancy side channel concatenated (34 ch) + 2 SubM 'subm2';
    SparseConv 32->64 s2 'spconv3' + 2 SubM; SparseConv 64->64 s2 p(0,1,1) 'spconv4' + 2 SubM; conv_out (3,1,1) s(2,1,1);
    res_combine (:905-918): down2 / down3 reuse the cached 'spconv3' / 'spconv4' rulebooks, squeezeBev dense + gather,
    down_combine 256->128->128."""

    def __init__(self, input_channels=6, raw_channels=4, grid_size=(1408, 1600, 40)):
        super().__init__()
        self.sparse_shape = [grid_size[2] + 1, grid_size[1], grid_size[0]]
        ch = [16, 32, 64, 64, 128]
        occ = input_channels - raw_channels
        self.occ_code_num = occ
        norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        blk = _post_act_block
        self.occ_conv2 = spconv.SparseSequential(spconv.SparseSequential(spconv.SparseMaxPool3d(3, stride=2, padding=1)))
        self.conv1 = spconv.SparseSequential(spconv.SubMConv3d(input_channels, ch[0], 3, padding=1, bias=False,
                                                               indice_key="subm1"), norm(ch[0]), nn.ReLU())
        self.conv1_combine = spconv.SparseSequential(blk(ch[0], ch[0], 3, "subm1", padding=1))
        self.conv2 = spconv.SparseSequential(blk(ch[0], ch[1], 3, "spconv2", stride=2, padding=1, conv_type="spconv"))
        self.conv2_combine = spconv.SparseSequential(blk(ch[1] + occ, ch[1], 3, "subm2", padding=1),
                                                     blk(ch[1], ch[1], 3, "subm2", padding=1))
        self.conv3 = spconv.SparseSequential(blk(ch[1], ch[2], 3, "spconv3", stride=2, padding=1, conv_type="spconv"))
        self.conv3_combine = spconv.SparseSequential(blk(ch[2], ch[2], 3, "subm3", padding=1),
                                                     blk(ch[2], ch[2], 3, "subm3", padding=1))
        self.conv4 = spconv.SparseSequential(blk(ch[2], ch[3], 3, "spconv4", stride=2, padding=(0, 1, 1), conv_type="spconv"))
        self.conv4_combine = spconv.SparseSequential(blk(ch[3], ch[3], 3, "subm4", padding=1),
                                                     blk(ch[3], ch[3], 3, "subm4", padding=1))
        self.conv_out = spconv.SparseSequential(spconv.SparseConv3d(ch[3], ch[4], (3, 1, 1), stride=(2, 1, 1), padding=0,
                                                                    bias=False, indice_key="spconv_down2"),
                                                norm(ch[4]), nn.ReLU())
        self.down2 = spconv.SparseSequential(blk(ch[1], ch[1], 3, "spconv3", stride=2, padding=1, conv_type="spconv"),
                                             blk(ch[1], ch[2], 3, "spconv4", stride=2, padding=(0, 1, 1), conv_type="spconv"))
        self.down3 = spconv.SparseSequential(blk(ch[2], ch[2], 3, "spconv4", stride=2, padding=(0, 1, 1), conv_type="spconv"))
        self.squeezeBev = spconv.SparseSequential(blk(ch[4], ch[3], (2, 1, 1), "subm_down2", stride=(2, 1, 1), padding=0,
                                                      conv_type="spconv"))
        self.down_combine = spconv.SparseSequential(blk(ch[2] * 2 + ch[3] * 2, ch[3] * 2, 3, "subm4", padding=1),
                                                    blk(ch[3] * 2, ch[3] * 2, 3, "subm4", padding=1))
        self.num_point_features = 128

    def forward(self, batch_dict):
        coords = batch_dict["voxel_coords"].int()
        B = batch_dict["batch_size"]
        n_dev = batch_dict.get("voxel_n_dev")             # static mode when the count is on the device
        x = spconv.SparseConvTensor(features=batch_dict["voxel_features"], indices=coords, spatial_shape=self.sparse_shape,
                                    batch_size=B, n_dev=n_dev)
        x_conv1 = self.conv1(x)
        occ_in = spconv.SparseConvTensor(features=batch_dict["occ_voxel_features"], indices=coords,
                                         spatial_shape=self.sparse_shape, batch_size=B, n_dev=n_dev)   # fresh rulebook cache (:959-964)
        x_conv1 = self.conv1_combine(x_conv1)
        x_conv2 = self.conv2(x_conv1)
        x_occ2 = self.occ_conv2(occ_in)
        x_conv2.features = torch.cat((x_conv2.features, x_occ2.features), dim=1)             # sparse_cat (:869-873)
        x_conv2 = self.conv2_combine(x_conv2)
        x_conv3 = self.conv3_combine(self.conv3(x_conv2))
        x_conv4 = self.conv4_combine(self.conv4(x_conv3))
        out = self.conv_out(x_conv4)
        # res_combine (:905-918), big_bev_combine
        d2 = self.down2(x_conv2)
        d3 = self.down3(x_conv3)
        x_conv4.features = torch.cat((d2.features, d3.features, x_conv4.features), dim=1)
        bev = self.squeezeBev(out).dense()
        n, c, dd, h, w = bev.shape
        bev = bev.view(n, c * dd, h, w)
        inds = x_conv4.indices.long()
        x_conv4.features = torch.cat((x_conv4.features, bev[inds[:, 0], :, inds[:, 2], inds[:, 3]]), dim=1)
        x_combine = self.down_combine(x_conv4)
        batch_dict.update({"encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8,
                           "multi_scale_3d_features": {"x_conv1": None, "x_conv2": None, "x_conv3": None, "x_conv4": None,
                                                       "x_combine": x_combine}})
        return batch_dict


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
