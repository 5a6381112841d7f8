"""Test-only mirrors of the reference's two sparse backbones, written against the `spconv` surface.

The GPU box has no reference checkout, so the `-m gpu` tests cannot import
btcdet/models/backbones_3d/spconv_backbone.py.  These classes restate the layer tables of
VoxelBackBoneDeconv (:91-203, + the OccHead3D convs, occ_head_3D.py:25-31) and VoxelBackBone8xOcc (:630-1019, in the
shipped configuration OCC_CONV_TYPE ['identity','maxpool'], OCC_CONV_EXECUTE [False, True], OUT_FEAT_TYPE
[... 'big_bev_combine']) with the same attribute names, so that tests/test_reference_import.py can assert (here, where
the reference exists) that state-dict keys and shapes are identical to the real classes.  `forward` is written once
over a tiny backend (`apply`, `cat`, `new_tensor`, `dense_gather`) so the same dataflow runs on the CUDA shim and on
the CPU oracle.
"""
from functools import partial

import numpy as np
import torch
import torch.nn as nn

import spconv


def _block(cin, cout, k, key, stride=1, padding=0, kind="subm"):
    norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
    if kind == "subm":
        conv = spconv.SubMConv3d(cin, cout, k, bias=False, indice_key=key)
    elif kind == "spconv":
        conv = spconv.SparseConv3d(cin, cout, k, stride=stride, padding=padding, bias=False, indice_key=key)
    elif kind == "spdeconv":
        conv = spconv.SparseConvTranspose3d(cin, cout, k, stride=stride, padding=padding, bias=False, indice_key=key)
    elif kind == "maxpool":
        return spconv.SparseSequential(spconv.SparseMaxPool3d(k, stride=stride, padding=padding))
    return spconv.SparseSequential(conv, norm(cout), nn.ReLU())


class ShimBackend:
    """Runs modules on the CUDA spconv shim."""

    def apply(self, module, x):
        return module(x)

    def cat(self, a, b_list):
        a.features = torch.cat([a.features] + [b.features if hasattr(b, "features") else b for b in b_list], dim=1)
        return a

    def new_tensor(self, feats, like, shape):
        return spconv.SparseConvTensor(features=feats, indices=like.indices, spatial_shape=shape, batch_size=like.batch_size)

    def bev_gather(self, bev, at):
        d = bev.dense()
        n, c, dd, h, w = d.shape
        d = d.view(n, c * dd, h, w)
        i = at.indices.long()
        return d[i[:, 0], :, i[:, 2], i[:, 3]]


class OracleBackend:
    """Runs the same modules on oracle tensors (CPU)."""

    def apply(self, module, x):
        from tests import oracle_net
        return oracle_net.run(module, x)

    def cat(self, a, b_list):
        a.features = np.concatenate([a.features] + [b.features if hasattr(b, "features") else b for b in b_list], axis=1)
        return a

    def new_tensor(self, feats, like, shape):
        from oracle import oracle as O
        return O.SparseTensor(np.asarray(feats, np.float32), like.indices, shape, like.batch_size)

    def bev_gather(self, bev, at):
        from oracle import oracle as O
        d = O.dense(bev.features, bev.indices, bev.spatial_shape, bev.batch_size)
        n, c, dd, h, w = d.shape
        d = d.reshape(n, c * dd, h, w)
        i = at.indices.astype(np.int64)
        return d[i[:, 0], :, i[:, 2], i[:, 3]]


class OccBackboneMirror(nn.Module):
    """VoxelBackBoneDeconv on the cylindrical occ grid + the two OccHead3D convolutions."""

    def __init__(self, input_channels=4, grid_size=(209, 157, 9)):
        super().__init__()
        self.sparse_shape = list(grid_size[::-1])
        c = [16, 32, 64]
        self.conv1 = spconv.SparseSequential(_block(input_channels, c[0], 3, "spconv1", padding=1, kind="spconv"))
        self.conv2 = spconv.SparseSequential(_block(c[0], c[1], 3, "spconv2", stride=2, padding=1, kind="spconv"),
                                             _block(c[1], c[1], 3, "subm2", padding=1))
        self.conv3 = spconv.SparseSequential(_block(c[1], c[2], 3, "spconv3", stride=2, padding=1, kind="spconv"),
                                             _block(c[2], c[2], 3, "subm3", padding=1))
        self.deconv4 = spconv.SparseSequential(_block(c[2], c[1], 3, "spconv4", stride=2, padding=1, kind="spdeconv"),
                                               _block(c[1], c[1], 3, "subm4", padding=1))
        self.deconv5 = spconv.SparseSequential(_block(c[1], c[1], 3, "spconv5", stride=2, padding=1, kind="spdeconv"),
                                               _block(c[1], c[1], 3, "subm5", padding=1))
        # OccHead3D (softmax, class-agnostic, REG): cls 32->2 with bias, residual 32->3 without
        self.conv_cls = spconv.SparseSequential(spconv.SubMConv3d(c[1], 2, 3, padding=1, bias=True, indice_key="cls_ind"))
        self.conv_res = spconv.SparseSequential(spconv.SubMConv3d(c[1], 3, 3, padding=1, bias=False, indice_key="res_ind"))

    def run(self, be, x):
        for name in ("conv1", "conv2", "conv3", "deconv4", "deconv5"):
            x = be.apply(getattr(self, name), x)
        return {"encoded": x, "cls": be.apply(self.conv_cls, x), "res": be.apply(self.conv_res, x)}


class DetBackboneMirror(nn.Module):
    """VoxelBackBone8xOcc in the shipped configuration (max-pooled occupancy side channel, big_bev_combine)."""

    def __init__(self, input_channels=6, raw_channels=4, grid_size=(1408, 1600, 40)):
        super().__init__()
        self.sparse_shape = [grid_size[2] + 1, grid_size[1], grid_size[0]]
        ch = [16, 32, 64, 64, 128]
        occ = input_channels - raw_channels
        self.occ_code_num = occ
        norm = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.occ_conv2 = spconv.SparseSequential(_block(occ, occ, 3, "spconv2", stride=2, padding=1, kind="maxpool"))
        self.conv1 = spconv.SparseSequential(spconv.SubMConv3d(input_channels, ch[0], 3, padding=1, bias=False,
                                                               indice_key="subm1"), norm(ch[0]), nn.ReLU())
        self.conv1_combine = spconv.SparseSequential(_block(ch[0], ch[0], 3, "subm1", padding=1))
        self.conv2 = spconv.SparseSequential(_block(ch[0], ch[1], 3, "spconv2", stride=2, padding=1, kind="spconv"))
        self.conv2_combine = spconv.SparseSequential(_block(ch[1] + occ, ch[1], 3, "subm2", padding=1),
                                                     _block(ch[1], ch[1], 3, "subm2", padding=1))
        self.conv3 = spconv.SparseSequential(_block(ch[1], ch[2], 3, "spconv3", stride=2, padding=1, kind="spconv"))
        self.conv3_combine = spconv.SparseSequential(_block(ch[2], ch[2], 3, "subm3", padding=1),
                                                     _block(ch[2], ch[2], 3, "subm3", padding=1))
        self.conv4 = spconv.SparseSequential(_block(ch[2], ch[3], 3, "spconv4", stride=2, padding=(0, 1, 1), kind="spconv"))
        self.conv4_combine = spconv.SparseSequential(_block(ch[3], ch[3], 3, "subm4", padding=1),
                                                     _block(ch[3], ch[3], 3, "subm4", padding=1))
        self.conv_out = spconv.SparseSequential(spconv.SparseConv3d(ch[3], ch[4], (3, 1, 1), stride=(2, 1, 1), padding=0,
                                                                    bias=False, indice_key="spconv_down2"),
                                                norm(ch[4]), nn.ReLU())
        self.down2 = spconv.SparseSequential(_block(ch[1], ch[1], 3, "spconv3", stride=2, padding=1, kind="spconv"),
                                             _block(ch[1], ch[2], 3, "spconv4", stride=2, padding=(0, 1, 1), kind="spconv"))
        self.down3 = spconv.SparseSequential(_block(ch[2], ch[2], 3, "spconv4", stride=2, padding=(0, 1, 1), kind="spconv"))
        self.squeezeBev = spconv.SparseSequential(_block(ch[4], ch[3], (2, 1, 1), "subm_down2", stride=(2, 1, 1), padding=0,
                                                         kind="spconv"))
        self.down_combine = spconv.SparseSequential(_block(ch[2] * 2 + ch[3] * 2, ch[3] * 2, 3, "subm4", padding=1),
                                                    _block(ch[3] * 2, ch[3] * 2, 3, "subm4", padding=1))

    def run(self, be, x, occ_feats):
        x = be.apply(self.conv1, x)
        occ_in = be.new_tensor(occ_feats, x, self.sparse_shape)        # fresh tensor: empty rulebook cache
        x1 = be.apply(self.conv1_combine, x)
        x2 = be.apply(self.conv2, x1)
        occ2 = be.apply(self.occ_conv2, occ_in)
        x2 = be.cat(x2, [occ2])                                          # relies on deterministic output ordering
        x2 = be.apply(self.conv2_combine, x2)
        x3 = be.apply(self.conv3_combine, be.apply(self.conv3, x2))
        x4 = be.apply(self.conv4_combine, be.apply(self.conv4, x3))
        out = be.apply(self.conv_out, x4)
        d2 = be.apply(self.down2, x2)                                    # cached 'spconv3' / 'spconv4' rulebooks
        d3 = be.apply(self.down3, x3)
        x4 = be.cat(x4, [d2, d3])
        x4.features = x4.features[:, list(range(64, 192)) + list(range(0, 64))]   # reference order: (d2, d3, x4)
        bev = be.apply(self.squeezeBev, out)
        x4 = be.cat(x4, [be.bev_gather(bev, x4)])
        x_combine = be.apply(self.down_combine, x4)
        return {"x_conv2": x2, "x_conv3": x3, "out": out, "x_combine": x_combine}
