"""spconv convolution layers (spconv 1.2.1 `spconv/conv.py`) on the B200 gather-GEMM kernels.

Classes the reference instantiates (btcdet/models/backbones_3d/spconv_backbone.py:12-29,45-48,
58-66; occ_head_3D.py:25-31; conv_head.py:122): SubMConv3d, SubMConv2d, SparseConv3d,
SparseConvTranspose3d, SparseInverseConv3d (+2-D twins).  `forward` mirrors
SparseConvolution.forward of spconv 1.2.1 step by step: look the rulebook up in
`input.indice_dict[indice_key]`, else build it, then one fused gather-GEMM launch instead of
K x {gather, SGEMM, scatter-add}.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import init
from torch.nn.parameter import Parameter

from btcdet_b200 import ops as _ops

from .modules import SparseModule
from .tensor import SparseConvTensor


def _ntuple(v, ndim):
    if isinstance(v, (list, tuple)):
        assert len(v) == ndim, (v, ndim)
        return [int(x) for x in v]
    return [int(v)] * ndim


def _pad3(v, ndim, fill):
    """Lift 2-D geometry to 3-D with a singleton leading (z) axis."""
    return list(v) if ndim == 3 else [fill] + list(v)


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, output_padding=0, transposed=False, inverse=False, indice_key=None,
                 fused_bn=False, use_hash=False, algo=None):
        super(SparseConvolution, self).__init__()
        assert groups == 1
        assert ndim in (2, 3), "only 2-D and 3-D sparse convolutions are supported"
        kernel_size = _ntuple(kernel_size, ndim)
        stride = _ntuple(stride, ndim)
        padding = _ntuple(padding, ndim)
        dilation = _ntuple(dilation, ndim)
        output_padding = _ntuple(output_padding, ndim)
        for d, s in zip(dilation, stride):
            assert any([s == 1, d == 1]), "don't support this."
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.conv1x1 = np.prod(kernel_size) == 1
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.transposed = transposed
        self.inverse = inverse
        self.output_padding = output_padding
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.fused_bn = fused_bn
        self.use_hash = use_hash
        self.algo = 0 if algo is None else int(algo)
        # spconv 1.2.1 layout: [*kernel, Cin, Cout]; released BtcDet checkpoints load unchanged
        self.weight = Parameter(torch.Tensor(*kernel_size, in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return ("{in_channels}, {out_channels}, kernel_size={kernel_size}, stride={stride}, padding={padding}, "
                "subm={subm}, transposed={transposed}, inverse={inverse}, indice_key={indice_key}").format(
                    **self.__dict__)

    def forward(self, input, epilogue=None):
        """epilogue (extension): (scale [Cout], shift [Cout], relu) applied inside the kernel after the bias — what
        SparseSequential folds an eval-mode BatchNorm1d + ReLU into in static mode."""
        assert isinstance(input, SparseConvTensor)
        features = input.features
        indices = input.indices
        spatial_shape = input.spatial_shape
        batch_size = int(input.batch_size)
        nd = self.ndim
        if not features.is_cuda:
            raise RuntimeError("spconv (btcdet_b200) layers run on CUDA tensors only — there is no CPU path")

        if self.conv1x1 and all(s == 1 for s in self.stride) and all(p == 0 for p in self.padding) \
                and not self.transposed and not self.inverse:
            out_features = torch.mm(features, self.weight.view(self.in_channels, self.out_channels))
            if self.bias is not None:
                out_features = out_features + self.bias
            if epilogue is not None:
                out_features = out_features * epilogue[0] + epilogue[1]
                if epilogue[2]:
                    out_features = torch.relu(out_features)
            out_tensor = SparseConvTensor(out_features, indices, spatial_shape, batch_size, n_dev=input.n_dev)
            out_tensor.indice_dict = input.indice_dict
            out_tensor.grid = input.grid
            out_tensor._index = input._index
            return out_tensor

        datas = input.find_indice_pair(self.indice_key)
        if self.inverse:
            assert datas is not None and self.indice_key is not None
            rb_fwd, out_indices, out_spatial_shape = datas["rulebook"], datas["in_indices"], datas["in_spatial_shape"]
            assert rb_fwd.K == int(np.prod(self.kernel_size)), \
                "inverse conv must have same kernel size as its couple conv"
            rulebook = datas.get("inverse")
            if rulebook is None:
                rulebook = rb_fwd.inverse()
                datas["inverse"] = rulebook
            out_index = datas.get("in_index")
        else:
            if self.indice_key is not None and datas is not None:
                rulebook, out_indices, out_spatial_shape = datas["rulebook"], datas["out_indices"], \
                    datas["out_spatial_shape"]
                out_index = rulebook.out_index
            else:
                coords4 = input._coords4()
                shape3 = input._shape3()
                k3 = _pad3(self.kernel_size, nd, 1)
                d3 = _pad3(self.dilation, nd, 1)
                if self.subm:
                    if input._index is None:
                        input._index = _ops.build_hash(coords4, batch_size, shape3, n_dev=input.n_dev)
                    rulebook = _ops.rulebook_subm(coords4, batch_size, shape3, k3, d3, index=input._index, n_dev=input.n_dev)
                    out_indices, out_spatial_shape = indices, spatial_shape
                    out_index = input._index
                else:
                    rulebook = _ops.rulebook_conv(coords4, batch_size, shape3, k3, _pad3(self.stride, nd, 1),
                                                  _pad3(self.padding, nd, 0), d3, transposed=self.transposed,
                                                  output_padding=_pad3(self.output_padding, nd, 0), n_dev=input.n_dev)
                    out_indices = rulebook.out_coords if nd == 3 else rulebook.out_coords[:, [0, 2, 3]].contiguous()
                    out_spatial_shape = rulebook.out_shape if nd == 3 else rulebook.out_shape[1:]
                    out_index = rulebook.out_index
                input.indice_dict[self.indice_key] = {
                    "rulebook": rulebook,
                    "out_indices": out_indices,
                    "out_spatial_shape": out_spatial_shape,
                    "in_indices": indices,
                    "in_spatial_shape": spatial_shape,
                    "in_index": input._index,
                }

        out_features = _ops.SparseConvFunction.apply(features, self.weight, self.bias, rulebook, self.algo, epilogue)
        out_tensor = SparseConvTensor(out_features, out_indices, out_spatial_shape, batch_size, n_dev=rulebook.n_out_dev)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        out_tensor._index = out_index
        return out_tensor


class SparseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SparseConv2d, self).__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                           bias, indice_key=indice_key, use_hash=use_hash, algo=algo)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SparseConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                           bias, indice_key=indice_key, use_hash=use_hash, algo=algo)


class SparseConvTranspose2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SparseConvTranspose2d, self).__init__(2, in_channels, out_channels, kernel_size, stride, padding,
                                                    dilation, groups, bias, transposed=True, indice_key=indice_key,
                                                    use_hash=use_hash, algo=algo)


class SparseConvTranspose3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SparseConvTranspose3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding,
                                                    dilation, groups, bias, transposed=True, indice_key=indice_key,
                                                    use_hash=use_hash, algo=algo)


class SparseInverseConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None):
        super(SparseInverseConv2d, self).__init__(2, in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                                                  indice_key=indice_key, algo=algo)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None):
        super(SparseInverseConv3d, self).__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                                                  indice_key=indice_key, algo=algo)


class SubMConv2d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SubMConv2d, self).__init__(2, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                         bias, True, indice_key=indice_key, use_hash=use_hash, algo=algo)


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None, use_hash=False, algo=None):
        super(SubMConv3d, self).__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups,
                                         bias, True, indice_key=indice_key, use_hash=use_hash, algo=algo)
