"""spconv.SparseMaxPool3d (spconv 1.2.1 `spconv/pool.py`); reference use:
btcdet/models/backbones_3d/spconv_backbone.py:29 (post_act_block 'maxpool'), :831-847 (occ_conv2).
Rulebook as a regular conv (recomputed each call, no indice_key), output zero-initialised then
out[o] = max(out[o], in[i]) — negative inputs clamp to 0."""
from btcdet_b200 import ops as _ops

from .conv import _ntuple, _pad3
from .modules import SparseModule
from .tensor import SparseConvTensor


class SparseMaxPool(SparseModule):
    def __init__(self, ndim, kernel_size, stride=1, padding=0, dilation=1, subm=False):
        super(SparseMaxPool, self).__init__()
        assert ndim in (2, 3)
        self.ndim = ndim
        self.kernel_size = _ntuple(kernel_size, ndim)
        self.stride = _ntuple(stride, ndim)
        self.padding = _ntuple(padding, ndim)
        self.subm = subm
        self.dilation = _ntuple(dilation, ndim)

    def forward(self, input):
        assert isinstance(input, SparseConvTensor)
        features = input.features
        if not features.is_cuda:
            raise RuntimeError("spconv (btcdet_b200) layers run on CUDA tensors only — there is no CPU path")
        nd = self.ndim
        batch_size = int(input.batch_size)
        coords4, shape3 = input._coords4(), input._shape3()
        k3, d3 = _pad3(self.kernel_size, nd, 1), _pad3(self.dilation, nd, 1)
        if self.subm:
            if input._index is None:
                input._index = _ops.build_hash(coords4, batch_size, shape3, n_dev=input.n_dev)
            rulebook = _ops.rulebook_subm(coords4, batch_size, shape3, k3, d3, index=input._index, n_dev=input.n_dev)
            out_indices, out_spatial_shape, out_index = input.indices, input.spatial_shape, input._index
        else:
            rulebook = _ops.rulebook_conv(coords4, batch_size, shape3, k3, _pad3(self.stride, nd, 1),
                                          _pad3(self.padding, nd, 0), d3, transposed=False, n_dev=input.n_dev)
            out_indices = rulebook.out_coords if nd == 3 else rulebook.out_coords[:, [0, 2, 3]].contiguous()
            out_spatial_shape = rulebook.out_shape if nd == 3 else rulebook.out_shape[1:]
            out_index = rulebook.out_index
        out_features = _ops.SparseMaxPoolFunction.apply(features, rulebook)
        out_tensor = SparseConvTensor(out_features, out_indices, out_spatial_shape, batch_size, n_dev=rulebook.n_out_dev)
        out_tensor.indice_dict = input.indice_dict
        out_tensor.grid = input.grid
        out_tensor._index = out_index
        return out_tensor


class SparseMaxPool2d(SparseMaxPool):
    def __init__(self, kernel_size, stride=1, padding=0, dilation=1):
        super(SparseMaxPool2d, self).__init__(2, kernel_size, stride, padding, dilation)


class SparseMaxPool3d(SparseMaxPool):
    def __init__(self, kernel_size, stride=1, padding=0, dilation=1):
        super(SparseMaxPool3d, self).__init__(3, kernel_size, stride, padding, dilation)
