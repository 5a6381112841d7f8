"""spconv.SparseModule / SparseSequential (spconv 1.2.1 `spconv/modules.py`); the reference mixes
sparse layers with nn.BatchNorm1d / nn.ReLU inside SparseSequential
(btcdet/models/backbones_3d/spconv_backbone.py:33-38)."""
import sys
from collections import OrderedDict

import torch
from torch import nn

from .tensor import SparseConvTensor


class SparseModule(nn.Module):
    """Marker base class: modules deriving from it receive the SparseConvTensor itself."""
    pass


def is_spconv_module(module):
    return isinstance(module, SparseModule)


def is_sparse_conv(module):
    from .conv import SparseConvolution
    return isinstance(module, SparseConvolution)


class SparseSequential(SparseModule):
    """Sequential container; non-sparse modules are applied to `.features` in place on the same
    SparseConvTensor object."""

    def __init__(self, *args, **kwargs):
        super(SparseSequential, self).__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if sys.version_info < (3, 6):
                raise ValueError("kwargs only supported in py36+")
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)
        self._sparity_dict = {}

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError('index {} is out of range'.format(idx))
        if idx < 0:
            idx += len(self)
        it = iter(self._modules.values())
        for i in range(idx):
            next(it)
        return next(it)

    def __len__(self):
        return len(self._modules)

    @property
    def sparity_dict(self):
        return self._sparity_dict

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        for k, module in self._modules.items():
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                self._sparity_dict[k] = input.sparity
                input = module(input)
            else:
                if isinstance(input, SparseConvTensor):
                    if input.indices.shape[0] != 0:
                        input.features = module(input.features)
                else:
                    input = module(input)
        return input

    def fused(self):
        """spconv's conv+BN fusion helper; not used by the reference."""
        raise NotImplementedError("SparseSequential.fused() is not part of the BtcDet hot path")


class ToDense(SparseModule):
    def forward(self, x: SparseConvTensor):
        return x.dense()


class RemoveGrid(SparseModule):
    def forward(self, x: SparseConvTensor):
        x.grid = None
        return x
