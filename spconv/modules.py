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
        mods = list(self._modules.items())
        skip = 0
        for i, (k, module) in enumerate(mods):
            if skip:
                skip -= 1
                continue
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                self._sparity_dict[k] = input.sparity
                # static mode (SparseConvTensor.n_dev): an eval-mode BatchNorm1d (+ ReLU) behind a convolution is folded into
                # the convolution's epilogue (per-channel affine + max) — no elementwise passes over capacity-sized rows
                fuse = _fusable_epilogue(module, mods, i) if getattr(input, "n_dev", None) is not None else None
                if fuse is not None:
                    scale, shift, relu, skip = fuse
                    input = module(input, epilogue=(scale, shift, relu))
                else:
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


def _fusable_epilogue(conv, mods, i):
    """(scale, shift, relu, modules consumed) when mods[i + 1] is an eval-mode BatchNorm1d (optionally followed by ReLU)."""
    from .conv import SparseConvolution
    if not isinstance(conv, SparseConvolution) or i + 1 >= len(mods):
        return None
    bn = mods[i + 1][1]
    if not isinstance(bn, nn.BatchNorm1d) or bn.training or not bn.track_running_stats or bn.running_var is None:
        return None
    key = (bn.running_var._version, bn.running_mean._version,
           None if bn.weight is None else bn.weight._version, None if bn.bias is None else bn.bias._version)
    cached = getattr(bn, "_btc_fold", None)
    if cached is None or cached[0] != key:
        with torch.no_grad():
            w = bn.weight if bn.weight is not None else torch.ones_like(bn.running_var)
            b = bn.bias if bn.bias is not None else torch.zeros_like(bn.running_var)
            scale = (w / torch.sqrt(bn.running_var + bn.eps)).float().contiguous()
            shift = (b - bn.running_mean * scale).float().contiguous()
        cached = (key, scale, shift)
        bn._btc_fold = cached
    relu = i + 2 < len(mods) and isinstance(mods[i + 2][1], nn.ReLU)
    return cached[1], cached[2], relu, (2 if relu else 1)


class ToDense(SparseModule):
    def forward(self, x: SparseConvTensor):
        return x.dense()


class RemoveGrid(SparseModule):
    def forward(self, x: SparseConvTensor):
        x.grid = None
        return x
