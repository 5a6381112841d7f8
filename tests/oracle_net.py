"""Runs spconv-style module trees on the CPU oracle (test helper; mirrors SparseSequential.forward and
SparseConvolution.forward of spconv 1.2.1 on oracle/ ops)."""
import numpy as np
import torch.nn as nn

import spconv
from oracle import oracle as O


def to_oracle_tensor(features, indices, spatial_shape, batch_size):
    return O.SparseTensor(np.asarray(features, np.float32), np.asarray(indices, np.int32), spatial_shape, batch_size)


def run(module, x):
    """Apply `module` (SparseSequential / SparseConvolution / SparseMaxPool / BN / ReLU) to oracle tensor x."""
    if isinstance(module, spconv.SparseSequential):
        for m in module.children():
            x = run(m, x)
        return x
    if isinstance(module, spconv.SparseConvolution):
        assert module.ndim == 3 and not module.inverse
        w = module.weight.detach().cpu().numpy().reshape(-1, module.in_channels, module.out_channels)
        b = None if module.bias is None else module.bias.detach().cpu().numpy()
        return O.sparse_conv(x, w, module.kernel_size, module.stride, module.padding, module.dilation,
                             subm=module.subm, transpose=module.transposed, indice_key=module.indice_key, bias=b,
                             out_padding=module.output_padding)
    if isinstance(module, spconv.SparseMaxPool):
        outids, pairs, pair_num, oshape = O.get_indice_pairs(x.indices, x.batch_size, x.spatial_shape,
                                                             module.kernel_size, module.stride, module.padding,
                                                             module.dilation, 0, module.subm, False)
        y = O.SparseTensor(O.indice_maxpool(x.features, pairs, pair_num, outids.shape[0]), outids, oshape, x.batch_size)
        y.indice_dict = x.indice_dict
        return y
    if isinstance(module, nn.BatchNorm1d):
        assert not module.training, "oracle BN is eval-mode (running statistics)"
        rm, rv = module.running_mean.cpu().numpy(), module.running_var.cpu().numpy()
        g, b = module.weight.detach().cpu().numpy(), module.bias.detach().cpu().numpy()
        x.features = ((x.features - rm) / np.sqrt(rv + module.eps) * g + b).astype(np.float32)
        return x
    if isinstance(module, nn.ReLU):
        x.features = np.maximum(x.features, 0)
        return x
    raise NotImplementedError(type(module))
