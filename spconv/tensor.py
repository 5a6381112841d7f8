"""spconv.SparseConvTensor (spconv 1.2.1 `spconv/__init__.py`), reference ctor call sites
btcdet/models/backbones_3d/spconv_backbone.py:155-160,950-964."""
import numpy as np
import torch

from btcdet_b200 import ops as _ops


class SparseConvTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, n_dev=None):
        """
        features: [N, C] float32; indices: [N, ndim+1] int32 rows (batch, *spatial) — 3-D (b,z,y,x)
        or 2-D (b,y,x); spatial_shape: list/array of ndim ints; grid: kept for API compatibility
        (spconv's dense lookup grid; the B200 path uses a rank bitmap kept in `_index`).
        n_dev (extension, not in spconv): int32 [1] CUDA tensor = number of live rows; features / indices are then
        capacity-sized ("static mode"): every layer keeps its counts on the device, allocates capacity-sized outputs
        and never reads the host, so a whole forward can be captured in one CUDA graph (inference; train-mode
        BatchNorm would average over the dead rows).
        """
        self.features = features
        self.indices = indices
        self.spatial_shape = spatial_shape
        self.batch_size = batch_size
        self.indice_dict = {}
        self.grid = grid
        self._index = None  # btcdet_b200.ops.CoordIndex of `indices`, filled lazily
        self.n_dev = n_dev

    @property
    def spatial_size(self):
        return np.prod(self.spatial_shape)

    def find_indice_pair(self, key):
        if key is None:
            return None
        if key in self.indice_dict:
            return self.indice_dict[key]
        return None

    # -- helpers shared by the layers -------------------------------------------------------
    def _shape3(self):
        s = [int(v) for v in self.spatial_shape]
        return s if len(s) == 3 else [1] + s

    def _coords4(self):
        """int32 contiguous (b, z, y, x) rows; 2-D tensors get a singleton z."""
        ind = self.indices
        if ind.dtype != torch.int32:
            ind = ind.int()
        if ind.shape[1] == 3:
            ind = torch.cat([ind[:, :1], torch.zeros_like(ind[:, :1]), ind[:, 1:]], dim=1)
        return ind.contiguous()

    def dense(self, channels_first=True):
        """zeros [B, *spatial, C], index-assign rows, permute to channels first."""
        shape3 = self._shape3()
        out = _ops.ToDenseFunction.apply(self.features, self._coords4(), int(self.batch_size), shape3, self.n_dev)
        if len(self.spatial_shape) == 2:
            out = out.squeeze(2)
        if not channels_first:
            ndim = len(self.spatial_shape)
            out = out.permute(0, *range(2, ndim + 2), 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / np.prod(self.spatial_shape) / self.batch_size
