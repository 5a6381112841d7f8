"""Drop-in `spconv` package (spconv v1.2.1 surface) backed by the B200 CUDA library.

The reference imports this name at module level — btcdet/models/backbones_3d/spconv_backbone.py:3,
btcdet/models/occ_pnt/occ_dense_heads/occ_head_3D.py:6,
btcdet/models/occ_pnt/occ_training_targets/occ_targets_template.py:8,
btcdet/models/roi_heads/conv_head.py:9 — and uses exactly the names exported here
(SURVEY.md §8b).  Semantics follow spconv 1.2.1: weight layout
[*kernel, Cin, Cout], state-dict keys `weight`/`bias`, plain mutable `.features`/`.indices`,
a shared `indice_dict` rulebook cache, outputs of strided convs in ascending flat-key order,
`dense()` channels-first.  Every forward/backward runs in libbtcdet_b200.so; there is no CPU path.
"""
from . import utils  # noqa: F401
from .conv import (SparseConv2d, SparseConv3d, SparseConvTranspose2d, SparseConvTranspose3d,  # noqa: F401
                   SparseConvolution, SparseInverseConv2d, SparseInverseConv3d, SubMConv2d, SubMConv3d)
from .modules import SparseModule, SparseSequential, is_spconv_module  # noqa: F401
from .pool import SparseMaxPool, SparseMaxPool2d, SparseMaxPool3d  # noqa: F401
from .tensor import SparseConvTensor  # noqa: F401

__version__ = "1.2.1+btcdet_b200"
