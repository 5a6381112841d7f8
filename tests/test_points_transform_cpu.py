"""Pins oracle/coords.py against the reference's own numpy functions
(btcdet/utils/coords_utils.py:268-292); skipped where no reference checkout / staged copy exists."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


def test_oracle_equals_reference_transforms():
    from btcdet_b200 import synthetic as S
    from oracle import coords
    mods = ref_loader.load_reference_modules()
    cu = mods["coords_utils"]
    for seed, n in ((0, 20000), (1, 5000)):
        pts = S.lidar_like(n, seed=seed)
        for cols in (4, 3):
            p = np.ascontiguousarray(pts[:, :cols])
            np.testing.assert_array_equal(coords.absxyz_2_cylinxyz(p), cu.absxyz_2_cylinxyz_np(p))
            np.testing.assert_array_equal(coords.absxyz_2_spherexyz(p), cu.absxyz_2_spherexyz_np(p))
    # the two-rounding spelling the CUDA kernel follows: (atan2 * float32(180)) / float32(pi), sqrt(x*x + y*y)
    p = S.lidar_like(20000, seed=2)
    a = np.arctan2(-p[:, 1], p[:, 0])
    want = coords.absxyz_2_cylinxyz(p)
    np.testing.assert_array_equal(want[:, 1], (a * np.float32(180.)) / np.float32(np.pi))
    np.testing.assert_array_equal(want[:, 0], np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]))
    assert want.dtype == np.float32
