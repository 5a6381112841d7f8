"""Host logic of the static (capacity) mode that needs no GPU: capacity policy and the check registry."""
import pytest
import torch

from btcdet_b200 import _lib, ops


def test_static_capacity_policy():
    cells = 2 * 21 * 800 * 704
    # strided level on a large grid: STATIC_GROWTH x the input capacity (never above the true bound or the grid)
    assert ops._static_cap(80000, [3, 3, 3], [2, 2, 2], False, cells) == int(ops.STATIC_GROWTH * 80000)
    assert ops._static_cap(10, [3, 3, 3], [2, 2, 2], False, cells) == 80            # true bound 10 * 2^3 below the floor of 1024
    assert ops._static_cap(80000, [3, 1, 1], [2, 1, 1], False, cells) == 80000
    # stride-1 (dilating) and transposed layers: the true bound k^3 * n, capped by the grid
    small = 2 * 9 * 157 * 209
    assert ops._static_cap(40000, [3, 3, 3], [1, 1, 1], False, small) == small
    assert ops._static_cap(100, [3, 3, 3], [2, 2, 2], True, small) == 2700
    assert ops._static_cap(0, [3, 3, 3], [2, 2, 2], False, cells) == 1


def test_static_checks_registry():
    assert ops._static_checks is None
    ops._register_cap(torch.tensor([5], dtype=torch.int32), 3, "outside a context: ignored")
    with ops.static_checks() as chk:
        ops._register_cap(torch.tensor([5], dtype=torch.int32), 8, "fits")
        with ops.static_checks() as inner:          # contexts nest; the inner one collects its own
            ops._register_cap(torch.tensor([9], dtype=torch.int32), 8, "inner overflow")
        ops._register_cap(torch.tensor([8], dtype=torch.int32), 8, "exactly full")
    assert ops._static_checks is None
    assert chk.verify() == [5, 8]
    with pytest.raises(_lib.BtcError, match="inner overflow has 9 rows > capacity 8"):
        inner.verify()
    assert ops.StaticChecks().verify() == []
