"""The C-ABI library loads and exports every symbol include/btcdet_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "btcdet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(btc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib_path():
    from btcdet_b200 import build
    return build.build()


def test_header_declares_something():
    names = _declared()
    assert "btc_voxelize" in names and "btc_sparse_conv_fwd" in names and len(names) >= 20


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(lib_path):
    from btcdet_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.btc_compiled_sm() == 100
    assert lib.btc_abi_version() >= 1


def test_host_only_queries(lib_path):
    from btcdet_b200 import _lib
    lib = _lib.load()
    # KITTI det grid [41,1600,1408], one scene: 92.3 M cells / 32 per entry
    assert lib.btc_index_entries(1, _lib.int3([41, 1600, 1408])) == (41 * 1600 * 1408 + 31) // 32
    assert lib.btc_voxelize_workspace_bytes(20000, 1, 16000, 5) > 0
    assert lib.btc_voxelize_workspace_bytes(-1, 1, 16000, 5) < 0


def test_bad_arguments_return_status_not_crash(lib_path):
    from btcdet_b200 import _lib
    lib = _lib.load()
    rc = lib.btc_sparse_conv_fwd(None, None, None, None, None, None, 0, None, 0, None, 27, 16, 16, 0, None)
    assert rc == -1 and b"null" in lib.btc_last_error()


def test_product_does_not_import_oracle():
    """The shipped packages must never touch oracle/ (tests, smoke and bench baselines only)."""
    for pkg in ("btcdet_b200", "spconv"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|orc_[a-z_]+\(", src, flags=re.M), \
                        os.path.join(dirpath, f)
