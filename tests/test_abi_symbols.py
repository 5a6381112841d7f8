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
    rc = lib.btc_sparse_conv_fwd(None, None, None, None, None, None, 0, None, 5, None, 27, 16, 16, 0, None)
    assert rc == -1 and b"null" in lib.btc_last_error()
    # an EMPTY output is not an error: torch hands out null data pointers for 0-row tensors (spconv handles empty inputs)
    assert lib.btc_sparse_conv_fwd(None, None, None, None, None, None, 0, None, 0, None, 27, 16, 16, 0, None) == 0
    assert lib.btc_maxpool_fwd(None, None, None, 0, None, 27, 2, None) == 0
    assert lib.btc_sparse_conv_fwd_tc(None, None, None, None, None, None, 0, None, 0, None, 27, 32, 32, None) == 0


def test_product_does_not_import_oracle():
    """The shipped packages must never touch oracle/ (tests, smoke and bench baselines only)."""
    for pkg in ("btcdet_b200", "spconv"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|orc_[a-z_]+\(", src, flags=re.M), \
                        os.path.join(dirpath, f)


def test_tensor_core_tile_host_queries(lib_path):
    """Host-side contract of the tcgen05 tile: supported shapes (incl. the shared-memory bound on K), packed sizes, knobs."""
    from btcdet_b200 import _lib
    lib = _lib.load()
    sup = lib.btc_sparse_conv_tc_supported
    # every layer shape of VoxelBackBone8x / VoxelBackBone8xOcc / the occupancy backbone that is a multiple of 4 channels
    for K, cin, cout in [(27, 4, 16), (27, 16, 16), (27, 16, 32), (27, 32, 32), (27, 32, 64), (27, 64, 64), (3, 64, 128),
                         (2, 128, 64), (27, 256, 128), (27, 128, 128), (27, 32, 2 * 2)]:
        assert sup(K, cin, cout) == 1, (K, cin, cout)
    assert sup(27, 6, 16) == 0 and sup(27, 34, 32) == 0          # c_in % 4 != 0 -> FFMA tile
    assert sup(27, 32, 6) == 0 and sup(27, 32, 256) == 0         # c_out % 4 != 0 (>= 4: misaligned vector stores) / > 128
    assert sup(27, 32, 2) == 1 and sup(27, 32, 3) == 1           # the occupancy head's 32 -> 2 / 32 -> 3 (scalar stores)
    assert sup(1, 4, 16) == 0                                    # reduction shorter than one 32-element stage
    assert sup(33, 32, 64) == 1 and sup(34, 32, 64) == 0         # two [128 x K] index tiles must fit next to the rings
    assert sup(49, 32, 32) == 1 and sup(50, 32, 32) == 0         # (the N = 32 weight ring is 16 KB smaller)
    assert sup(64, 32, 32) == 0 and sup(125, 16, 16) == 0
    # packed image: per 32-element chunk a hi and a lo tile of N x 128 bytes, N = c_out padded to 32 / 64 / 128
    assert lib.btc_sparse_conv_tc_packed_bytes(27, 64, 64) == (27 * 64 // 32) * 2 * 64 * 128
    assert lib.btc_sparse_conv_tc_packed_bytes(27, 4, 16) == 4 * 2 * 32 * 128          # ceil(108 / 32) = 4 chunks
    assert lib.btc_sparse_conv_tc_packed_bytes(3, 64, 128) == 6 * 2 * 128 * 128
    assert lib.btc_sparse_conv_tc_packed_bytes(27, 6, 16) < 0
    # knobs validate their arguments and keep the defaults otherwise
    assert lib.btc_sparse_conv_tc_config(-1, -1, -1) == 0
    assert lib.btc_sparse_conv_tc_config(12, -1, -1) == -1 and b"producer_warps" in lib.btc_last_error()
    assert lib.btc_sparse_conv_tc_config(16, 0, 1) == 0
    assert lib.btc_sparse_conv_tc_grid(0) == -1 and lib.btc_sparse_conv_tc_grid(149) == -1
    assert lib.btc_sparse_conv_tc_grid(148) == 0
    assert lib.btc_sparse_conv_tc_diag(0) == 0


def test_workspace_and_capacity_queries(lib_path):
    from btcdet_b200 import _lib
    lib = _lib.load()
    n = lib.btc_hash_slots(20000)
    assert n >= 2 * 20000 and n & (n - 1) == 0                   # open addressing at load <= 0.5, power of two
    assert lib.btc_hash_slots(0) >= 1
    e = lib.btc_index_entries(2, _lib.int3([21, 800, 704]))
    assert e == (2 * 21 * 800 * 704 + 31) // 32
    assert lib.btc_index_workspace_bytes(e) > 0
    assert lib.btc_rulebook_pairs_workspace_bytes(100000, 27) >= ((100000 + 2047) // 2048) * 27 * 4
    assert lib.btc_sparse_conv_bwd_workspace_bytes(27, 64, 64) >= 27 * 64 * 64 * 4
    assert lib.btc_revoxelize_workspace_bytes(40000, e) > 0
    assert lib.btc_occ_select_workspace_bytes(2, _lib.int3([209, 157, 9])) > 0
    assert lib.btc_occ_box_targets_workspace_bytes(2, 12, 40000, 1000) > 0


def test_survey_named_aliases_forward(lib_path):
    """btc_voxelize_cuda / btc_rulebook_pool / btc_occ_inject_revoxelize are thin forwards: same argument checking."""
    from btcdet_b200 import _lib
    lib = _lib.load()
    three = _lib.int3([1, 1, 1])
    assert lib.btc_rulebook_pool(None, 0, None, 1, None, None, None, None, None, None, None, 0, None, 0, None, None, None, None, 0,
                                 None) == lib.btc_rulebook_conv(None, 0, None, 1, None, None, None, None, None, None, 0, None, 0,
                                                                None, 0, None, None, None, None, 0, None) == -1
    assert lib.btc_voxelize_cuda(None, 10, 4, None, 1, None, None, three, 5, 100, None, None, None, None, None, None, 0, None) == \
        lib.btc_voxelize(None, 10, 4, None, 1, None, None, three, 5, 100, None, None, None, None, None, None, 0, None)
    assert lib.btc_occ_inject_revoxelize(None, 10, None, 1, None, None, 0, None, 0, None, None, None, None, None, None, 0, None) == \
        lib.btc_revoxelize(None, 10, None, 1, None, None, 0, None, 0, None, None, None, None, None, None, 0, None)
