"""ctypes binding of the C-ABI library `libbtcdet_b200.so` (include/btcdet_b200.h).

There is NO fallback: if the library is missing or a call fails, the caller gets an
exception.  The product path never routes through `oracle/` or CPU code.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbtcdet_b200.so")

_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64

# name -> (restype, argtypes); mirrors include/btcdet_b200.h one to one
SIGNATURES = {
    "btc_abi_version": (_i, []),
    "btc_compiled_sm": (_i, []),
    "btc_last_error": (ctypes.c_char_p, []),
    "btc_voxelize_workspace_bytes": (_i64, [_i64, _i, _i, _i]),
    "btc_voxelize": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_voxelize_group": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_voxelize_hash_view": (_i, [_i64, _i, _i, _i, _p, _p, _p]),
    "btc_voxelize_fill": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_points_to_cylinder": (_i, [_p, _i, _p, _i, _i, _p, _p]),
    "btc_index_entries": (_i64, [_i, _p]),
    "btc_index_workspace_bytes": (_i64, [_i64]),
    "btc_index_build": (_i, [_p, _i, _p, _i, _p, _p, _i64, _p, _p, _p, _i64, _p]),
    "btc_index_clear": (_i, [_p, _i, _p, _i, _p, _p, _i64, _p]),
    "btc_hash_slots": (_i64, [_i]),
    "btc_hash_build": (_i, [_p, _i, _p, _i, _p, _p, _p, _i64, _p]),
    "btc_rulebook_subm_hash": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _i64, _p, _p]),
    "btc_rulebook_subm": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "btc_rulebook_conv": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _i64, _p, _i, _p, _p, _p, _p, _i64, _p]),
    "btc_rulebook_tile_order_ints": (_i64, [_i]),
    "btc_index_summary_words": (_i64, [_i64]),
    "btc_rulebook_conv_sparse_workspace_bytes": (_i64, [_i64]),
    "btc_rulebook_conv_sparse": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _i64, _p, _p, _i, _p, _p, _p, _p, _i64, _p]),
    "btc_index_clear_sparse": (_i, [_p, _i, _p, _i, _p, _p, _i64, _p, _p]),
    "btc_rulebook_pairs_workspace_bytes": (_i64, [_i, _i]),
    "btc_rulebook_pairs": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _i64, _p]),
    "btc_sparse_conv_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _p]),
    "btc_sparse_conv_tc_supported": (_i, [_i, _i, _i]),
    "btc_sparse_conv_tc_config": (_i, [_i, _i, _i]),
    "btc_sparse_conv_tc_diag": (_i, [_i]),
    "btc_sparse_conv_tc_trace": (_i, [_p]),
    "btc_sparse_conv_tc_grid": (_i, [_i]),
    "btc_sparse_conv_tc_split_supported": (_i, [_i, _i, _i, _i, _i]),
    "btc_sparse_conv_tc_split_packed_bytes": (_i64, [_i, _i, _i]),
    "btc_sparse_conv_tc_pack_split": (_i, [_p, _i, _i, _i, _p, _p]),
    "btc_features_to_split": (_i, [_p, _i, _p, _i, _p, _p]),
    "btc_features_from_split": (_i, [_p, _i, _p, _i, _p, _p]),
    "btc_sparse_conv_fwd_tc_split": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _p, _p, _p]),
    "btc_sparse_conv_tc_packed_bytes": (_i64, [_i, _i, _i]),
    "btc_sparse_conv_tc_pack": (_i, [_p, _i, _i, _i, _p, _p]),
    "btc_sparse_conv_fwd_tc": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p]),
    "btc_rulebook_tile_meta": (_i, [_p, _i, _p, _i, _p, _p, _p]),
    "btc_sparse_conv_fwd_tc_meta": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p, _p, _p]),
    "btc_sparse_conv_fwd_tc_rows": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _p]),
    "btc_rulebook_sort_rows": (_i, [_p, _i, _p, _i, _p, _p, _p]),
    "btc_sparse_conv_bwd_workspace_bytes": (_i64, [_i, _i, _i]),
    "btc_sparse_conv_bwd_data": (_i, [_p, _p, _i, _p, _p, _i, _p, _i, _i, _i, _p, _i64, _p]),
    "btc_sparse_conv_bwd_weight": (_i, [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _p]),
    "btc_maxpool_fwd": (_i, [_p, _p, _p, _i, _p, _i, _i, _p]),
    "btc_maxpool_bwd": (_i, [_p, _p, _p, _p, _p, _i, _i, _p, _i, _i, _p]),
    "btc_copy_rows": (_i, [_p, _p, _i, _p, _i, _p]),
    "btc_to_dense": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p]),
    "btc_from_dense": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p]),
    "btc_occ_targets_workspace_bytes": (_i64, [_i, _p, _p]),
    "btc_occ_targets": (_i, [_p, _i, _i, _p, _p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_select_workspace_bytes": (_i64, [_i, _p]),
    "btc_occ_select": (_i, [_p, _p, _i, _p, ctypes.c_float, _p, _p, _p, ctypes.c_float, _i, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_abs_mean_vfe": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _p]),
    "btc_occ_vfe": (_i, [_p, _p, _i, _p, _i, _i, _i, _p, _p, _p]),
    "btc_occ_head_prob": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _p]),
    "btc_boxes_bev": (_i, [_p, _i, _p, _i, _i, _p, _p]),
    "btc_nms_workspace_bytes": (_i64, [_i]),
    "btc_nms": (_i, [_p, _i, ctypes.c_float, _i, _p, _p, _p, _i64, _p]),
    "btc_ball_query_stack": (_i, [_i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p]),
    "btc_group_points_stack": (_i, [_i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "btc_group_points_stack_grad": (_i, [_i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p]),
    "btc_trilinear_sparse_workspace_bytes": (_i64, [_i64, _i, _p]),
    "btc_trilinear_sparse_flag": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _i64, _i64, _i, _p, _p, _i64, _p]),
    "btc_trilinear_sparse_grad": (_i, [_p, _p, _i, _p, _i, _i, _p, _p, _p, _i64, _i64, _i, _p, _p, _i64, _p]),
    "btc_trilinear_sparse_emit": (_i, [_p, _i, _i, _p, _p, _p, _i64, _i64, _i, _i, _p, _i, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_box_targets_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "btc_occ_box_targets": (_i, [_p, _i, _i, _p, _p, _i, _p, _i, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i,
                                 _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_box_targets_v2": (_i, [_p, _i, _i, _p, _p, _i, _p, _i, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _i, _i, _i,
                                    _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_loss_maps": (_i, [_p] * 10 + [_i, _p] + [_p] * 11),
    "btc_revoxelize_workspace_bytes": (_i64, [_i, _i64]),
    "btc_revoxelize": (_i, [_p, _i, _p, _i, _p, _p, _i64, _p, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_revoxelize_fill": (_i, [_p, _p, _p, _i, _p, _i, _i, _p, _i, _p]),
    # aliases under the names of SURVEY §8(b)'s minimum export set
    "btc_voxelize_cuda": (_i, [_p, _i, _i, _p, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
    "btc_rulebook_pool": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _i, _p, _p, _p, _p, _i64, _p]),
    "btc_occ_inject_revoxelize": (_i, [_p, _i, _p, _i, _p, _p, _i64, _p, _i, _p, _p, _p, _p, _p, _p, _i64, _p]),
}

_lib = None


class BtcError(RuntimeError):
    pass


def load():
    """Load the library once; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BtcError(
            "CUDA extension %s is missing — run `python -m btcdet_b200.build` "
            "(there is no CPU fallback in this package)" % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.btc_compiled_sm() != 100:
        raise BtcError("libbtcdet_b200.so was not built for sm_100a")
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().btc_last_error()
        raise BtcError("%s failed with status %d: %s" % (what, rc, msg.decode() if msg else ""))


def int3(values):
    """Host int[3] array for geometry arguments."""
    a = (ctypes.c_int * 3)(*[int(v) for v in values])
    return a


def float_array(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])
