"""Drop-in for the reference's compiled extension `btcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda`, for the
entry points the RoI head uses (`StackSAModuleMSG` under `ConvHead.roi_conv_pool`).

Same names and calling convention as `btcdet/ops/pointnet2/pointnet2_stack/src/pointnet2_api.cpp:11-23` — the caller
pre-allocates the outputs, the functions return an int — so `pointnet2_stack/pointnet2_utils.py` (BallQuery,
GroupingOperation, QueryAndGroup) and `pointnet2_modules.py` (StackSAModuleMSG) run unchanged on top of it:

    import sys, btcdet_b200.pointnet2_stack_cuda as m
    sys.modules["btcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda"] = m     # before pointnet2_utils is imported

The kernels are this library's (`csrc/roi_pool_kernels.cuh`); results are identical to the reference kernels' bit for
bit (`tests/test_roi_pool_gpu.py` runs both on the same inputs).  Not provided (not on the BtcDet path, the yaml's
ConvHead uses neither): shell_query, furthest_point_sampling, three_nn / three_interpolate.
"""
import ctypes

import torch

from . import _lib
from ._lib import check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(*ts):
    for t in ts:   # the reference's CHECK_INPUT (ball_query.cpp:14-18): CUDA + contiguous
        if not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("pointnet2_stack: tensors must be contiguous CUDA tensors")


def ball_query_multi(radii, nsamples, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx_list):
    """All radii of one module in ONE launch: idx_list[r] (M, nsamples[r]) int32 receives what the reference's kernel
    writes into a zero-initialised idx for radius radii[r] (1 <= len(radii) <= 4)."""
    _check(new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, *idx_list)
    assert new_xyz.dtype == xyz.dtype == torch.float32
    assert new_xyz_batch_cnt.dtype == xyz_batch_cnt.dtype == torch.int32
    n = len(radii)
    assert 1 <= n <= 4 and len(nsamples) == n and len(idx_list) == n
    M, B = new_xyz.shape[0], xyz_batch_cnt.shape[0]
    for t, ns in zip(idx_list, nsamples):
        assert t.dtype == torch.int32 and tuple(t.shape) == (M, int(ns))
    rad = (ctypes.c_float * n)(*[float(r) for r in radii])
    nsm = (ctypes.c_int * n)(*[int(s) for s in nsamples])
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in idx_list])
    check(_lib.load().btc_ball_query_stack(B, M, n, rad, nsm, _ptr(new_xyz), _ptr(new_xyz_batch_cnt), _ptr(xyz),
                                           _ptr(xyz_batch_cnt), ptrs, _stream()), "btc_ball_query_stack")
    return 1


def ball_query_wrapper(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx):
    """ball_query_wrapper_stack (ball_query.cpp:21-39): idx (M, nsample) int32, zero-initialised by the caller."""
    assert xyz_batch_cnt.shape[0] == B and new_xyz.shape[0] == M
    return ball_query_multi([radius], [nsample], new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, [idx])


def group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    """group_points_wrapper_stack (group_points.cpp): out (M, C, nsample) = features[start_b + idx[m, s], c]."""
    _check(features, features_batch_cnt, idx, idx_batch_cnt, out)
    assert features.dtype == out.dtype == torch.float32 and idx.dtype == torch.int32
    assert tuple(out.shape) == (M, C, nsample) and tuple(idx.shape) == (M, nsample) and features.shape[1] == C
    check(_lib.load().btc_group_points_stack(B, M, C, nsample, _ptr(features), _ptr(features_batch_cnt), _ptr(idx),
                                             _ptr(idx_batch_cnt), _ptr(out), _stream()), "btc_group_points_stack")
    return 1


def group_points_grad_wrapper(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    """group_points_grad_wrapper_stack: scatter-add of grad_out (M, C, nsample) into the zeroed grad_features (N, C)."""
    _check(grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features)
    assert grad_out.dtype == grad_features.dtype == torch.float32 and idx.dtype == torch.int32
    assert tuple(grad_out.shape) == (M, C, nsample) and tuple(grad_features.shape) == (N, C)
    check(_lib.load().btc_group_points_stack_grad(B, M, C, N, nsample, _ptr(grad_out), _ptr(idx), _ptr(idx_batch_cnt),
                                                  _ptr(features_batch_cnt), _ptr(grad_features), _stream()),
          "btc_group_points_stack_grad")
    return 1
