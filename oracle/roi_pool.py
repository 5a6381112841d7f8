"""Oracle of the RoI grid pooling ops — TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may import this module; nothing under btcdet_b200/
or spconv/ does.

* ball query / grouping: C restatement of the reference's own in-tree CUDA kernels, thread for thread
  (oracle.c: orc_ball_query_stack, orc_group_points_stack(_grad); reference
  btcdet/ops/pointnet2/pointnet2_stack/src/ball_query_gpu.cu:16-60, group_points_gpu.cu:16-95).  PINNED on the GPU box
  against those kernels themselves, compiled from the checkout (`make -C oracle ref` -> oracle/_ref/libpointnet2_ref.so,
  loaded by `RefPointnet2` below).
* reverse trilinear gather + compaction: torch restatement, operation for operation, of
  `reverse_sparse_trilinear_interpolate_torch` (btcdet/utils/common_utils.py:247-311) and
  `ConvHead.interpolate_from_3d_features` (btcdet/models/roi_heads/conv_head.py:505-528); device agnostic, so it is the
  checker on the CPU (PINNED there against the reference's own function, tests/test_roi_pool_cpu.py) and on the GPU
  (same ATen kernels as the reference's code).
"""
import ctypes
import os

import numpy as np
import torch

from . import oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libpointnet2_ref.so")


def _lib():
    lib = _o._load()
    if not getattr(lib, "_roi_bound", False):
        P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        lib.orc_ball_query_stack.restype = None
        lib.orc_ball_query_stack.argtypes = [I, I, F, I, P, P, P, P, P]
        lib.orc_group_points_stack.restype = None
        lib.orc_group_points_stack.argtypes = [I, I, I, I, P, P, P, P, P]
        lib.orc_group_points_grad_stack.restype = None
        lib.orc_group_points_grad_stack.argtypes = [I, I, I, I, I, P, P, P, P, P]
        lib._roi_bound = True
    return lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def ball_query_stack(radius, nsample, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt):
    """What the reference kernel leaves in the zero-initialised idx [M, nsample] (before pointnet2_utils.py:38-40
    turns the -1 of empty balls into 0)."""
    xyz, new_xyz = _f32(xyz), _f32(new_xyz)
    cnt, ncnt = _i32(xyz_batch_cnt), _i32(new_xyz_batch_cnt)
    M = new_xyz.shape[0]
    idx = np.zeros((M, int(nsample)), dtype=np.int32)
    _lib().orc_ball_query_stack(len(cnt), M, ctypes.c_float(float(radius)), int(nsample), _o._p(new_xyz), _o._p(ncnt),
                                _o._p(xyz), _o._p(cnt), _o._p(idx))
    return idx


def group_points_stack(features, features_batch_cnt, idx, idx_batch_cnt):
    features, idx = _f32(features), _i32(idx)
    fc, ic = _i32(features_batch_cnt), _i32(idx_batch_cnt)
    M, ns = idx.shape
    C = features.shape[1]
    out = np.zeros((M, C, ns), dtype=np.float32)
    _lib().orc_group_points_stack(len(ic), M, C, ns, _o._p(features), _o._p(fc), _o._p(idx), _o._p(ic), _o._p(out))
    return out


def group_points_grad_stack(grad_out, idx, idx_batch_cnt, features_batch_cnt, N):
    grad_out, idx = _f32(grad_out), _i32(idx)
    fc, ic = _i32(features_batch_cnt), _i32(idx_batch_cnt)
    M, C, ns = grad_out.shape
    g = np.zeros((int(N), C), dtype=np.float32)
    _lib().orc_group_points_grad_stack(len(ic), M, C, int(N), ns, _o._p(grad_out), _o._p(idx), _o._p(ic), _o._p(fc), _o._p(g))
    return g


class RefPointnet2(object):
    """The reference's OWN ball-query / grouping CUDA kernels (compiled from /root/reference by `make -C oracle ref`,
    launched on the default stream like the reference's extension does).  GPU box only."""

    def __init__(self):
        self.lib = ctypes.CDLL(REF_LIB)
        P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
        self.lib.ref_ball_query_stack.argtypes = [I, I, F, I, P, P, P, P, P]
        self.lib.ref_group_points_stack.argtypes = [I, I, I, I, P, P, P, P, P]
        self.lib.ref_group_points_grad_stack.argtypes = [I, I, I, I, I, P, P, P, P, P]

    @staticmethod
    def available():
        return os.path.exists(REF_LIB)

    def ball_query(self, radius, nsample, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt):
        M = new_xyz.shape[0]
        idx = torch.zeros((M, nsample), dtype=torch.int32, device=xyz.device)
        torch.cuda.synchronize()
        self.lib.ref_ball_query_stack(xyz_batch_cnt.shape[0], M, ctypes.c_float(float(radius)), nsample, new_xyz.data_ptr(),
                                      new_xyz_batch_cnt.data_ptr(), xyz.data_ptr(), xyz_batch_cnt.data_ptr(), idx.data_ptr())
        torch.cuda.synchronize()
        return idx

    def group_points(self, features, features_batch_cnt, idx, idx_batch_cnt):
        M, ns = idx.shape
        C = features.shape[1]
        out = torch.zeros((M, C, ns), dtype=torch.float32, device=features.device)
        torch.cuda.synchronize()
        self.lib.ref_group_points_stack(idx_batch_cnt.shape[0], M, C, ns, features.data_ptr(), features_batch_cnt.data_ptr(),
                                        idx.data_ptr(), idx_batch_cnt.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        return out

    def group_points_grad(self, grad_out, idx, idx_batch_cnt, features_batch_cnt, N):
        M, C, ns = grad_out.shape
        g = torch.zeros((N, C), dtype=torch.float32, device=grad_out.device)
        torch.cuda.synchronize()
        self.lib.ref_group_points_grad_stack(idx_batch_cnt.shape[0], M, C, N, ns, grad_out.data_ptr(), idx.data_ptr(),
                                             idx_batch_cnt.data_ptr(), features_batch_cnt.data_ptr(), g.data_ptr())
        torch.cuda.synchronize()
        return g


# ---- reverse trilinear gather (torch, device agnostic) ----------------------------------------------------------------
def dense_volume(feats, coords, batch, spatial_shape):
    """What SparseConvTensor.dense() returns: zeros [B, C, Z, Y, X] with the rows index-assigned."""
    Z, Y, X = [int(v) for v in spatial_shape]
    im = torch.zeros((batch, Z, Y, X, feats.shape[1]), dtype=feats.dtype, device=feats.device)
    c = coords.long()
    im[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = feats
    return im.permute(0, 4, 1, 2, 3).contiguous()


def reverse_trilinear(feats, coords, batch, spatial_shape, b, zyx, normalize=False):
    """common_utils.py:247-311, line for line in meaning: returns [T, C]."""
    im = dense_volume(feats, coords, batch, spatial_shape)
    x, y, z = zyx[..., 2], zyx[..., 1], zyx[..., 0]
    x0 = torch.floor(x).long(); x1 = x0 + 1
    y0 = torch.floor(y).long(); y1 = y0 + 1
    z0 = torch.floor(z).long(); z1 = z0 + 1
    S = [int(v) for v in spatial_shape]
    if normalize:
        zm0 = zm1 = ym0 = ym1 = xm0 = xm1 = 1
    else:
        zm0 = ((z0 >= 0) & (z0 < S[0])).unsqueeze(-1); zm1 = ((z1 >= 0) & (z1 < S[0])).unsqueeze(-1)
        ym0 = ((y0 >= 0) & (y0 < S[1])).unsqueeze(-1); ym1 = ((y1 >= 0) & (y1 < S[1])).unsqueeze(-1)
        xm0 = ((x0 >= 0) & (x0 < S[2])).unsqueeze(-1); xm1 = ((x1 >= 0) & (x1 < S[2])).unsqueeze(-1)
    dz1, dz0 = z1.type_as(z) - z, z0.type_as(z) - z
    dy1, dy0 = y1.type_as(y) - y, y0.type_as(y) - y
    dx1, dx0 = x1.type_as(x) - x, x0.type_as(x) - x
    w000 = torch.abs(dz1 * dy1 * dx1); w010 = torch.abs(-dz1 * dy0 * dx1)
    w001 = torch.abs(-dz1 * dy1 * dx0); w011 = torch.abs(dz1 * dy0 * dx0)
    w100 = torch.abs(-dz0 * dy1 * dx1); w110 = torch.abs(dz0 * dy0 * dx1)
    w101 = torch.abs(dz0 * dy1 * dx0); w111 = torch.abs(-dz0 * dy0 * dx0)
    x0 = torch.clamp(x0, 0, S[2] - 1); x1 = torch.clamp(x1, 0, S[2] - 1)
    y0 = torch.clamp(y0, 0, S[1] - 1); y1 = torch.clamp(y1, 0, S[1] - 1)
    z0 = torch.clamp(z0, 0, S[0] - 1); z1 = torch.clamp(z1, 0, S[0] - 1)
    I000 = im[b, :, z0, y0, x0] * zm0 * ym0 * xm0; I010 = im[b, :, z0, y1, x0] * zm0 * ym1 * xm0
    I001 = im[b, :, z0, y0, x1] * zm0 * ym0 * xm1; I011 = im[b, :, z0, y1, x1] * zm0 * ym1 * xm1
    I100 = im[b, :, z1, y0, x0] * zm1 * ym0 * xm0; I110 = im[b, :, z1, y1, x0] * zm1 * ym1 * xm0
    I101 = im[b, :, z1, y0, x1] * zm1 * ym0 * xm1; I111 = im[b, :, z1, y1, x1] * zm1 * ym1 * xm1
    return (I000 * w000.unsqueeze(-1) + I010 * w010.unsqueeze(-1) + I001 * w001.unsqueeze(-1) + I011 * w011.unsqueeze(-1)
            + I100 * w100.unsqueeze(-1) + I110 * w110.unsqueeze(-1) + I101 * w101.unsqueeze(-1) + I111 * w111.unsqueeze(-1))


def target_indices(conv_grid_points, point_cloud_range, voxel_size, stride):
    """conv_head.py:513-517: [B, NP, 3] world points -> [B*NP, 3] fractional (z, y, x) indices."""
    if isinstance(stride, int):
        stride = [stride, stride, stride]
    x = (conv_grid_points[:, :, 0] - point_cloud_range[0]) / voxel_size[0] / stride[2] - 0.5
    y = (conv_grid_points[:, :, 1] - point_cloud_range[1]) / voxel_size[1] / stride[1] - 0.5
    z = (conv_grid_points[:, :, 2] - point_cloud_range[2]) / voxel_size[2] / stride[0] - 0.5
    return torch.stack([z.view(-1), y.view(-1), x.view(-1)], dim=-1)


def interpolate_rows(feats, coords, batch, spatial_shape, zyx, per_scene, local_shape, normalize=False):
    """conv_head.py:518-528 for targets zyx [T, 3]: (coords int64 [n, 4] = (t // P, z, y, x of cell t % P), rows [n, C],
    target indices [n])."""
    T = zyx.shape[0]
    t = torch.arange(T, device=zyx.device)
    b = t // int(per_scene)
    feat = reverse_trilinear(feats, coords, batch, spatial_shape, b, zyx, normalize)
    inds = torch.nonzero(torch.any(torch.abs(feat) > 0.0, dim=-1))[..., 0]
    lz, ly, lx = [int(v) for v in local_shape]
    P = lz * ly * lx
    cell = inds % P
    bzyx = torch.stack([inds // P, cell // (ly * lx), (cell // lx) % ly, cell % lx], dim=-1)
    return bzyx, feat[inds, :], inds
