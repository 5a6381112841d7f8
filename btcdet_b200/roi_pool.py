"""RoI grid pooling of the second stage — host side above the C ABI.

Mirrors the pieces of `ConvHead.roi_conv_pool` (btcdet/models/roi_heads/conv_head.py:247-379) that are not plain dense
torch layers, with the reference's names, argument meaning and return values:

* `interpolate_from_3d_features` — `ConvHead.interpolate_from_3d_features` (conv_head.py:505-528) =
  `reverse_sparse_trilinear_interpolate_torch` (btcdet/utils/common_utils.py:247-311) + non-zero-row compaction, fused:
  no `feat.dense()` volume, no eight `[T, C]` corner gathers (T = B * N_roi * 27 * 96 targets, 340 MB each at the yaml's
  training shape); rows are bit-identical to the reference expression's.
* `stack_sa_msg_forward` — `StackSAModuleMSG.forward` (pointnet2_stack/pointnet2_modules.py:55-108) with the ball queries
  of all its radii in ONE launch (`btc_ball_query_stack`) instead of one launch per radius; grouping, the reference's
  rotation / scaling and the shared MLPs are unchanged.
* `patch_conv_head(head)` — binds both onto a reference `ConvHead` instance (everything else of the head is the
  reference's own code running on the `spconv` shim).

Exact mode reads one count from the device (the reference's `torch.nonzero` synchronises at the same place);
`out_cap=` gives the static form (capacity-sized outputs + device count, no host read).
"""
import ctypes
import types

import torch
import torch.nn.functional as F

from . import _lib
from . import pointnet2_stack_cuda as _p2
from ._lib import check


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def target_indices(conv_grid_points, point_cloud_range, voxel_size, stride):
    """[B, NP, 3] world points -> [B*NP, 3] fractional (z, y, x) voxel indices at the feature level `stride` (z, y, x):
    the expression of conv_head.py:513-516, evaluated by the same torch kernels (tensor / python scalar)."""
    if isinstance(stride, int):
        stride = [stride, stride, stride]
    x = (conv_grid_points[:, :, 0] - point_cloud_range[0]) / voxel_size[0] / stride[2] - 0.5
    y = (conv_grid_points[:, :, 1] - point_cloud_range[1]) / voxel_size[1] / stride[1] - 0.5
    z = (conv_grid_points[:, :, 2] - point_cloud_range[2]) / voxel_size[2] / stride[0] - 0.5
    return torch.stack([z.reshape(-1), y.reshape(-1), x.reshape(-1)], dim=-1)


class _TrilinearRows(torch.autograd.Function):
    """rows = non-zero rows of the reverse trilinear gather; differentiable w.r.t. the sparse source features."""

    @staticmethod
    def forward(ctx, feats, coords, n_dev, shape, batch, zyx, bt, per_scene, normalize, lshape, out_cap):
        lib = _lib.load()
        T, C = int(zyx.shape[0]), int(feats.shape[1])
        P = lshape[0] * lshape[1] * lshape[2]
        dev = feats.device
        shp = _lib.int3(shape)
        ws_bytes = int(lib.btc_trilinear_sparse_workspace_bytes(T, batch, shp))
        if ws_bytes < 0:
            raise RuntimeError("btc_trilinear_sparse_workspace_bytes: bad sizes (T=%d, batch=%d, shape=%s)" % (T, batch, shape))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        check(lib.btc_trilinear_sparse_flag(_ptr(feats), _ptr(coords), int(feats.shape[0]), _ptr(n_dev), C, batch, shp,
                                            _ptr(zyx), _ptr(bt), T, per_scene, normalize, _ptr(count), _ptr(ws), ws_bytes,
                                            _stream()), "btc_trilinear_sparse_flag")
        static = out_cap is not None
        n = int(out_cap) if static else int(count.item())
        alloc = torch.zeros if static else torch.empty
        out_f = alloc((n, C), dtype=torch.float32, device=dev)
        out_c = alloc((n, 4), dtype=torch.int32, device=dev)
        out_t = torch.full((n,), -1, dtype=torch.int64, device=dev)
        check(lib.btc_trilinear_sparse_emit(_ptr(feats), C, batch, shp, _ptr(zyx), _ptr(bt), T, per_scene, normalize, P,
                                            _lib.int3(lshape), n, _ptr(out_f), _ptr(out_c), _ptr(out_t), _ptr(ws), ws_bytes,
                                            _stream()), "btc_trilinear_sparse_emit")
        ctx.save_for_backward(zyx, out_t, ws, count)
        ctx.bt = bt
        ctx.meta = (shape, batch, per_scene, normalize, tuple(feats.shape), static)
        ctx.mark_non_differentiable(out_c, out_t, count)
        return out_f, out_c, out_t, count

    @staticmethod
    def backward(ctx, g_rows, *_):
        zyx, out_t, ws, count = ctx.saved_tensors
        shape, batch, per_scene, normalize, fshape, static = ctx.meta
        g_rows = g_rows.contiguous().float()
        grad = torch.zeros(fshape, dtype=torch.float32, device=g_rows.device)
        check(_lib.load().btc_trilinear_sparse_grad(_ptr(g_rows), _ptr(out_t), int(g_rows.shape[0]),
                                                    _ptr(count) if static else None, fshape[1], batch, _lib.int3(shape),
                                                    _ptr(zyx), _ptr(ctx.bt), int(zyx.shape[0]), per_scene, normalize,
                                                    _ptr(grad), _ptr(ws), int(ws.numel()), _stream()),
              "btc_trilinear_sparse_grad")
        return (grad,) + (None,) * 10


def trilinear_gather_rows(features, zyx, per_scene, local_shape, normalize=False, out_cap=None, b_target=None,
                          want_target=False):
    """Rows of `reverse_sparse_trilinear_interpolate_torch(features, b, zyx)` that have a non-zero channel, in target
    order, with their mini-grid coordinates (t // P, unravel(t % P, local_shape)), P = prod(local_shape).

    features: spconv.SparseConvTensor (3-D); zyx [T, 3] float32; scene of target t = b_target[t] or t // per_scene.
    Returns (coords int32 [n, 4], rows float32 [n, C][, target int64 [n]]); with `out_cap` the outputs are capacity
    sized and the int32 [1] device count is appended (no host read).  Differentiable w.r.t. `features.features`."""
    feats = features.features
    if not feats.is_cuda:
        raise RuntimeError("btcdet_b200.roi_pool: CUDA tensors only (there is no CPU path)")
    feats = feats.contiguous().float()
    bt = None if b_target is None else b_target.contiguous().long()
    out_f, out_c, out_t, count = _TrilinearRows.apply(
        feats, features._coords4(), features.n_dev, [int(v) for v in features._shape3()], int(features.batch_size),
        zyx.detach().contiguous().float(), bt, int(per_scene), int(bool(normalize)), [int(v) for v in local_shape],
        out_cap)
    res = (out_c, out_f) + ((out_t,) if want_target else ())
    return res + ((count,) if out_cap is not None else ())


def interpolate_from_3d_features(conv_grid_points, dense_idx, features, stride, point_cloud_range, voxel_size,
                                 normalize=False):
    """`ConvHead.interpolate_from_3d_features` (conv_head.py:505-528): conv_grid_points [B, NP, 3], dense_idx
    [BN, P, 3] (z, y, x of every cell of a mini grid, `get_dense_grid_points`), features = the sparse source tensor.
    Returns (bzyx [n, 4] float32, feat [n, C]) exactly like the reference (which stacks a long with float columns)."""
    B, NP, _ = list(conv_grid_points.shape)
    BN, P, _ = list(dense_idx.shape)
    assert B * NP == BN * P, "one target per cell of every mini grid"
    # the mini-grid shape from dense_idx (row-major nonzero() order of ones(lz, ly, lx): the last row is the far corner)
    local_shape = [int(v) + 1 for v in dense_idx[0, -1].tolist()]
    zyx = target_indices(conv_grid_points, point_cloud_range, voxel_size, stride)
    coords, rows = trilinear_gather_rows(features, zyx, NP, local_shape, normalize=normalize)
    return coords.float(), rows


def stack_sa_msg_forward(module, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features=None, empty_voxel_set_zeros=True,
                         rotateMatrix=None, xyscales=None, zscales=None, vis=False):
    """`StackSAModuleMSG.forward` (pointnet2_modules.py:55-108) for a reference module instance `module`, with one
    fused multi-radius ball query.  Returns (new_xyz, new_features [M, sum_k C_k]) like the reference (vis unsupported)."""
    assert not vis, "visualisation outputs are not produced by the fused path"
    groupers = list(module.groupers)
    radii = [g.radius for g in groupers]
    nsamples = [int(g.nsample) for g in groupers]
    assert all(not isinstance(r, (list, tuple)) for r in radii), "shell queries are not on this path"
    M = int(new_xyz.shape[0])
    idx_all = []
    for lo in range(0, len(radii), 4):
        part = [torch.empty((M, ns), dtype=torch.int32, device=xyz.device) for ns in nsamples[lo:lo + 4]]
        _p2.ball_query_multi(radii[lo:lo + 4], nsamples[lo:lo + 4], new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, part)
        idx_all += part
    B = int(xyz_batch_cnt.shape[0])
    out = []
    for k, g in enumerate(groupers):
        idx = idx_all[k]
        empty = idx[:, 0] == -1                                   # pointnet2_utils.py:38-40
        idx[empty] = 0
        ns = nsamples[k]

        def group(t):
            t = t.contiguous()
            o = torch.empty((M, t.shape[1], ns), dtype=torch.float32, device=t.device)
            _p2.group_points_wrapper(B, M, int(t.shape[1]), ns, t, xyz_batch_cnt, idx, new_xyz_batch_cnt, o)
            return o

        grouped_xyz = group(xyz)                                  # (M, 3, ns)
        grouped_xyz -= new_xyz.unsqueeze(-1)
        grouped_xyz[empty] = 0
        if rotateMatrix is not None:
            grouped_xyz = g.rotate(grouped_xyz, rotateMatrix)     # the reference's einsum (pointnet2_utils.py:188-196)
        if xyscales is not None:
            grouped_xyz[..., :2, :] = grouped_xyz[..., :2, :] / xyscales
            grouped_xyz[..., 2:3, :] = grouped_xyz[..., 2:3, :] / zscales
        if features is not None:
            grouped_features = group(features)
            grouped_features[empty] = 0
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if g.use_xyz else grouped_features
        else:
            new_features = grouped_xyz
        new_features = new_features.permute(1, 0, 2).unsqueeze(dim=0)          # (1, C, M, ns)
        new_features = module.mlps[k](new_features)
        if module.pool_method == 'max_pool':
            new_features = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(dim=-1)
        elif module.pool_method == 'avg_pool':
            new_features = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)]).squeeze(dim=-1)
        else:
            raise NotImplementedError
        out.append(new_features.squeeze(dim=0).permute(1, 0))
    return new_xyz, torch.cat(out, dim=1)


def patch_conv_head(head, fuse_ball_queries=True):
    """Bind the fused ops onto a reference `ConvHead` instance: `interpolate_from_3d_features` and (inference /
    no-grad use: the fused grouping has no autograd edge to the point features) the `StackSAModuleMSG.forward`s."""
    def _interp(self, conv_grid_points, dense_idx, features, stride):
        return interpolate_from_3d_features(conv_grid_points, dense_idx, features, stride, self.point_cloud_range,
                                            self.det_voxel_size, normalize=self.intrp_norm)

    head.interpolate_from_3d_features = types.MethodType(_interp, head)
    if fuse_ball_queries:
        for name in ("SA_rawpoints", "SA_occpoints"):
            mod = getattr(head, name, None)
            if mod is not None:
                mod.forward = types.MethodType(
                    lambda self, *a, **kw: stack_sa_msg_forward(self, *a, **kw), mod)
    return head
