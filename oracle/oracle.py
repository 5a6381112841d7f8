"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle/oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  PARITY UNPINNED at the spconv-1.2.1 boundary (no reference golden vectors
exist) — pinned instead against dense stock-torch formulations in tests/.

Integer work (voxel grouping, rulebooks) is restated in C (oracle.c, built by oracle/Makefile);
the per-offset gather -> mm -> scatter-add arithmetic is restated with CPU torch ops, which is what
spconv 1.2.1's CPU path does (torch::mm per kernel offset, src/spconv/spconv_ops.cc) [App. A.5].
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None


def build():
    """gcc-compile oracle.c (building the checker is not using it)."""
    src, mk = os.path.join(_HERE, "oracle.c"), os.path.join(_HERE, "Makefile")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < max(os.path.getmtime(src), os.path.getmtime(mk)):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        lib = ctypes.CDLL(_LIB)
        P, I = ctypes.c_void_p, ctypes.c_int
        lib.orc_points_to_voxel.restype = I
        lib.orc_points_to_voxel.argtypes = [P, I, I, P, P, P, I, I, P, P, P, P]
        lib.orc_rulebook.restype = I
        lib.orc_rulebook.argtypes = [P, I, I, P, P, P, P, P, P, I, I, P, I, P, P]
        lib.orc_indice_conv.restype = None
        lib.orc_indice_conv.argtypes = [P, I, P, I, I, I, P, P, I, I, P]
        lib.orc_indice_maxpool.restype = None
        lib.orc_indice_maxpool.argtypes = [P, I, I, I, P, P, I, P]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i3(v):
    return np.ascontiguousarray(np.array(v, dtype=np.int32))


def _triple(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple, np.ndarray)) else [int(v)] * 3


# ------------------------------------------------------------------------------------------------
class VoxelGeneratorV2:
    """spconv.utils.VoxelGeneratorV2 (CPU, sequential) — follows spconv 1.2.1
    spconv/utils/__init__.py + points_to_voxel_3d_np [App. A.1].  Keeps the dense int32 lookup
    volume allocated across calls like the original (360 MB for the KITTI det grid)."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = np.round((point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size).astype(np.int64)
        self._voxel_size, self._range, self._grid = voxel_size, point_cloud_range, grid_size
        self._max_num_points, self._max_voxels = int(max_num_points), int(max_voxels)
        self._lookup = None

    @property
    def grid_size(self):
        return self._grid

    def generate(self, points, max_voxels=None):
        lib = _load()
        max_voxels = int(max_voxels or self._max_voxels)
        points = np.ascontiguousarray(points, dtype=np.float32)
        n, c = points.shape
        if self._lookup is None:
            self._lookup = -np.ones(int(np.prod(self._grid)), dtype=np.int32)
        voxels = np.zeros((max_voxels, self._max_num_points, c), dtype=np.float32)
        coors = np.zeros((max_voxels, 3), dtype=np.int32)
        num = np.zeros((max_voxels,), dtype=np.int32)
        grid = _i3(self._grid)
        m = lib.orc_points_to_voxel(_p(points), n, c, _p(self._voxel_size), _p(self._range), _p(grid),
                                    self._max_num_points, max_voxels, _p(voxels), _p(coors), _p(num),
                                    _p(self._lookup))
        return {"voxels": voxels[:m], "coordinates": coors[:m], "num_points_per_voxel": num[:m], "voxel_num": m}


def voxelize_batch(scenes, voxel_size, point_cloud_range, max_num_points, max_voxels, gen=None):
    """Per-scene generate + collate_batch padding of the batch index
    (btcdet/datasets/dataset.py:187-192): coords become (b, z, y, x)."""
    gen = gen or VoxelGeneratorV2(voxel_size, point_cloud_range, max_num_points, max_voxels)
    vs, cs, ns = [], [], []
    for b, pts in enumerate(scenes):
        r = gen.generate(pts)
        vs.append(r["voxels"])
        cs.append(np.pad(r["coordinates"], ((0, 0), (1, 0)), mode="constant", constant_values=b))
        ns.append(r["num_points_per_voxel"])
    return np.concatenate(vs, 0), np.concatenate(cs, 0).astype(np.int32), np.concatenate(ns, 0)


# ------------------------------------------------------------------------------------------------
def conv_output_shape(in_shape, ksize, stride, padding, dilation):
    return [(i + 2 * p - d * (k - 1) - 1) // s + 1 for i, k, s, p, d in zip(in_shape, ksize, stride, padding, dilation)]


def deconv_output_shape(in_shape, ksize, stride, padding, dilation, output_padding):
    return [(i - 1) * s - 2 * p + k + op for i, k, s, p, op in zip(in_shape, ksize, stride, padding, output_padding)]


def get_indice_pairs(indices, batch_size, spatial_shape, ksize, stride=1, padding=0, dilation=1, out_padding=0,
                     subm=False, transpose=False):
    """spconv.ops.get_indice_pairs -> (outids [M,4], indice_pairs [2,K,N], indice_pair_num [K], out_shape),
    outputs ascending in flat key, pairs in canonical (offset, input row) order [App. A.4]."""
    lib = _load()
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    n = indices.shape[0]
    ksize, stride, padding, dilation, out_padding = map(_triple, (ksize, stride, padding, dilation, out_padding))
    in_shape = [int(v) for v in spatial_shape]
    if subm:
        out_shape, stride, padding = in_shape, [1, 1, 1], [k // 2 for k in ksize]
    elif transpose:
        out_shape = deconv_output_shape(in_shape, ksize, stride, padding, dilation, out_padding)
    else:
        out_shape = conv_output_shape(in_shape, ksize, stride, padding, dilation)
    K = int(np.prod(ksize))
    cap = n if subm else max(1, min(n * K, batch_size * int(np.prod(out_shape))))
    outids = np.zeros((max(cap, 1), 4), dtype=np.int32)
    pairs = np.zeros((2, K, max(n, 1)), dtype=np.int32)
    pair_num = np.zeros((K,), dtype=np.int32)
    m = lib.orc_rulebook(_p(indices), n, int(batch_size), _p(_i3(in_shape)), _p(_i3(out_shape)), _p(_i3(ksize)),
                         _p(_i3(stride)), _p(_i3(padding)), _p(_i3(dilation)), int(subm), int(transpose), _p(outids),
                         cap, _p(pairs), _p(pair_num))
    if m < 0:
        raise RuntimeError("orc_rulebook failed (%d)" % m)
    return outids[:m], pairs[:, :, :n], pair_num, out_shape


def indice_conv(features, weight, pairs, pair_num, n_out, subm=False, bias=None, use_c=False):
    """spconv Native indice_conv on CPU: out=0; subm centre first; per offset gather -> mm -> scatter-add."""
    K = pairs.shape[1]
    if use_c:
        lib = _load()
        f = np.ascontiguousarray(features, dtype=np.float32)
        w = np.ascontiguousarray(weight, dtype=np.float32).reshape(K, f.shape[1], -1)
        out = np.zeros((n_out, w.shape[2]), dtype=np.float32)
        pr = np.ascontiguousarray(pairs, dtype=np.int32)
        pn = np.ascontiguousarray(pair_num, dtype=np.int32)
        lib.orc_indice_conv(_p(f), f.shape[0], _p(w), K, f.shape[1], w.shape[2], _p(pr), _p(pn), n_out, int(subm),
                            _p(out))
        if bias is not None:
            out += np.asarray(bias, dtype=np.float32)
        return out
    f = torch.as_tensor(np.asarray(features), dtype=torch.float32)
    w = torch.as_tensor(np.asarray(weight), dtype=torch.float32).reshape(K, f.shape[1], -1)
    pr = torch.as_tensor(np.asarray(pairs)).long()
    out = torch.zeros((n_out, w.shape[2]), dtype=torch.float32)
    center = -1
    if subm:
        center = K // 2
        out = torch.mm(f, w[center])
    for k in range(K):
        nh = int(pair_num[k])
        if nh == 0 or k == center:
            continue
        buf = torch.mm(f[pr[0, k, :nh]], w[k])
        out.index_add_(0, pr[1, k, :nh], buf)
    if bias is not None:
        out += torch.as_tensor(np.asarray(bias), dtype=torch.float32)
    return out.numpy()


def indice_maxpool(features, pairs, pair_num, n_out):
    lib = _load()
    f = np.ascontiguousarray(features, dtype=np.float32)
    K = pairs.shape[1]
    out = np.zeros((n_out, f.shape[1]), dtype=np.float32)
    pr = np.ascontiguousarray(pairs, dtype=np.int32)
    pn = np.ascontiguousarray(pair_num, dtype=np.int32)
    lib.orc_indice_maxpool(_p(f), f.shape[0], f.shape[1], K, _p(pr), _p(pn), n_out, _p(out))
    return out


def dense(features, indices, spatial_shape, batch_size):
    """SparseConvTensor.dense(): [B, C, *spatial] [App. A.9]."""
    features = np.asarray(features, dtype=np.float32)
    indices = np.asarray(indices).astype(np.int64)
    out = np.zeros((batch_size, *[int(s) for s in spatial_shape], features.shape[1]), dtype=np.float32)
    out[indices[:, 0], indices[:, 1], indices[:, 2], indices[:, 3]] = features
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3))


def pairs_to_tables(pairs, pair_num, n_in, n_out):
    """Canonical pair list -> (nbr_out [n_out,K], nbr_in [n_in,K]) neighbour tables (for comparing
    with the CUDA library's table layout)."""
    K = pairs.shape[1]
    nbr_out = -np.ones((n_out, K), dtype=np.int32)
    nbr_in = -np.ones((n_in, K), dtype=np.int32)
    for k in range(K):
        nh = int(pair_num[k])
        i, o = pairs[0, k, :nh], pairs[1, k, :nh]
        nbr_out[o, k] = i
        nbr_in[i, k] = o
    return nbr_out, nbr_in


# ------------------------------------------------------------------------------------------------
class SparseTensor:
    """Minimal CPU stand-in for SparseConvTensor used by the oracle backbones."""

    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], int(batch_size)
        self.indice_dict = {}


def sparse_conv(x, weight, ksize, stride=1, padding=0, dilation=1, subm=False, transpose=False, indice_key=None,
                bias=None, out_padding=0):
    """SparseConvolution.forward of spconv 1.2.1 on the oracle ops (rulebook cache by indice_key)."""
    datas = x.indice_dict.get(indice_key) if indice_key is not None else None
    if datas is None:
        outids, pairs, pair_num, out_shape = get_indice_pairs(x.indices, x.batch_size, x.spatial_shape, ksize, stride,
                                                              padding, dilation, out_padding, subm, transpose)
        datas = (outids, pairs, pair_num, out_shape)
        if indice_key is not None:
            x.indice_dict[indice_key] = datas
    outids, pairs, pair_num, out_shape = datas
    feats = indice_conv(x.features, weight, pairs, pair_num, outids.shape[0], subm=subm, bias=bias)
    y = SparseTensor(feats, outids, out_shape, x.batch_size)
    y.indice_dict = x.indice_dict
    return y
