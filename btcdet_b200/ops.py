"""Torch-facing wrappers over the C ABI (include/btcdet_b200.h).

PyTorch is plumbing here: it owns device memory (caching allocator), streams and autograd
bookkeeping.  Every computation is a call into libbtcdet_b200.so on the current stream.
Mirrors the operator surface of spconv 1.2.1 `spconv.ops` (get_indice_pairs, indice_conv,
indice_maxpool) that the reference reaches from btcdet/models/backbones_3d/spconv_backbone.py.
"""
import ctypes
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import check, float_array, int3


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.BtcError("btcdet_b200 ops need CUDA tensors (no CPU fallback)")


def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3, v
        return [int(x) for x in v]
    return [int(v)] * 3


def conv_output_shape(in_shape, ksize, stride, padding, dilation):
    """spconv get_conv_output_size: (in + 2p - d(k-1) - 1)//s + 1."""
    return [(i + 2 * p - d * (k - 1) - 1) // s + 1 for i, k, s, p, d in zip(in_shape, ksize, stride, padding, dilation)]


def deconv_output_shape(in_shape, ksize, stride, padding, dilation, output_padding):
    """spconv get_deconv_output_size: (in-1)s - 2p + k + output_padding."""
    return [(i - 1) * s - 2 * p + k + op for i, k, s, p, op in zip(in_shape, ksize, stride, padding, output_padding)]


# ------------------------------------------------------------------------------------------
# coordinate index
# ------------------------------------------------------------------------------------------
@dataclass
class CoordIndex:
    """Coordinate -> row lookup of one site set.

    Sorted site sets (outputs of strided / transposed convs) carry the rank bitmap that produced
    them (`entries`, rows == ranks).  Unsorted site sets (voxeliser order, user tensors) carry a
    coordinate hash (`hash_keys`/`hash_vals`) — far cheaper than a bitmap on a 92 M-cell grid.
    """
    entries: Optional[torch.Tensor]  # int64 [n_entries]  ({bits, rank} packed) or None
    perm: Optional[torch.Tensor]  # int32 [N] rank->row, only for bitmaps over unsorted rows
    batch: int
    shape: Sequence[int]
    hash_keys: Optional[torch.Tensor] = None  # int64 [n_slots]
    hash_vals: Optional[torch.Tensor] = None  # int32 [n_slots]


def index_entries(batch, shape):
    return int(_lib.load().btc_index_entries(int(batch), int3(shape)))


def build_index(coords: torch.Tensor, batch: int, shape, need_perm=True, n_dev=None) -> CoordIndex:
    _require_cuda(coords)
    lib = _lib.load()
    assert coords.dtype == torch.int32 and coords.dim() == 2 and coords.shape[1] == 4 and coords.is_contiguous()
    n = coords.shape[0]
    n_entries = index_entries(batch, shape)
    entries = torch.zeros(n_entries, dtype=torch.int64, device=coords.device)
    ws_bytes = int(lib.btc_index_workspace_bytes(n_entries))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=coords.device)
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=coords.device) if need_perm else None
    check(lib.btc_index_build(_ptr(coords), n, _ptr(n_dev), int(batch), int3(shape), _ptr(entries), n_entries,
                              _ptr(perm), None, _ptr(ws), ws_bytes, _stream()), "btc_index_build")
    return CoordIndex(entries, perm, int(batch), list(shape))


def build_hash(coords: torch.Tensor, batch: int, shape, n_dev=None) -> CoordIndex:
    """Coordinate hash of an (unsorted) site set."""
    _require_cuda(coords)
    lib = _lib.load()
    assert coords.dtype == torch.int32 and coords.dim() == 2 and coords.shape[1] == 4 and coords.is_contiguous()
    n = coords.shape[0]
    n_slots = int(lib.btc_hash_slots(n))
    keys = torch.empty(n_slots, dtype=torch.int64, device=coords.device)
    vals = torch.empty(n_slots, dtype=torch.int32, device=coords.device)
    check(lib.btc_hash_build(_ptr(coords), n, _ptr(n_dev), int(batch), int3(shape), _ptr(keys), _ptr(vals), n_slots,
                             _stream()), "btc_hash_build")
    return CoordIndex(None, None, int(batch), list(shape), keys, vals)


# ------------------------------------------------------------------------------------------
# static (capacity) mode: device-side counts instead of exact shapes
# ------------------------------------------------------------------------------------------
class StaticChecks:
    """(device count, capacity, what) triples registered by the capacity-sized allocations of a static-mode forward;
    `verify()` reads them all with ONE device->host copy (after the forward / graph replay) and raises if a capacity
    was exceeded — the kernels clamp silently at their capacities."""

    def __init__(self):
        self.items = []

    def add(self, n_dev, cap, what):
        self.items.append((n_dev, int(cap), what))

    def verify(self):
        if not self.items:
            return []
        counts = torch.cat([n.reshape(1).to(torch.int32) for n, _, _ in self.items]).tolist()
        for c, (_, cap, what) in zip(counts, self.items):
            if c > cap:
                raise _lib.BtcError("static-mode capacity exceeded: %s has %d rows > capacity %d" % (what, c, cap))
        return counts


_static_checks: Optional[StaticChecks] = None


class static_checks:
    """`with ops.static_checks() as chk:` collects the capacity checks of the static-mode calls made inside."""

    def __enter__(self):
        global _static_checks
        self.prev, _static_checks = _static_checks, StaticChecks()
        return _static_checks

    def __exit__(self, *exc):
        global _static_checks
        _static_checks = self.prev
        return False


def _register_cap(n_dev, cap, what):
    if _static_checks is not None:
        _static_checks.add(n_dev, cap, what)


STATIC_GROWTH = 1.0   # capacity of a strided level relative to its input level (static mode; overflow is checked)


def _static_cap(n_in_cap, ksize, stride, transposed, cells):
    """Output capacity of a static-mode rulebook: the true bound (every input touches prod(ceil(k/s)) outputs, k^3 for
    stride-1 / transposed layers) where the grid is small, else STATIC_GROWTH x the input capacity (checked afterwards)."""
    true_bound = _out_bound(n_in_cap, ksize, stride, transposed, cells)
    if transposed or all(int(st) == 1 for st in stride):
        return true_bound
    return max(1, min(true_bound, max(1024, int(STATIC_GROWTH * n_in_cap))))


# ------------------------------------------------------------------------------------------
# rulebooks
# ------------------------------------------------------------------------------------------
@dataclass
class Rulebook:
    """What spconv keeps in `indice_dict[key]`, in the layout the B200 kernels want.

    nbr_out [N_out, K]: input row feeding output row o through offset k (or -1).
    nbr_in  [N_in, K] : output row fed by input row i through offset k (None for submanifold
                        rulebooks, whose nbr_in is the mirror image of nbr_out).
    """
    nbr_out: torch.Tensor
    nbr_in: Optional[torch.Tensor]
    out_coords: torch.Tensor
    n_in: int
    n_out: int
    K: int
    subm: bool
    in_shape: list
    out_shape: list
    ksize: list
    stride: list
    padding: list
    dilation: list
    transposed: bool = False
    out_index: Optional[CoordIndex] = None
    # static (capacity) mode: the tables are capacity-sized and the live row counts stay on the device — no host read,
    # the whole forward is CUDA-graph capturable (SparseConvTensor.n_dev); None = exact shapes
    n_in_dev: Optional[torch.Tensor] = None
    n_out_dev: Optional[torch.Tensor] = None
    _pairs: Optional[tuple] = field(default=None, repr=False)
    _sorted: Optional[tuple] = field(default=None, repr=False)
    _meta: Optional[tuple] = field(default=None, repr=False)

    def sorted_rows(self):
        """(nbr_sorted, out_rows): the output-stationary table with rows reordered by valid-offset mask inside
        windows of 2048 rows — what the block-skipping tcgen05 tile consumes (computed once per rulebook)."""
        if self._sorted is None:
            self._sorted = rulebook_sort_rows(self.nbr_out)
        return self._sorted

    def tile_meta(self):
        """(tile_mask [tiles] i64, tile_order) of nbr_out for the tcgen05 tile (btc_rulebook_tile_meta), computed once per
        rulebook; (None, None) for kernels with more than 64 offsets."""
        if self._meta is None:
            self._meta = rulebook_tile_meta(self.nbr_out, self.n_out_dev) if self.K <= 64 and self.nbr_out.shape[0] > 0 \
                else (None, None)
        return self._meta

    def inverse(self) -> "Rulebook":
        """Rulebook of SparseInverseConv3d: swap the pair directions."""
        if self.subm:
            nbr_out_inv = self.nbr_out.flip(1).contiguous()
            nbr_in_inv = None
        else:
            nbr_out_inv, nbr_in_inv = self.nbr_in, self.nbr_out
        return Rulebook(nbr_out_inv, nbr_in_inv, None, self.n_out, self.n_in, self.K, self.subm, self.out_shape,
                        self.in_shape, self.ksize, self.stride, self.padding, self.dilation, not self.transposed,
                        n_in_dev=self.n_out_dev, n_out_dev=self.n_in_dev)

    def pairs(self):
        """spconv-1.2.1-format (indice_pairs [2,K,N_in], indice_pair_num [K]) in canonical order."""
        if self._pairs is None:
            table = self.nbr_out if self.subm else self.nbr_in
            self._pairs = rulebook_pairs(table, self.n_in, self.K, mirror=self.subm)
        return self._pairs


def rulebook_subm(coords: torch.Tensor, batch: int, shape, ksize, dilation=1,
                  index: Optional[CoordIndex] = None, n_dev: Optional[torch.Tensor] = None) -> Rulebook:
    _require_cuda(coords)
    lib = _lib.load()
    ksize, dilation = _triple(ksize), _triple(dilation)
    for k in ksize:
        if k % 2 != 1:
            raise _lib.BtcError("submanifold convolution needs odd kernel sizes")
    n = coords.shape[0]
    K = ksize[0] * ksize[1] * ksize[2]
    if index is None:
        index = build_hash(coords, batch, shape, n_dev=n_dev)
    nbr = torch.empty((n, K), dtype=torch.int32, device=coords.device)
    if index.hash_keys is not None:
        check(lib.btc_rulebook_subm_hash(_ptr(coords), n, _ptr(n_dev), int(batch), int3(shape), int3(ksize), int3(dilation),
                                         _ptr(index.hash_keys), _ptr(index.hash_vals), index.hash_keys.numel(),
                                         _ptr(nbr), _stream()), "btc_rulebook_subm_hash")
    else:
        check(lib.btc_rulebook_subm(_ptr(coords), n, _ptr(n_dev), int(batch), int3(shape), int3(ksize), int3(dilation),
                                    _ptr(index.entries), index.entries.numel(), _ptr(index.perm), _ptr(nbr), _stream()),
              "btc_rulebook_subm")
    return Rulebook(nbr, None, coords, n, n, K, True, list(shape), list(shape), ksize, [1, 1, 1],
                    [k // 2 for k in ksize], dilation, False, index, n_in_dev=n_dev, n_out_dev=n_dev)


def _out_bound(n_in, ksize, stride, transposed, cells):
    if transposed:
        b = n_in * ksize[0] * ksize[1] * ksize[2]
    else:
        b = n_in
        for k, s in zip(ksize, stride):
            b *= -(-k // s)
    return max(1, min(b, cells))


def rulebook_conv(coords: torch.Tensor, batch: int, in_shape, ksize, stride=1, padding=0, dilation=1,
                  transposed=False, output_padding=0, out_cap: Optional[int] = None,
                  n_dev: Optional[torch.Tensor] = None) -> Rulebook:
    """Regular / transposed sparse conv (and pooling) rulebook.  One host read of n_out — none in static mode (`n_dev`
    given: `coords` is capacity-sized with n_dev live rows, the result is capacity-sized with its count on the device)."""
    _require_cuda(coords)
    lib = _lib.load()
    ksize, stride, padding, dilation = _triple(ksize), _triple(stride), _triple(padding), _triple(dilation)
    output_padding = _triple(output_padding)
    in_shape = [int(v) for v in in_shape]
    if transposed:
        out_shape = deconv_output_shape(in_shape, ksize, stride, padding, dilation, output_padding)
    else:
        out_shape = conv_output_shape(in_shape, ksize, stride, padding, dilation)
    if min(out_shape) <= 0:
        raise _lib.BtcError("sparse conv output shape %s is empty" % (out_shape,))
    n_in = coords.shape[0]
    K = ksize[0] * ksize[1] * ksize[2]
    dev = coords.device
    out_entries = index_entries(batch, out_shape)
    cells = batch * out_shape[0] * out_shape[1] * out_shape[2]
    static = n_dev is not None
    if out_cap is not None:
        cap = out_cap
    else:
        cap = _static_cap(n_in, ksize, stride, transposed, cells) if static else _out_bound(n_in, ksize, stride, transposed, cells)
    out_index = torch.zeros(out_entries, dtype=torch.int64, device=dev)
    summary = torch.zeros(int(lib.btc_index_summary_words(out_entries)), dtype=torch.int32, device=dev)
    # static mode: dead coordinate rows stay zero so that torch gathers through them remain in bounds
    out_coords = (torch.zeros if static else torch.empty)((cap, 4), dtype=torch.int32, device=dev)
    nbr_out = torch.empty((cap, K), dtype=torch.int32, device=dev)
    nbr_in = torch.empty((max(n_in, 1), K), dtype=torch.int32, device=dev)
    n_out_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.btc_rulebook_conv_sparse_workspace_bytes(out_entries))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    # sparse two-level build (freshly zeroed bitmaps; the rank bitmap stays populated for sub-manifold layers on this level)
    check(lib.btc_rulebook_conv_sparse(_ptr(coords), n_in, _ptr(n_dev), int(batch), int3(in_shape), int3(out_shape), int3(ksize),
                                       int3(stride), int3(padding), int3(dilation), int(bool(transposed)), _ptr(out_index),
                                       out_entries, _ptr(summary), _ptr(out_coords), cap, _ptr(n_out_dev), _ptr(nbr_out),
                                       _ptr(nbr_in), _ptr(ws), ws_bytes, _stream()), "btc_rulebook_conv_sparse")
    idx = CoordIndex(out_index, None, int(batch), out_shape)
    if static:
        _register_cap(n_out_dev, cap, "rulebook %s -> %s" % (in_shape, out_shape))
        return Rulebook(nbr_out, nbr_in, out_coords, n_in, cap, K, False, in_shape, out_shape, ksize, stride, padding,
                        dilation, bool(transposed), idx, n_in_dev=n_dev, n_out_dev=n_out_dev)
    n_out = int(n_out_dev.item())  # the one host sync of a new rulebook (exact tensor shapes for torch)
    if n_out > cap:
        raise _lib.BtcError("rulebook capacity exceeded: %d output sites > capacity %d" % (n_out, cap))
    return Rulebook(nbr_out[:n_out], nbr_in[:n_in], out_coords[:n_out], n_in, n_out, K, False, in_shape, out_shape,
                    ksize, stride, padding, dilation, bool(transposed), idx)


def rulebook_pairs(table: torch.Tensor, n_in: int, K: int, mirror: bool):
    lib = _lib.load()
    dev = table.device
    table = table.contiguous()
    pairs = torch.empty((2, K, max(n_in, 1)), dtype=torch.int32, device=dev)
    pair_num = torch.empty(K, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.btc_rulebook_pairs_workspace_bytes(n_in, K))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.btc_rulebook_pairs(_ptr(table), n_in, None, K, int(mirror), _ptr(pairs), _ptr(pair_num), _ptr(ws),
                                 ws_bytes, _stream()), "btc_rulebook_pairs")
    return pairs[:, :, :n_in], pair_num


# ------------------------------------------------------------------------------------------
# arithmetic
# ------------------------------------------------------------------------------------------
def sparse_conv_fwd(features, nbr_out, weight, bias=None, scale=None, shift=None, relu=False, algo=0,
                    n_out_dev=None, out=None):
    """features [N_in,Cin], nbr_out [N_out,K], weight [K,Cin,Cout] (any [...,Cin,Cout] view) -> [N_out,Cout]."""
    _require_cuda(features, nbr_out, weight)
    lib = _lib.load()
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    n_out, K = nbr_out.shape
    assert features.dtype == torch.float32 and weight.dtype == torch.float32
    assert features.shape[1] == c_in and weight.numel() == K * c_in * c_out
    features, weight, nbr_out = features.contiguous(), weight.contiguous(), nbr_out.contiguous()
    if out is None:
        out = torch.empty((n_out, c_out), dtype=torch.float32, device=features.device)
    check(lib.btc_sparse_conv_fwd(_ptr(features), _ptr(nbr_out), _ptr(weight), _ptr(bias), _ptr(scale), _ptr(shift),
                                  int(bool(relu)), _ptr(out), n_out, _ptr(n_out_dev), K, c_in, c_out, int(algo),
                                  _stream()), "btc_sparse_conv_fwd")
    return out


def tc_supported(K, c_in, c_out):
    return bool(_lib.load().btc_sparse_conv_tc_supported(int(K), int(c_in), int(c_out)))


def tc_config(producer_warps=-1, concat_b=-1, dynamic_tiles=-1):
    """Process-wide variant knobs of the tcgen05 tile (A/B measurements, tests); -1 keeps a setting.
    Defaults: 16 producer warps, dynamic tile scheduling.  concat_b must stay 0 / -1 (variant removed, no gain measured)."""
    check(_lib.load().btc_sparse_conv_tc_config(int(producer_warps), int(concat_b), int(dynamic_tiles)),
          "btc_sparse_conv_tc_config")


def tc_split_supported(K, c_in, c_out, in_split, out_split):
    return bool(_lib.load().btc_sparse_conv_tc_split_supported(int(K), int(c_in), int(c_out), int(bool(in_split)),
                                                               int(bool(out_split))))


def tc_pack_weight_split(weight):
    """Pack [K,Cin,Cout] fp32 weights for a SPLIT-input layer (bf16 hi / lo tiles per 32-element reduction chunk, 64-byte swizzled rows)."""
    lib = _lib.load()
    _require_cuda(weight)
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    K = weight.numel() // (c_in * c_out)
    nbytes = int(lib.btc_sparse_conv_tc_split_packed_bytes(K, c_in, c_out))
    if nbytes < 0:
        raise _lib.BtcError("split-format tile does not support K=%d Cin=%d Cout=%d" % (K, c_in, c_out))
    weight = weight.detach().to(torch.float32).contiguous()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    check(lib.btc_sparse_conv_tc_pack_split(_ptr(weight), K, c_in, c_out, _ptr(packed), _stream()), "btc_sparse_conv_tc_pack_split")
    return packed


def features_to_split(features, n_dev=None):
    """fp32 rows [N, C] (C % 32 == 0) -> split rows (same shape / bytes, float32-typed storage of packed bf16 hi / lo)."""
    _require_cuda(features)
    features = features.to(torch.float32).contiguous()
    out = torch.empty_like(features)
    check(_lib.load().btc_features_to_split(_ptr(features), features.shape[0], _ptr(n_dev), features.shape[1], _ptr(out),
                                            _stream()), "btc_features_to_split")
    return out


def features_from_split(split, n_dev=None):
    _require_cuda(split)
    split = split.contiguous()
    out = torch.empty_like(split)
    check(_lib.load().btc_features_from_split(_ptr(split), split.shape[0], _ptr(n_dev), split.shape[1], _ptr(out), _stream()),
          "btc_features_from_split")
    return out


def sparse_conv_fwd_tc_split(features, nbr_out, packed_weight, c_in, c_out, in_split, out_split, bias=None, scale=None,
                             shift=None, relu=False, n_out_dev=None, out=None, tile_mask=None, tile_order=None):
    """tcgen05 gather-GEMM with the split (bf16 hi / lo) feature format on the input and / or output side.
    packed_weight: tc_pack_weight_split(w) when in_split else tc_pack_weight(w)."""
    _require_cuda(features, nbr_out, packed_weight)
    lib = _lib.load()
    n_out, K = nbr_out.shape
    features, nbr_out = features.contiguous(), nbr_out.contiguous()
    assert features.dtype == torch.float32 and features.shape[1] == c_in
    if out is None:
        out = torch.empty((n_out, c_out), dtype=torch.float32, device=features.device)
    check(lib.btc_sparse_conv_fwd_tc_split(_ptr(features), _ptr(nbr_out), _ptr(packed_weight), _ptr(bias), _ptr(scale),
                                           _ptr(shift), int(bool(relu)), _ptr(out), n_out, _ptr(n_out_dev), K, int(c_in),
                                           int(c_out), int(bool(in_split)), int(bool(out_split)), _ptr(tile_mask),
                                           _ptr(tile_order), _stream()), "btc_sparse_conv_fwd_tc_split")
    return out


def tc_pack_weight(weight):
    """Pack [K,Cin,Cout] (or [*k,Cin,Cout]) fp32 weights into the tcgen05 operand image (hi/lo tf32 split,
    K-major, 128-byte swizzle) consumed by sparse_conv_fwd_tc."""
    lib = _lib.load()
    _require_cuda(weight)
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    K = weight.numel() // (c_in * c_out)
    nbytes = int(lib.btc_sparse_conv_tc_packed_bytes(K, c_in, c_out))
    if nbytes < 0:
        raise _lib.BtcError("tensor-core tile does not support K=%d Cin=%d Cout=%d" % (K, c_in, c_out))
    weight = weight.detach().to(torch.float32).contiguous()
    packed = torch.empty(nbytes, dtype=torch.uint8, device=weight.device)
    check(lib.btc_sparse_conv_tc_pack(_ptr(weight), K, c_in, c_out, _ptr(packed), _stream()), "btc_sparse_conv_tc_pack")
    return packed


def rulebook_tile_meta(nbr_out, n_out_dev=None):
    """btc_rulebook_tile_meta: per 128-row tile the mask of offsets with at least one valid neighbour, and the tiles
    bucketed by cost class (heaviest-first hand-out order of the tcgen05 tile)."""
    _require_cuda(nbr_out)
    lib = _lib.load()
    nbr_out = nbr_out.contiguous()
    n_out, K = nbr_out.shape
    tiles = (n_out + 127) // 128
    tile_mask = torch.empty(max(tiles, 1), dtype=torch.int64, device=nbr_out.device)
    tile_order = torch.empty(int(lib.btc_rulebook_tile_order_ints(n_out)), dtype=torch.int32, device=nbr_out.device)
    check(lib.btc_rulebook_tile_meta(_ptr(nbr_out), n_out, _ptr(n_out_dev), K, _ptr(tile_mask), _ptr(tile_order), _stream()),
          "btc_rulebook_tile_meta")
    return tile_mask, tile_order


def rulebook_sort_rows(nbr_out, n_out_dev=None, out=None):
    """btc_rulebook_sort_rows: rows of nbr_out reordered by valid-offset mask within 2048-row windows.
    Returns (nbr_sorted [N_out, K], out_rows [N_out]); row i of nbr_sorted is row out_rows[i] of nbr_out."""
    _require_cuda(nbr_out)
    lib = _lib.load()
    nbr_out = nbr_out.contiguous()
    n_out, K = nbr_out.shape
    nbr_sorted, out_rows = out if out is not None else (torch.empty_like(nbr_out),
                                                        torch.empty(n_out, dtype=torch.int32, device=nbr_out.device))
    check(lib.btc_rulebook_sort_rows(_ptr(nbr_out), n_out, _ptr(n_out_dev), K, _ptr(nbr_sorted), _ptr(out_rows), _stream()),
          "btc_rulebook_sort_rows")
    return nbr_sorted, out_rows


def sparse_conv_fwd_tc(features, nbr_out, packed_weight, c_in, c_out, bias=None, scale=None, shift=None, relu=False,
                       n_out_dev=None, out=None, out_rows=None):
    """tcgen05 3xTF32 gather-GEMM; same contract as sparse_conv_fwd with pre-packed weights.  With `out_rows`,
    `nbr_out` is a mask-sorted table (rulebook_sort_rows) and row i's result goes to out[out_rows[i]]."""
    _require_cuda(features, nbr_out, packed_weight)
    lib = _lib.load()
    n_out, K = nbr_out.shape
    features, nbr_out = features.contiguous(), nbr_out.contiguous()
    assert features.dtype == torch.float32 and features.shape[1] == c_in
    if out is None:
        out = torch.empty((n_out, c_out), dtype=torch.float32, device=features.device)
    if out_rows is not None:
        check(lib.btc_sparse_conv_fwd_tc_rows(_ptr(features), _ptr(nbr_out), _ptr(out_rows), _ptr(packed_weight), _ptr(bias),
                                              _ptr(scale), _ptr(shift), int(bool(relu)), _ptr(out), n_out, _ptr(n_out_dev),
                                              K, int(c_in), int(c_out), _stream()), "btc_sparse_conv_fwd_tc_rows")
        return out
    check(lib.btc_sparse_conv_fwd_tc(_ptr(features), _ptr(nbr_out), _ptr(packed_weight), _ptr(bias), _ptr(scale),
                                     _ptr(shift), int(bool(relu)), _ptr(out), n_out, _ptr(n_out_dev), K, int(c_in),
                                     int(c_out), _stream()), "btc_sparse_conv_fwd_tc")
    return out


def sparse_conv_bwd_data(d_out, table, mirror, weight, n_in):
    lib = _lib.load()
    c_in, c_out = weight.shape[-2], weight.shape[-1]
    K = table.shape[1]
    d_out, table, weight = d_out.contiguous(), table.contiguous(), weight.contiguous()
    d_in = torch.empty((n_in, c_in), dtype=torch.float32, device=d_out.device)
    ws_bytes = int(lib.btc_sparse_conv_bwd_workspace_bytes(K, c_in, c_out))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d_out.device)
    check(lib.btc_sparse_conv_bwd_data(_ptr(d_out), _ptr(table), int(bool(mirror)), _ptr(weight), _ptr(d_in), n_in,
                                       None, K, c_in, c_out, _ptr(ws), ws_bytes, _stream()),
          "btc_sparse_conv_bwd_data")
    return d_in


def sparse_conv_bwd_weight(features, d_out, nbr_out, weight_shape, want_bias):
    lib = _lib.load()
    c_in, c_out = weight_shape[-2], weight_shape[-1]
    n_out, K = nbr_out.shape
    features, d_out, nbr_out = features.contiguous(), d_out.contiguous(), nbr_out.contiguous()
    d_w = torch.empty(weight_shape, dtype=torch.float32, device=d_out.device)
    d_b = torch.empty(c_out, dtype=torch.float32, device=d_out.device) if want_bias else None
    check(lib.btc_sparse_conv_bwd_weight(_ptr(features), _ptr(d_out), _ptr(nbr_out), _ptr(d_w), _ptr(d_b), n_out,
                                         None, K, c_in, c_out, _stream()), "btc_sparse_conv_bwd_weight")
    return d_w, d_b


def _packed_weight(weight):
    """tcgen05 operand image of a layer's weights, re-packed on every eager forward: the pack kernel touches only
    K*Cin*Cout elements, and any cache keyed on (data_ptr, _version) goes stale under `weight.data.fill_()` (the
    reference's own idiom, spconv_backbone.py:48) or when the allocator recycles a freed Parameter's storage."""
    return tc_pack_weight(weight)


class SparseConvFunction(torch.autograd.Function):
    """indice_conv of spconv.ops with autograd (forward + dX + dW + db on the CUDA library).

    algo: 0 = auto (tcgen05 3xTF32 tile when the shape qualifies, fp32 FFMA tile otherwise), 1 = FFMA, 2 = tcgen05.
    Backward: dX is the same gather-GEMM over the transposed relation — the sub-manifold table read with the mirrored
    offsets W'[k] = W[K-1-k]^T, or nbr_in with W[k]^T for strided / transposed layers — and runs on the tcgen05 tile
    whenever that shape qualifies (algo != 1); dW / db are fp32 outer products with atomics."""

    @staticmethod
    def forward(ctx, features, weight, bias, rulebook: Rulebook, algo, epilogue=None):
        c_in, c_out = weight.shape[-2], weight.shape[-1]
        scale, shift, relu = epilogue if epilogue is not None else (None, None, False)
        if epilogue is not None and c_in % 4 != 0 and algo != 1:
            # inference fast path: zero-pad the input channels to a multiple of 4 (16-byte gather pieces) so that thin odd
            # layers (the detection backbone's 6 -> 16 input convolution) run on the tcgen05 tile as well
            pad = 4 - c_in % 4
            features = torch.nn.functional.pad(features, (0, pad))
            weight = torch.nn.functional.pad(weight, (0, 0, 0, pad))
            c_in += pad
        if epilogue is not None and (features.requires_grad or weight.requires_grad) and torch.is_grad_enabled():
            raise _lib.BtcError("the fused conv epilogue is inference-only (no backward through the folded BatchNorm)")
        use_tc = algo != 1 and tc_supported(rulebook.K, c_in, c_out) and features.shape[0] > 0 \
            and features.data_ptr() % 16 == 0
        if algo == 2 and not use_tc:
            raise _lib.BtcError("tensor-core tile requested for an unsupported shape (K=%d Cin=%d Cout=%d)" %
                                (rulebook.K, c_in, c_out))
        if use_tc:
            # per-tile offset masks + heaviest-first tile order, computed once per rulebook (one launch; the mask sort this
            # replaced cost 18 launches / 1.26 ms of a 5.5 ms config-3 step and saved less than it cost)
            # (thin sub-manifold layers run every chunk: the mask launch does not pay there)
            tile_mask, tile_order = rulebook.tile_meta() if (c_in > 16 or not rulebook.subm) else (None, None)
            out = sparse_conv_fwd_tc_split(features.contiguous(), rulebook.nbr_out, _packed_weight(weight), c_in, c_out,
                                           False, False, bias=bias, scale=scale, shift=shift, relu=relu,
                                           n_out_dev=rulebook.n_out_dev, tile_mask=tile_mask, tile_order=tile_order)
        else:
            out = sparse_conv_fwd(features, rulebook.nbr_out, weight, bias, scale, shift, relu, algo=1,
                                  n_out_dev=rulebook.n_out_dev)
        ctx.save_for_backward(features, weight)
        ctx.rulebook = rulebook
        ctx.has_bias = bias is not None
        ctx.algo = algo
        return out

    @staticmethod
    def backward(ctx, d_out):
        features, weight = ctx.saved_tensors
        rb = ctx.rulebook
        d_out = d_out.contiguous()
        d_feat = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            c_in, c_out = weight.shape[-2], weight.shape[-1]
            n_in = features.shape[0]
            if ctx.algo != 1 and n_in > 0 and d_out.shape[0] > 0 and tc_supported(rb.K, c_out, c_in) \
                    and d_out.data_ptr() % 16 == 0:
                wk = weight.detach().reshape(rb.K, c_in, c_out)
                if rb.subm:     # nbr_out[o][k] = i  <=>  nbr_out[i][K-1-k] = o
                    table, wb = rb.nbr_out, wk.flip(0).transpose(1, 2).contiguous()
                else:
                    table, wb = rb.nbr_in, wk.transpose(1, 2).contiguous()
                d_feat = sparse_conv_fwd_tc(d_out, table, tc_pack_weight(wb), c_out, c_in)
            else:
                table, mirror = (rb.nbr_out, True) if rb.subm else (rb.nbr_in, False)
                d_feat = sparse_conv_bwd_data(d_out, table, mirror, weight, n_in)
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            d_w, d_b = sparse_conv_bwd_weight(features, d_out, rb.nbr_out, tuple(weight.shape), ctx.has_bias)
        return d_feat, d_w, d_b, None, None, None


def maxpool_fwd(features, nbr_out, n_out_dev=None):
    lib = _lib.load()
    _require_cuda(features, nbr_out)
    n_out, K = nbr_out.shape
    c = features.shape[1]
    features, nbr_out = features.contiguous(), nbr_out.contiguous()
    out = torch.empty((n_out, c), dtype=torch.float32, device=features.device)
    check(lib.btc_maxpool_fwd(_ptr(features), _ptr(nbr_out), _ptr(out), n_out, _ptr(n_out_dev), K, c, _stream()),
          "btc_maxpool_fwd")
    return out


class SparseMaxPoolFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, rulebook: Rulebook):
        out = maxpool_fwd(features, rulebook.nbr_out, rulebook.n_out_dev)
        ctx.save_for_backward(features, out)
        ctx.rulebook = rulebook
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        features, out = ctx.saved_tensors
        rb = ctx.rulebook
        d_out = d_out.contiguous()
        n_in, c = features.shape
        d_in = torch.empty_like(features)
        check(lib.btc_maxpool_bwd(_ptr(features), _ptr(out), _ptr(d_out), _ptr(rb.nbr_out.contiguous()), _ptr(d_in),
                                  n_in, rb.nbr_out.shape[0], None, rb.K, c, _stream()), "btc_maxpool_bwd")
        return d_in, None


class ToDenseFunction(torch.autograd.Function):
    """SparseConvTensor.dense(): [N,C] rows -> [B,C,D,H,W] (channels first)."""

    @staticmethod
    def forward(ctx, features, coords, batch, shape, n_dev=None):
        lib = _lib.load()
        _require_cuda(features, coords)
        features, coords = features.contiguous(), coords.contiguous()
        n, c = features.shape
        out = torch.empty((batch, c, shape[0], shape[1], shape[2]), dtype=torch.float32, device=features.device)
        check(lib.btc_to_dense(_ptr(features), _ptr(coords), n, _ptr(n_dev), c, int(batch), int3(shape), _ptr(out), _stream()),
              "btc_to_dense")
        ctx.save_for_backward(coords)
        ctx.geom = (n, c, int(batch), list(shape))
        return out

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        (coords,) = ctx.saved_tensors
        n, c, batch, shape = ctx.geom
        d_out = d_out.contiguous()
        d_feat = torch.empty((n, c), dtype=torch.float32, device=d_out.device)
        check(lib.btc_from_dense(_ptr(d_out), _ptr(coords), n, None, c, batch, int3(shape), _ptr(d_feat), _stream()),
              "btc_from_dense")
        return d_feat, None, None, None, None


# ------------------------------------------------------------------------------------------
# voxelisation
# ------------------------------------------------------------------------------------------
def voxelize(points: torch.Tensor, scene_offsets: torch.Tensor, voxel_size, coors_range, max_points: int,
             max_voxels: int, want_mean=False, grid=None):
    """GPU VoxelGeneratorV2.generate over a batch of scenes.

    points [N, C] f32 cuda (all scenes concatenated), scene_offsets [B+1] i32 cuda.
    Returns (voxels [cap,P,C], coords [cap,4] (b,z,y,x), num_points [cap], mean or None, n_voxels [B+1] device).
    Rows beyond n_voxels[B] are unspecified; no host sync is performed.
    """
    _require_cuda(points, scene_offsets)
    lib = _lib.load()
    assert points.dtype == torch.float32 and points.dim() == 2 and points.is_contiguous()
    assert scene_offsets.dtype == torch.int32
    n, c = points.shape
    n_scenes = scene_offsets.numel() - 1
    if grid is None:
        grid = voxel_grid_size(voxel_size, coors_range)
    dev = points.device
    cap = n_scenes * max_voxels
    voxels = torch.empty((cap, max_points, c), dtype=torch.float32, device=dev)
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num_points = torch.empty(cap, dtype=torch.int32, device=dev)
    mean = torch.empty((cap, c), dtype=torch.float32, device=dev) if want_mean else None
    n_voxels = torch.empty(n_scenes + 1, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.btc_voxelize_workspace_bytes(n, n_scenes, max_voxels, max_points))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.btc_voxelize(_ptr(points), n, c, _ptr(scene_offsets), n_scenes, float_array(voxel_size),
                           float_array(coors_range), int3(grid), int(max_points), int(max_voxels), _ptr(voxels),
                           _ptr(coords), _ptr(num_points), _ptr(mean), _ptr(n_voxels), _ptr(ws), ws_bytes, _stream()),
          "btc_voxelize")
    return voxels, coords, num_points, mean, n_voxels


def points_to_cylinder(points: torch.Tensor, sphere=False, n_dev=None, out=None):
    """absxyz_2_cylinxyz_np / absxyz_2_spherexyz_np (btcdet/utils/coords_utils.py:268-292) on device points [N, C>=3]:
    (rho, phi_deg, z, extra...) or (r, az_deg, el_deg, extra...), numpy's float32 op order."""
    _require_cuda(points)
    assert points.dtype == torch.float32 and points.dim() == 2 and points.is_contiguous()
    n, c = points.shape
    out = torch.empty_like(points) if out is None else out
    check(_lib.load().btc_points_to_cylinder(_ptr(points), n, _ptr(n_dev), c, int(bool(sphere)), _ptr(out), _stream()),
          "btc_points_to_cylinder")
    return out


def voxelize_occ_and_det(points, scene_offsets, occ_voxel_size, occ_range, occ_max_points, occ_max_voxels,
                         det_voxel_size, det_range, det_max_points, det_max_voxels, want_mean=True):
    """GPU-resident input pipeline: raw points feed BOTH voxelisers of data_processor.py:105-190 without
    leaving the device — the cylindrical occupancy voxeliser (a2 transform + VoxelGeneratorV2, :128-136) and the Cartesian
    detection voxeliser (:176-177).  Returns (occ, det), each the tuple voxelize() returns."""
    cyl = points_to_cylinder(points)
    occ = voxelize(cyl, scene_offsets, occ_voxel_size, occ_range, occ_max_points, occ_max_voxels, want_mean=want_mean)
    det = voxelize(points, scene_offsets, det_voxel_size, det_range, det_max_points, det_max_voxels, want_mean=False)
    return occ, det


def voxel_grid_size(voxel_size, coors_range):
    """grid = round((max - min) / voxel_size) with float32 operands, as VoxelGeneratorV2.__init__ computes it."""
    import numpy as np
    r = np.array(coors_range, dtype=np.float32)
    v = np.array(voxel_size, dtype=np.float32)
    return [int(g) for g in np.round((r[3:] - r[:3]) / v).astype(np.int64)]


def revoxelize_sorted(pt_coords: torch.Tensor, pt_feat: torch.Tensor, batch: int, shape):
    """torch.unique(coords, dim=0, sorted)+pad of add_occ_template.py:248-268 on the rank bitmap.

    Returns (voxels [M, Pmax, C], voxel_num_points [M] i64, voxel_coords [M,4] i64) like the reference.
    """
    _require_cuda(pt_coords, pt_feat)
    lib = _lib.load()
    dev = pt_coords.device
    pc = pt_coords.to(torch.int32).contiguous()
    pt_feat = pt_feat.to(torch.float32).contiguous()
    n, c = pt_feat.shape
    n_entries = index_entries(batch, shape)
    index = torch.zeros(n_entries, dtype=torch.int64, device=dev)
    vox_coords = torch.empty((max(n, 1), 4), dtype=torch.int32, device=dev)
    vox_count = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    slots = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    pt_voxel = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.btc_revoxelize_workspace_bytes(n, n_entries))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.btc_revoxelize(_ptr(pc), n, None, int(batch), int3(shape), _ptr(index), n_entries, _ptr(vox_coords), n,
                             _ptr(vox_count), _ptr(slots), _ptr(pt_voxel), _ptr(counts[0:1]), _ptr(counts[1:2]),
                             _ptr(ws), ws_bytes, _stream()), "btc_revoxelize")
    m, pmax = [int(v) for v in counts.tolist()]  # one host read: exact shapes for the torch modules downstream
    voxels = torch.empty((m, pmax, c), dtype=torch.float32, device=dev)
    check(lib.btc_revoxelize_fill(_ptr(pt_feat), _ptr(pt_voxel), _ptr(slots), n, None, c, pmax, _ptr(voxels), m,
                                  _stream()), "btc_revoxelize_fill")
    return voxels, vox_count[:m].to(torch.int64), vox_coords[:m].to(torch.int64)


# ------------------------------------------------------------------------------------------
# occupancy / occlusion masks
# ------------------------------------------------------------------------------------------
def occ_geometry_arrays(voxel_size, point_cloud_range, support_sphere_range, dist_kern, half_x, empt_sur_thresh,
                        det_point_cloud_range):
    """Pack DATA_CONFIG.OCC geometry (btcdet_kitti_car.yaml:72-92) into the geom_f / geom_i arrays of btc_occ_targets."""
    import numpy as np
    r, v = np.array(point_cloud_range, dtype=np.float64), np.array(voxel_size, dtype=np.float64)
    grid = np.round((r[3:6] - r[0:3]) / v).astype(np.int64)                       # data_processor.py:119-120
    sr = np.asarray(support_sphere_range, dtype=np.float64)
    svs = np.array([voxel_size[0], voxel_size[1], sr[6]], dtype=np.float64)        # occ_targets_template.py:48
    sgrid = ((sr[3:6] - sr[:3]) / svs).astype(int)                                 # :51 (truncation)
    use_empty = empt_sur_thresh != "None" and float(empt_sur_thresh) < 9
    gf = list(voxel_size) + list(r[:3]) + list(r[3:6]) + list(svs) + list(sr[:3]) + list(sr[3:6]) + \
        [float(empt_sur_thresh) if use_empty else 0.0, float(det_point_cloud_range[2]), float(det_point_cloud_range[5])]
    gi = [int(g) for g in grid] + [int(g) for g in sgrid] + [int(k) for k in dist_kern] + \
        [int(dist_kern[-1] // 2 if half_x else 0), int(use_empty)]
    return gf, gi


def occ_targets(voxels, voxel_coords, voxel_num_points, batch_size, geom_f, geom_i, rot_z=None, want_sphere=False,
                n_dev=None):
    """Fused GPU occupancy / occlusion masks.  Returns a dict of uint8 [B,nz,ny,nx] tensors."""
    _require_cuda(voxels, voxel_coords, voxel_num_points)
    lib = _lib.load()
    dev = voxels.device
    voxels = voxels.to(torch.float32).contiguous()
    coords = voxel_coords.to(torch.int32).contiguous()
    nump = voxel_num_points.to(torch.int32).contiguous()
    m, P, C = voxels.shape
    B = int(batch_size)
    gf, gi = float_array(geom_f), (ctypes.c_int * len(geom_i))(*[int(v) for v in geom_i])
    nx, ny, nz = geom_i[0:3]
    shape = (B, nz, ny, nx)
    out = {k: torch.empty(shape, dtype=torch.uint8, device=dev)
           for k in ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask")}
    sphere = torch.empty((B, geom_i[5], geom_i[4], geom_i[3]), dtype=torch.uint8, device=dev) if want_sphere else None
    ws_bytes = int(lib.btc_occ_targets_workspace_bytes(B, gf, gi))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rz = None if rot_z is None else rot_z.to(torch.float32).contiguous()
    check(lib.btc_occ_targets(_ptr(voxels), P, C, _ptr(coords), _ptr(nump), m, _ptr(n_dev), B, _ptr(rz), gf, gi,
                              _ptr(out["voxelwise_mask"]), _ptr(out["vcc_mask"]), _ptr(out["occ_voxelwise_mask"]),
                              _ptr(out["general_cls_loss_mask"]), _ptr(sphere), _ptr(ws), ws_bytes, _stream()),
          "btc_occ_targets")
    if want_sphere:
        out["sphere_map"] = sphere
    return out


DEFAULT_LOSS_WEIGHTS = {"occ_fore_cls_weight": 1.0, "occ_mirr_cls_weight": 1.0, "occ_bm_cls_weight": 1.0,
                        "occ_neg_cls_weight": 1.0, "occ_fore_res_weight": 0.1, "occ_mirr_res_weight": 0.0,
                        "occ_bm_res_weight": 0.0}


_CENTERS2D = {}


def occ_voxel_centers_2d(geom_f, geom_i, device):
    """The reference's `voxel_centers["all_voxel_centers_2d"]` (detector3d_template.py:52-63): cylinder voxel centres ->
    Cartesian -> mean over z of (x, y), flattened [ny * nx, 2].  Computed with the same torch ops as the reference at model
    build (so every rounding, including the 9-term mean, is torch's) and cached per geometry and device."""
    key = (tuple(float(v) for v in geom_f[:6]), tuple(int(v) for v in geom_i[:3]), str(device))
    hit = _CENTERS2D.get(key)
    if hit is None:
        import numpy as np
        nx, ny, nz = (int(v) for v in geom_i[:3])
        vs = torch.tensor([float(geom_f[2]), float(geom_f[1]), float(geom_f[0])], dtype=torch.float32, device=device)
        org = torch.tensor([float(geom_f[5]), float(geom_f[4]), float(geom_f[3])], dtype=torch.float32, device=device)
        # ^ get_all_voxel_centers_zyx (coords_utils.py:166-178): float32 tensors built from python floats
        z, y, x = torch.meshgrid(torch.arange(nz, device=device), torch.arange(ny, device=device),
                                 torch.arange(nx, device=device), indexing="ij")
        c = (0.5 + torch.stack([z, y, x], dim=0).to(torch.float32)) * vs.view(3, 1, 1, 1) + org.view(3, 1, 1, 1)
        rho, phi = c[2], c[1]                                                          # cylinder_uvd2absxyz (:198-204)
        cx = rho * torch.cos(phi * np.pi / 180.)
        cy = -rho * torch.sin(phi * np.pi / 180.)
        centers = torch.stack([cx, cy, c[0]], dim=-1)          # [nz, ny, nx, 3], the layout the reference reduces over
        hit = torch.mean(centers[:, :, :, :2], dim=0).view(-1, 2).contiguous()
        assert hit.dtype == torch.float32
        _CENTERS2D[key] = hit
    return hit


def occ_box_targets(voxels, voxel_coords, voxel_num_points, batch_size, gt_boxes, gt_boxes_num, geom_f, geom_i,
                    box_mirr_flag=None, bm_points=None, rot_z=None, num_class=1, want_forebox=True, want_point_label=False,
                    mirr_cap=None, bm_cap=None):
    """Foreground / mirrored / best-match masks with their mean-residual volumes and the forebox label
    (occ_targets_3d.py:70-86,95-171) in one sync-free call.
    gt_boxes [B, M, >=8] f32, gt_boxes_num list or int tensor [B], box_mirr_flag [B, M], bm_points [n, 4] (b,x,y,z).
    Returns a dict of dense tensors; "status" is a device int32 (1 = an accumulator capacity overflowed)."""
    _require_cuda(voxels, voxel_coords, voxel_num_points, gt_boxes)
    lib = _lib.load()
    dev = voxels.device
    voxels = voxels.to(torch.float32).contiguous()
    coords = voxel_coords.to(torch.int32).contiguous()
    nump = voxel_num_points.to(torch.int32).contiguous()
    m, P, C = voxels.shape
    B = int(batch_size)
    boxes = gt_boxes.to(torch.float32).contiguous()
    max_boxes, box_dim = int(boxes.shape[1]), int(boxes.shape[2])
    bnum = torch.as_tensor(gt_boxes_num, dtype=torch.int32).to(dev).contiguous()
    flag = None if box_mirr_flag is None else box_mirr_flag.to(device=dev, dtype=torch.float32).contiguous()
    bm = None if bm_points is None or len(bm_points) == 0 else bm_points.to(device=dev, dtype=torch.float32).contiguous()
    n_bm = 0 if bm is None else int(bm.shape[0])
    gf, gi = float_array(geom_f), (ctypes.c_int * len(geom_i))(*[int(v) for v in geom_i])
    nx, ny, nz = geom_i[0:3]
    cells = B * nx * ny * nz
    mirr_cap = int(mirr_cap if mirr_cap is not None else min(cells, max(1024, m * P)))
    bm_cap = int(bm_cap if bm_cap is not None else min(cells, max(1024, n_bm)))
    shape = (B, nz, ny, nx)
    out = {"fore_voxelwise_mask": torch.empty(shape, dtype=torch.uint8, device=dev),
           "mirr_fore_voxelwise_mask": torch.empty(shape, dtype=torch.uint8, device=dev),
           "fore_res_mtrx": torch.empty((B, 3, nz, ny, nx), dtype=torch.float32, device=dev),
           "mirr_res_mtrx": torch.empty((B, 3, nz, ny, nx), dtype=torch.float32, device=dev),
           "bm_voxelwise_mask": torch.empty(shape, dtype=torch.uint8, device=dev) if n_bm else None,
           "bm_res_mtrx": torch.empty((B, 3, nz, ny, nx), dtype=torch.float32, device=dev) if n_bm else None,
           "forebox_label": torch.empty(shape, dtype=torch.int8, device=dev) if want_forebox else None,
           "point_label": torch.empty((m, P), dtype=torch.int8, device=dev) if want_point_label else None,
           "status": torch.empty(1, dtype=torch.int32, device=dev)}
    ws_bytes = int(lib.btc_occ_box_targets_workspace_bytes(B, max_boxes, mirr_cap, bm_cap))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    rz = None if rot_z is None else rot_z.to(torch.float32).contiguous()
    c2d = occ_voxel_centers_2d(geom_f, geom_i, dev) if want_forebox else None
    check(lib.btc_occ_box_targets_v2(_ptr(voxels), P, C, _ptr(coords), _ptr(nump), m, None, B, _ptr(boxes), max_boxes, box_dim,
                                  _ptr(bnum), _ptr(flag), _ptr(bm), n_bm, _ptr(rz), gf, gi, int(num_class), mirr_cap, bm_cap,
                                  _ptr(c2d), _ptr(out["fore_voxelwise_mask"]), _ptr(out["fore_res_mtrx"]),
                                  _ptr(out["mirr_fore_voxelwise_mask"]), _ptr(out["mirr_res_mtrx"]),
                                  _ptr(out["bm_voxelwise_mask"]), _ptr(out["bm_res_mtrx"]), _ptr(out["forebox_label"]),
                                  _ptr(out["point_label"]), _ptr(out["status"]), _ptr(ws), ws_bytes, _stream()),
          "btc_occ_box_targets")
    return out


def occ_loss_maps(occ, box, weights=None, box_weight=0.2):
    """prepare_cls_loss_map + prepare_reg_loss_map (occ_targets_template.py:330-401, dropout off) on the outputs of
    occ_targets() and occ_box_targets(): every derived mask, the float weight maps and res_mtrx in one pass."""
    lib = _lib.load()
    w = dict(DEFAULT_LOSS_WEIGHTS)
    w.update(weights or {})
    vm = occ["voxelwise_mask"]
    dev = vm.device
    B, nz, ny, nx = vm.shape
    forebox = box.get("forebox_label") if box_weight != 1.0 else None
    wf = float_array([w["occ_fore_cls_weight"], w["occ_mirr_cls_weight"], w["occ_bm_cls_weight"], w["occ_neg_cls_weight"],
                      w["occ_fore_res_weight"], w["occ_mirr_res_weight"], w["occ_bm_res_weight"],
                      float(box_weight) - w["occ_neg_cls_weight"]])
    u8 = lambda: torch.empty((B, nz, ny, nx), dtype=torch.uint8, device=dev)   # noqa: E731
    f32 = lambda: torch.empty((B, nz, ny, nx), dtype=torch.float32, device=dev)   # noqa: E731
    out = {"occ_fore_cls_mask": u8(), "occ_mirr_cls_mask": u8(), "occ_bm_cls_mask": u8(), "pos_mask": u8(),
           "bm_voxelwise_mask": u8(), "general_cls_loss_mask_float": f32(), "general_reg_loss_mask": u8(),
           "general_reg_loss_mask_float": f32(), "res_mtrx": torch.empty((B, 3, nz, ny, nx), dtype=torch.float32, device=dev),
           "pos_all_num": torch.empty(1, dtype=torch.int32, device=dev)}
    check(lib.btc_occ_loss_maps(_ptr(vm), _ptr(occ["general_cls_loss_mask"]), _ptr(box["fore_voxelwise_mask"]),
                                _ptr(box["mirr_fore_voxelwise_mask"]), _ptr(box.get("bm_voxelwise_mask")), _ptr(forebox),
                                _ptr(box["fore_res_mtrx"]), _ptr(box["mirr_res_mtrx"]), _ptr(box.get("bm_res_mtrx")), wf, B,
                                int3([nx, ny, nz]), _ptr(out["occ_fore_cls_mask"]), _ptr(out["occ_mirr_cls_mask"]),
                                _ptr(out["occ_bm_cls_mask"]), _ptr(out["pos_mask"]), _ptr(out["bm_voxelwise_mask"]),
                                _ptr(out["general_cls_loss_mask_float"]), _ptr(out["general_reg_loss_mask"]),
                                _ptr(out["general_reg_loss_mask_float"]), _ptr(out["res_mtrx"]), _ptr(out["pos_all_num"]),
                                _stream()), "btc_occ_loss_maps")
    out["forebox_label"] = forebox
    return out


def occ_training_targets(voxels, voxel_coords, voxel_num_points, batch_size, gt_boxes, gt_boxes_num, geom_f, geom_i,
                         box_mirr_flag=None, bm_points=None, rot_z=None, num_class=1, weights=None, box_weight=0.2):
    """OccTargets3D.create_voxel_res_label (occ_targets_3d.py:44-92) end to end on the GPU: three C-ABI calls, no host
    synchronisation.  Returns the reference's batch_dict entries (masks u8, weight maps f32, res_mtrx f32)."""
    occ = occ_targets(voxels, voxel_coords, voxel_num_points, batch_size, geom_f, geom_i, rot_z=rot_z)
    box = occ_box_targets(voxels, voxel_coords, voxel_num_points, batch_size, gt_boxes, gt_boxes_num, geom_f, geom_i,
                          box_mirr_flag=box_mirr_flag, bm_points=bm_points, rot_z=rot_z, num_class=num_class,
                          want_forebox=box_weight != 1.0)
    out = occ_loss_maps(occ, box, weights, box_weight)
    out.update({k: occ[k] for k in ("voxelwise_mask", "vcc_mask", "occ_voxelwise_mask", "general_cls_loss_mask")})
    out["fore_voxelwise_mask"] = box["fore_voxelwise_mask"]
    out["status"] = box["status"]
    return out


# ------------------------------------------------------------------------------------------
# occupancy-point injection (PassOccVox) and OccVFE
# ------------------------------------------------------------------------------------------
def occ_select(probs, residuals, thresh, max_points, occ_voxel_size, occ_origin, det_voxel_size, det_range, det_grid,
               rot_z=None, inten=0.0):
    """filter_occ_points + occ_coords2absxyz + trans_voxel_grid + assemble_occ_points on the GPU
    (btcdet/models/occ_pnt/add_occ_template.py:78-165).  One host read (the counts) for exact shapes; the rare
    top-k branch (more than `max_points` cells above threshold in a scene) is applied with torch.topk on the
    compacted probabilities.  Returns a dict of tensors or None when no cell passes."""
    _require_cuda(probs)
    lib = _lib.load()
    dev = probs.device
    probs = probs.to(torch.float32).contiguous()
    res = None if residuals is None else residuals.to(torch.float32).contiguous()
    B, nz, ny, nx = probs.shape
    grid = int3([nx, ny, nz])
    n_above = B * nz * ny * nx
    cap = min(n_above, max(1024, B * max(int(max_points), 1) * 4))
    ws_bytes = int(lib.btc_occ_select_workspace_bytes(B, grid))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    gf = float_array(list(occ_voxel_size) + list(occ_origin) + list(det_voxel_size) + list(det_range[:3]))
    rz = None if rot_z is None else rot_z.to(torch.float32).contiguous()
    while True:
        out = {"occ_coords": torch.empty((cap, 4), dtype=torch.int32, device=dev),
               "occ_probs": torch.empty(cap, dtype=torch.float32, device=dev),
               "occ_xyz": torch.empty((cap, 3), dtype=torch.float32, device=dev),
               "det_coords": torch.empty((cap, 4), dtype=torch.int32, device=dev),
               "occ_points": torch.empty((cap, 6), dtype=torch.float32, device=dev)}
        counts = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        check(lib.btc_occ_select(_ptr(probs), _ptr(res), B, grid, ctypes.c_float(thresh), _ptr(rz), gf, int3(det_grid),
                                 ctypes.c_float(inten), cap, _ptr(out["occ_coords"]), _ptr(out["occ_probs"]),
                                 _ptr(out["occ_xyz"]), _ptr(out["det_coords"]), _ptr(out["occ_points"]), _ptr(counts),
                                 _ptr(ws), ws_bytes, _stream()), "btc_occ_select")
        c = counts.tolist()
        if c[B] <= cap:
            break
        cap = c[B]                      # capacity guess too small: rerun once with the exact size
    total = c[B]
    if total == 0:
        return None
    out = {k: v[:total] for k, v in out.items()}
    if max(c[:B]) > max_points:         # top-k branch of the reference (order unspecified there: sorted=False)
        keep, start = [], 0
        for b in range(B):
            n_b = c[b]
            if n_b > max_points:
                _, top = torch.topk(out["occ_probs"][start:start + n_b], int(max_points), largest=True, sorted=False)
                keep.append(top + start)
            elif n_b > 0:
                keep.append(torch.arange(start, start + n_b, device=dev))
            start += n_b
        keep = torch.cat(keep)
        out = {k: v[keep] for k, v in out.items()}
    out["counts"] = c
    return out


def occ_head_prob(logits, coords, batch_size, grid_xyz, mask=None, n_dev=None):
    """softmax(dense(logits), dim=1)[:, -1] * mask of OccHead3D.forward (occ_head_3D.py:46-49) from the head's sparse rows:
    logits [N, n_cls], coords [N, 4] (b,z,y,x), grid_xyz = (nx, ny, nz), mask u8 [B,nz,ny,nx] or None -> prob f32 [B,nz,ny,nx]."""
    _require_cuda(logits, coords)
    lib = _lib.load()
    logits = logits.to(torch.float32).contiguous()
    coords = coords.to(torch.int32).contiguous()
    nx, ny, nz = [int(v) for v in grid_xyz]
    prob = torch.empty((int(batch_size), nz, ny, nx), dtype=torch.float32, device=logits.device)
    m = None if mask is None else mask.to(torch.uint8).contiguous()
    check(lib.btc_occ_head_prob(_ptr(logits), _ptr(coords), logits.shape[0], _ptr(n_dev), logits.shape[1], int(batch_size),
                                int3([nx, ny, nz]), _ptr(m), _ptr(prob), _stream()), "btc_occ_head_prob")
    return prob


def occ_vfe(voxels, voxel_num_points, num_raw_features=4, n_dev=None):
    """OccVFE.forward (btcdet/models/backbones_3d/vfe/occ_vfe.py:24-55): returns (voxel_features [M,C], occ_voxel_features)."""
    _require_cuda(voxels, voxel_num_points)
    lib = _lib.load()
    voxels = voxels.to(torch.float32).contiguous()
    nump = voxel_num_points.to(torch.int32).contiguous()
    m, P, C = voxels.shape
    feats = torch.empty((m, C), dtype=torch.float32, device=voxels.device)
    occ = torch.empty((m, C - num_raw_features), dtype=torch.float32, device=voxels.device)
    check(lib.btc_occ_vfe(_ptr(voxels), _ptr(nump), m, _ptr(n_dev), P, C, int(num_raw_features), _ptr(feats), _ptr(occ),
                          _stream()), "btc_occ_vfe")
    return feats, occ


def pass_occ_vox(probs, residuals, det_voxels, det_voxel_num_points, det_voxel_coords, batch_size, thresh, max_points,
                 occ_voxel_size, occ_origin, det_voxel_size, det_range, det_grid, rot_z=None, inten=0.0):
    """PassOccVox.forward (btcdet/models/occ_pnt/pass_occ_vox.py:10-59) for the cylinder / REG configuration:
    select occupancy cells, build pseudo points, append the raw det-voxel points (two zero code channels) and
    re-voxelise everything sorted on the det grid.  Returns (voxels [M',Pmax,6], num_points, coords, selection)."""
    sel = occ_select(probs, residuals, thresh, max_points, occ_voxel_size, occ_origin, det_voxel_size, det_range, det_grid,
                     rot_z=rot_z, inten=inten)
    M, P, C = det_voxels.shape
    mask = torch.arange(P, device=det_voxels.device).view(1, -1) < det_voxel_num_points.view(-1, 1)
    gt_points = det_voxels[mask]
    gt_coords = det_voxel_coords.to(torch.int32)[mask.nonzero()[:, 0]]
    gt_points = torch.cat([gt_points, torch.zeros((gt_points.shape[0], 2), dtype=gt_points.dtype, device=gt_points.device)], dim=1)
    if sel is None:
        return None
    points = torch.cat([gt_points, sel["occ_points"]], dim=0)
    coords = torch.cat([gt_coords, sel["det_coords"]], dim=0)
    shape = [int(det_grid[2]), int(det_grid[1]), int(det_grid[0])]
    voxels, counts, vox_coords = revoxelize_sorted(coords, points, batch_size, shape)
    return voxels, counts, vox_coords, sel


# ------------------------------------------------------------------------------------------
# static (capacity) mode of the injection stage: no host read anywhere, CUDA-graph capturable
# ------------------------------------------------------------------------------------------
def occ_select_static(probs, residuals, thresh, max_points, cap, occ_voxel_size, occ_origin, det_voxel_size, det_range,
                      det_grid, rot_z=None, inten=0.0):
    """occ_select with a fixed capacity `cap` and the counts left on the device: returns (dict of capacity-sized tensors,
    counts [B+1] i32 with counts[B] = total).  The reference's top-k branch (more than `max_points` cells above threshold
    in one scene, add_occ_template.py:117-121) is not taken here: the per-scene counts are registered as capacity checks
    (ops.static_checks) and the caller falls back to the exact path when one fails."""
    _require_cuda(probs)
    lib = _lib.load()
    dev = probs.device
    probs = probs.to(torch.float32).contiguous()
    res = None if residuals is None else residuals.to(torch.float32).contiguous()
    B, nz, ny, nx = probs.shape
    grid = int3([nx, ny, nz])
    ws_bytes = int(lib.btc_occ_select_workspace_bytes(B, grid))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    gf = float_array(list(occ_voxel_size) + list(occ_origin) + list(det_voxel_size) + list(det_range[:3]))
    rz = None if rot_z is None else rot_z.to(torch.float32).contiguous()
    out = {"occ_coords": torch.zeros((cap, 4), dtype=torch.int32, device=dev),
           "occ_probs": torch.zeros(cap, dtype=torch.float32, device=dev),
           "occ_xyz": torch.zeros((cap, 3), dtype=torch.float32, device=dev),
           "det_coords": torch.zeros((cap, 4), dtype=torch.int32, device=dev),
           "occ_points": torch.zeros((cap, 6), dtype=torch.float32, device=dev)}
    counts = torch.zeros(B + 1, dtype=torch.int32, device=dev)
    check(lib.btc_occ_select(_ptr(probs), _ptr(res), B, grid, ctypes.c_float(thresh), _ptr(rz), gf, int3(det_grid),
                             ctypes.c_float(inten), cap, _ptr(out["occ_coords"]), _ptr(out["occ_probs"]),
                             _ptr(out["occ_xyz"]), _ptr(out["det_coords"]), _ptr(out["occ_points"]), _ptr(counts),
                             _ptr(ws), ws_bytes, _stream()), "btc_occ_select")
    _register_cap(counts[B:B + 1], cap, "occupancy cells above threshold")
    for b in range(B):
        _register_cap(counts[b:b + 1], int(max_points), "occupancy cells above threshold in scene %d (top-k branch)" % b)
    return out, counts


def pass_occ_vox_static(probs, residuals, det_voxels, det_voxel_num_points, det_voxel_coords, det_n_dev, batch_size, thresh,
                        max_points, occ_cap, vox_cap, p_max, occ_voxel_size, occ_origin, det_voxel_size, det_range, det_grid,
                        rot_z=None, inten=0.0):
    """pass_occ_vox without a host read.  det_voxels [Mcap, P, C] / det_voxel_num_points [Mcap] / det_voxel_coords
    [Mcap, 4] are capacity-sized with det_n_dev live rows.  Returns (voxels [vox_cap, p_max, C+2], num_points [vox_cap]
    i32, coords [vox_cap, 4] i32 (dead rows zero), m_dev [1] i32, selection dict, selection counts)."""
    lib = _lib.load()
    dev = det_voxels.device
    sel, sel_counts = occ_select_static(probs, residuals, thresh, max_points, occ_cap, occ_voxel_size, occ_origin, det_voxel_size,
                                        det_range, det_grid, rot_z=rot_z, inten=inten)
    Mcap, P, C = det_voxels.shape
    # raw points of the live det voxels in (voxel, slot) order == det_voxels[mask] of the reference (pass_occ_vox.py:33-41),
    # as static-shape index arithmetic: point j lives in the voxel whose inclusive count prefix first exceeds j
    nump = det_voxel_num_points.to(torch.int64)
    live = torch.arange(Mcap, device=dev) < det_n_dev.to(torch.int64)
    nump = torch.where(live, nump, torch.zeros_like(nump))
    ends = torch.cumsum(nump, 0)
    total_gt = ends[-1]
    j = torch.arange(Mcap * P, device=dev)
    vox = torch.searchsorted(ends, j, right=True).clamp_(max=Mcap - 1)
    slot = (j - (ends[vox] - nump[vox])).clamp_(0, P - 1)
    cap_pts = Mcap * P + occ_cap + 1                    # last row: dump for the dead occupancy rows
    points = torch.zeros((cap_pts, C + 2), dtype=torch.float32, device=dev)
    coords = torch.zeros((cap_pts, 4), dtype=torch.int32, device=dev)
    points[:Mcap * P, :C] = det_voxels.to(torch.float32)[vox, slot]
    coords[:Mcap * P] = det_voxel_coords.to(torch.int32)[vox]
    # occupancy pseudo points go right behind the live raw points
    n_occ = sel_counts[batch_size].to(torch.int64)
    i = torch.arange(occ_cap, device=dev)
    dst = torch.where(i < n_occ, total_gt + i, torch.full_like(i, cap_pts - 1))
    points.index_copy_(0, dst, sel["occ_points"])
    coords.index_copy_(0, dst, sel["det_coords"])
    n_pts = (total_gt + n_occ).to(torch.int32).reshape(1)
    shape = [int(det_grid[2]), int(det_grid[1]), int(det_grid[0])]
    n_entries = index_entries(batch_size, shape)
    index = torch.zeros(n_entries, dtype=torch.int64, device=dev)
    vox_coords = torch.zeros((vox_cap, 4), dtype=torch.int32, device=dev)
    vox_count = torch.zeros(vox_cap, dtype=torch.int32, device=dev)
    slots = torch.empty(cap_pts, dtype=torch.int32, device=dev)
    pt_voxel = torch.empty(cap_pts, dtype=torch.int32, device=dev)
    counts = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = int(lib.btc_revoxelize_workspace_bytes(max(cap_pts, vox_cap), n_entries))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    check(lib.btc_revoxelize(_ptr(coords), cap_pts, _ptr(n_pts), int(batch_size), int3(shape), _ptr(index), n_entries,
                             _ptr(vox_coords), vox_cap, _ptr(vox_count), _ptr(slots), _ptr(pt_voxel), _ptr(counts[0:1]),
                             _ptr(counts[1:2]), _ptr(ws), ws_bytes, _stream()), "btc_revoxelize")
    _register_cap(counts[0:1], vox_cap, "re-voxelised det voxels")
    _register_cap(counts[1:2], p_max, "points per re-voxelised voxel")
    voxels = torch.empty((vox_cap, p_max, C + 2), dtype=torch.float32, device=dev)
    check(lib.btc_revoxelize_fill(_ptr(points), _ptr(pt_voxel), _ptr(slots), cap_pts, _ptr(n_pts), C + 2, p_max, _ptr(voxels),
                                  vox_cap, _stream()), "btc_revoxelize_fill")
    return voxels, vox_count, vox_coords, counts[0:1], sel, sel_counts
