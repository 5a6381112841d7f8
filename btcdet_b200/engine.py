"""Sync-free planned forward of a sparse backbone, replayed as ONE CUDA graph per step.

The eager `spconv` shim keeps spconv-1.2.1 semantics for arbitrary torch code (exact tensor shapes,
one host read per new rulebook).  For serving, the same layers are compiled here into a static
launch sequence over capacity-sized buffers: every C-ABI call takes its live row count from device
memory (`n_dev`), so points -> voxels -> rulebooks -> convolutions run without a single host
synchronisation and the whole step is captured in a CUDA graph (one launch from the host instead of
the ~2 000 launches of the reference's per-offset gather/SGEMM/scatter path).
Eval-mode BatchNorm1d (btcdet/models/backbones_3d/spconv_backbone.py:33-38, eps 1e-3) is folded into
the conv epilogue as a per-channel affine, ReLU fused.
"""
import ctypes
from dataclasses import dataclass
from typing import List, Optional

import torch

import spconv

from . import _lib, ops
from ._lib import check, float_array, int3


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


@dataclass
class _Level:
    shape: list          # [D, H, W]
    cap: int
    coords: torch.Tensor     # [cap, 4] i32
    n_dev: torch.Tensor      # [1] i32 view
    index: Optional[torch.Tensor] = None   # rank bitmap (int64 entries) — sorted levels
    summary: Optional[torch.Tensor] = None  # one bit per index word (sparse two-level build / clear)
    perm: Optional[torch.Tensor] = None
    hash_keys: Optional[torch.Tensor] = None   # coordinate hash — the unsorted voxeliser level
    hash_vals: Optional[torch.Tensor] = None
    hash_shape: Optional[list] = None          # shape the hash keys were formed with (voxeliser's own hash), else `shape`


@dataclass
class _Step:
    kind: str            # "hash_build" | "subm_rb" | "conv_rb" | "sort_rb" | "conv"
    args: tuple


def fold_bn(bn: torch.nn.BatchNorm1d):
    """Eval-mode BN as y = x*scale + shift."""
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
    shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
    return scale, shift


class BackbonePlan:
    """points [N,4] -> GPU voxelisation -> MeanVFE -> sequential sparse backbone, one CUDA graph.

    layer_specs: list of (spconv.SparseConvolution, BatchNorm1d | None) in execution order
    (e.g. backbones.VoxelBackBone8x.layer_specs()), each consuming the previous layer's output.
    """

    def __init__(self, layer_specs, sparse_shape, batch, max_points_total, voxel_size, point_range, max_points=5,
                 max_voxels=16000, level_growth=2.0, algo=0, device="cuda", use_graph=True, sort_rows=False, side_priority=0,
                 tile_meta=True, split_format=True, early_conv_grid=(4, 116)):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.batch = int(batch)
        # (layers, ctas): the first `layers` convolutions run on at most `ctas` SMs, leaving the others to the rulebook
        # streams while most of the rulebook chain is still ahead.  The thin first layers are bound by their index-tile
        # pipeline, not by the SM count, and the rulebook kernels cannot share an SM with a persistent convolution CTA:
        # measured at batch 16 (tools/step_breakdown.py --early, profiles/r2_early_grid_sweep.json) the captured step drops
        # from 1 154 to 1 116 us with (4, 116); None = every layer on all SMs
        self.early_conv_grid = early_conv_grid
        self.voxel_size, self.point_range = list(voxel_size), list(point_range)
        self.grid = ops.voxel_grid_size(voxel_size, point_range)
        self.max_points, self.max_voxels = int(max_points), int(max_voxels)
        self.n_cap = int(max_points_total)
        self.algo = int(algo)
        self.use_graph = use_graph
        # mask-sorted tables for the block-skipping tensor-core tile: off by default — measured on B200 (batch 16) the eight
        # sort launches add 0.55 ms to the rulebook chain and save 0.07 ms of convolution
        self.sort_rows = bool(sort_rows)
        self._sorted = {}
        # per-rulebook tile masks + heaviest-first tile order for the tcgen05 tile (btc_rulebook_tile_meta)
        self.tile_meta = bool(tile_meta)
        self._meta = {}
        # split (bf16 hi / lo) feature format between consecutive tcgen05 layers (include/btcdet_b200.h): a layer writes it
        # when the next layer can gather it (c % 32 == 0 on both sides), the last layer always writes fp32
        specs = list(layer_specs)
        use_tc = [self.algo != 1 and not self.sort_rows and
                  ops.tc_supported(c.kernel_size[0] * c.kernel_size[1] * c.kernel_size[2], c.in_channels, c.out_channels)
                  for c, _ in specs]
        self._fmt = []
        for li, (c, _) in enumerate(specs):
            in_split = bool(self._fmt and self._fmt[-1][1])
            out_split = bool(split_format and use_tc[li] and li + 1 < len(specs) and use_tc[li + 1] and
                             c.out_channels % 32 == 0 and specs[li + 1][0].in_channels == c.out_channels)
            self._fmt.append((in_split, out_split))
        dev = self.device
        B = self.batch
        # ---- static input + voxelisation buffers ------------------------------------------
        self.points = torch.zeros((self.n_cap, 4), dtype=torch.float32, device=dev)
        self.scene_offsets = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        cap1 = B * self.max_voxels
        self.voxels = torch.empty((cap1, self.max_points, 4), dtype=torch.float32, device=dev)
        self.num_points = torch.empty(cap1, dtype=torch.int32, device=dev)
        self.n_voxels = torch.zeros(B + 1, dtype=torch.int32, device=dev)
        self.vox_ws = torch.empty(int(self.lib.btc_voxelize_workspace_bytes(self.n_cap, B, self.max_voxels,
                                                                             self.max_points)),
                                  dtype=torch.uint8, device=dev)
        self.counts = []          # device count tensors to read back (overflow check / output size)
        # live counts of every level in ONE small tensor (slot 0 = the voxeliser's total, copied; slot i = level i, written
        # in place by the rulebook build): a single D2H copy reads them all
        self.dev_counts = torch.zeros(16, dtype=torch.int32, device=dev)
        # ---- levels, rulebooks, feature buffers ---------------------------------------------
        shape0 = [int(s) for s in sparse_shape]
        lvl = _Level(shape0, cap1, torch.empty((cap1, 4), dtype=torch.int32, device=dev), self.n_voxels[B:B + 1])
        self.levels = [lvl]
        self.feat0 = torch.empty((cap1, 4), dtype=torch.float32, device=dev)   # MeanVFE output
        self.steps: List[_Step] = []
        self.rulebooks = {}
        self._ws = {}
        cur_feat, cur_lvl = self.feat0, lvl
        self.params = []  # keep folded tensors alive
        for li, (conv, bn) in enumerate(specs):
            assert isinstance(conv, spconv.SparseConvolution) and conv.ndim == 3 and not conv.inverse
            in_split, out_split = self._fmt[li]
            K = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
            key = conv.indice_key
            rb = self.rulebooks.get(key) if key is not None else None
            if rb is None:
                # (thin sub-manifold layers, c_in <= 16, run all chunks: no mask launch — see the index loader of conv_fwd_tc;
                # measured: 131 -> 124 us for the two level-1 layers plus 22 us of mask launch off the serial front of the
                # step, while the strided 16 -> 32 layer loses 15 us without its masks and keeps them)
                users = [c for c, _ in specs[li:] if c is conv or (key is not None and c.indice_key == key)]
                want_meta = self.tile_meta and not self.sort_rows and K <= 64 and self.algo != 1 and \
                    ops.tc_supported(K, conv.in_channels, conv.out_channels) and \
                    (not conv.subm or any(c.in_channels > 16 for c in users))
                if conv.subm:
                    if cur_lvl.index is None and cur_lvl.hash_keys is None:
                        self._add_index(cur_lvl)
                    nbr = torch.empty((cur_lvl.cap, K), dtype=torch.int32, device=dev)
                    self.steps.append(_Step("subm_rb", (cur_lvl, conv.kernel_size, conv.dilation, nbr)))
                    if want_meta:
                        meta = self._new_meta(nbr)
                        self.steps.append(_Step("tile_meta", (nbr, cur_lvl, meta[0], meta[1])))
                    rb = (nbr, cur_lvl)
                else:
                    assert not conv.transposed, "planned engine covers the det backbone (no transposed convs)"
                    out_shape = ops.conv_output_shape(cur_lvl.shape, conv.kernel_size, conv.stride, conv.padding,
                                                      conv.dilation)
                    cells = B * out_shape[0] * out_shape[1] * out_shape[2]
                    cap = int(min(cells, max(1024, level_growth * cur_lvl.cap)))
                    assert len(self.levels) < 16
                    n_dev = self.dev_counts[len(self.levels):len(self.levels) + 1]
                    out_lvl = _Level(out_shape, cap, torch.empty((cap, 4), dtype=torch.int32, device=dev), n_dev)
                    out_lvl.index = torch.zeros(ops.index_entries(B, out_shape), dtype=torch.int64, device=dev)
                    out_lvl.summary = torch.zeros(int(self.lib.btc_index_summary_words(out_lvl.index.numel())),
                                                  dtype=torch.int32, device=dev)
                    nbr = torch.empty((cap, K), dtype=torch.int32, device=dev)
                    self.steps.append(_Step("conv_rb", (cur_lvl, out_lvl, conv.kernel_size, conv.stride, conv.padding,
                                                        conv.dilation, nbr)))
                    if want_meta:
                        meta = self._new_meta(nbr)
                        self.steps.append(_Step("tile_meta", (nbr, out_lvl, meta[0], meta[1])))
                    self.levels.append(out_lvl)
                    self.counts.append((n_dev, cap))
                    rb = (nbr, out_lvl)
                if key is not None:
                    self.rulebooks[key] = rb
            nbr, out_lvl = rb
            w = conv.weight.detach().to(dev, torch.float32).reshape(K, conv.in_channels, conv.out_channels).contiguous()
            bias = None if conv.bias is None else conv.bias.detach().to(dev, torch.float32).contiguous()
            scale = shift = None
            if bn is not None:
                scale, shift = (t.to(dev) for t in fold_bn(bn))
            out_feat = torch.empty((out_lvl.cap, conv.out_channels), dtype=torch.float32, device=dev)
            # wide layers run on the tcgen05 tile (weights packed once into the UMMA operand image)
            packed = None
            if self.algo != 1 and ops.tc_supported(K, conv.in_channels, conv.out_channels):
                packed = ops.tc_pack_weight_split(w) if in_split else ops.tc_pack_weight(w)
            self.params.append((w, bias, scale, shift, packed))
            meta = self._meta.get(id(nbr)) if packed is not None else None
            rows = None
            if packed is not None and self.sort_rows:
                rows = self._sorted.get(id(nbr))
                if rows is None:
                    rows = (torch.empty_like(nbr), torch.empty(nbr.shape[0], dtype=torch.int32, device=dev))
                    self._sorted[id(nbr)] = rows
                    self.steps.append(_Step("sort_rb", (nbr, out_lvl, rows[0], rows[1])))
            self.steps.append(_Step("conv", (cur_feat, nbr, w, bias, scale, shift, bn is not None, out_feat, out_lvl, K,
                                             conv.in_channels, conv.out_channels, packed, rows, meta, (in_split, out_split))))
            cur_feat, cur_lvl = out_feat, out_lvl
        self.out_feat, self.out_lvl = cur_feat, cur_lvl
        # sparse clear of every sorted level's bitmaps right after their last reader (the rulebook that built them or a
        # sub-manifold rulebook on that level): the next step starts from all-zero bitmaps without a memset of the grid
        for lvl in self.levels:
            if lvl.summary is None:
                continue
            last = max(i for i, st in enumerate(self.steps)
                       if (st.kind == "conv_rb" and st.args[1] is lvl) or (st.kind == "subm_rb" and st.args[0] is lvl))
            self.steps.insert(last + 1, _Step("index_clear", (lvl,)))
        self.graph = None
        # (a high-priority rulebook stream, side_priority=-1, was measured: no effect on the captured step, 1515 us both ways)
        self._side_stream = torch.cuda.Stream(device=dev, priority=side_priority)
        self._side_stream2 = torch.cuda.Stream(device=dev, priority=side_priority)
        self.launches_per_step = 0
        self.host_counts = torch.zeros(len(self.levels) + 1, dtype=torch.int32).pin_memory()

    # ------------------------------------------------------------------------------------
    def _add_index(self, lvl: _Level):
        """Unsorted level (voxeliser order): coordinate hash instead of a 92 M-cell bitmap."""
        if lvl is self.levels[0] and lvl.shape[1] == self.grid[1] and lvl.shape[2] == self.grid[0] and lvl.shape[0] >= self.grid[2]:
            # the voxeliser's own hash (keys over its grid, vals = voxel rows) IS this level's coordinate hash: no build.
            # Its keys use depth grid_z (the sparse shape has one more layer that no voxel can occupy), so the probes run
            # with that depth — a neighbour in the extra layer is absent either way.
            ko, vo, ns = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
            check(self.lib.btc_voxelize_hash_view(self.n_cap, self.batch, self.max_voxels, self.max_points, ctypes.byref(ko),
                                                  ctypes.byref(vo), ctypes.byref(ns)), "btc_voxelize_hash_view")
            lvl.hash_keys = self.vox_ws[ko.value:ko.value + 8 * ns.value].view(torch.int64)
            lvl.hash_vals = self.vox_ws[vo.value:vo.value + 4 * ns.value].view(torch.int32)
            lvl.hash_shape = [int(self.grid[2]), lvl.shape[1], lvl.shape[2]]
            return
        n_slots = int(self.lib.btc_hash_slots(lvl.cap))
        lvl.hash_keys = torch.empty(n_slots, dtype=torch.int64, device=self.device)
        lvl.hash_vals = torch.empty(n_slots, dtype=torch.int32, device=self.device)
        self.steps.append(_Step("hash_build", (lvl,)))

    def _new_meta(self, nbr):
        tiles = (nbr.shape[0] + 127) // 128
        meta = (torch.zeros(tiles, dtype=torch.int64, device=self.device),
                torch.zeros(int(self.lib.btc_rulebook_tile_order_ints(nbr.shape[0])), dtype=torch.int32, device=self.device))
        self._meta[id(nbr)] = meta
        return meta

    def _workspace(self, nbytes):
        ws = self._ws.get("ws")
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
            self._ws["ws"] = ws
        return ws

    def _run(self):
        """Enqueue the whole step (no host synchronisation anywhere).

        Three streams (parallel branches of the captured graph).  The rulebook work depends only on coordinates and is a
        DAG, not a chain: the strided rulebooks form the critical path (voxelise -> level 2 -> level 3 -> ...), while
        each level's hash / sub-manifold table / tile metadata / bitmap clear hang off it.  `side` carries the critical
        path, `side2` everything else, `main` the convolutions, each behind the events of exactly what it reads.  The
        small latency-bound indexing kernels then overlap each other (they only get SM time in the gaps between the
        persistent convolution launches, which hold every SM's registers and shared memory).
        """
        main = torch.cuda.current_stream()
        side, side2 = self._side_stream, self._side_stream2
        launches = 0
        side.wait_stream(main)
        side2.wait_stream(main)
        written, readers = {}, {}     # resource -> event of its last writer / events of the reads since then

        def enqueue(stream, reads, writes, fn):
            for r in reads:
                ev = written.get(r)
                if ev is not None and ev[1] is not stream:
                    stream.wait_event(ev[0])
            for w in writes:                     # write-after-read / write-after-write across streams
                for ev in readers.get(w, []) + ([written[w]] if w in written else []):
                    if ev[1] is not stream:
                        stream.wait_event(ev[0])
            with torch.cuda.stream(stream):
                n = fn(ctypes.c_void_p(stream.cuda_stream))
                e = torch.cuda.Event()
                e.record(stream)
            for r in reads:
                readers.setdefault(r, []).append((e, stream))
            for w in writes:
                written[w] = (e, stream)
                readers[w] = []
            return n

        lvl0 = self.levels[0]
        n_conv = 0
        # grouping on the critical path of the rulebook chain; the voxel contents / MeanVFE features (only the first
        # convolution reads them) on the main stream, which is idle until then
        launches += enqueue(side, [], [("coords", id(lvl0)), ("index", id(lvl0))], lambda st: self.launch_voxelize(st, 1))
        launches += enqueue(main, [("coords", id(lvl0))], [("feat0", 0)], lambda st: self.launch_voxelize(st, 2))
        for s in self.steps:
            if s.kind == "conv":
                nbr, rows, meta = s.args[1], s.args[13], s.args[14]
                reads = [("nbr", id(nbr))] + ([("meta", id(nbr))] if (meta is not None or rows is not None) else [])
                for r in reads:
                    ev = written.get(r)
                    if ev is not None:
                        main.wait_event(ev[0])
                with torch.cuda.stream(main):
                    if self.early_conv_grid is not None:
                        capped = n_conv < self.early_conv_grid[0]
                        check(self.lib.btc_sparse_conv_tc_grid(int(self.early_conv_grid[1]) if capped else 148), "tc grid")
                    n_conv += 1
                    self.launch_conv(s.args, ctypes.c_void_p(main.cuda_stream))
                    e = torch.cuda.Event()
                    e.record(main)
                for r in reads:
                    readers.setdefault(r, []).append((e, main))
                launches += 1
                continue
            if s.kind == "hash_build":
                (lvl,) = s.args
                reads, writes, stream = [("coords", id(lvl))], [("index", id(lvl))], side2
            elif s.kind == "subm_rb":
                lvl, nbr = s.args[0], s.args[3]
                reads, writes, stream = [("coords", id(lvl)), ("index", id(lvl))], [("nbr", id(nbr))], side2
            elif s.kind == "conv_rb":
                lin, lout, nbr = s.args[0], s.args[1], s.args[6]
                reads, writes, stream = [("coords", id(lin))], [("coords", id(lout)), ("index", id(lout)), ("nbr", id(nbr))], side
            elif s.kind == "tile_meta":
                nbr, lvl = s.args[0], s.args[1]
                reads, writes, stream = [("nbr", id(nbr)), ("coords", id(lvl))], [("meta", id(nbr))], side2
            elif s.kind == "index_clear":
                (lvl,) = s.args
                reads, writes, stream = [("coords", id(lvl))], [("index", id(lvl))], side2
            elif s.kind == "sort_rb":
                nbr, lvl = s.args[0], s.args[1]
                reads, writes, stream = [("nbr", id(nbr)), ("coords", id(lvl))], [("meta", id(nbr))], side2
            else:
                raise ValueError(s.kind)
            launches += enqueue(stream, reads, writes, lambda st, s=s, stream=stream: self._launch_index_on(s, st, stream))
        if self.early_conv_grid is not None:
            check(self.lib.btc_sparse_conv_tc_grid(148), "tc grid")      # process-wide knob: back to one CTA per SM
        main.wait_stream(side)
        main.wait_stream(side2)
        # gather the live counts of every level into one small tensor (read back lazily by the caller)
        self.dev_counts[0:1].copy_(self.levels[0].n_dev)       # levels 1.. write their slot themselves
        self.launches_per_step = launches
        return launches

    def _launch_index_on(self, s, st, stream):
        with torch.cuda.stream(stream):
            return self.launch_index_step(s, st)

    def launch_voxelize(self, st, phase=3):
        """points -> voxels / coords / counts / MeanVFE features of level 0 (9 launches).  phase 1 = grouping only (coords and
        counts: what the rulebook chain waits for), phase 2 = contents (voxels, MeanVFE features: what the first convolution
        waits for), 3 = both."""
        lib, B, lvl0 = self.lib, self.batch, self.levels[0]
        fn = {1: lib.btc_voxelize_group, 2: lib.btc_voxelize_fill, 3: lib.btc_voxelize}[phase]
        check(fn(_ptr(self.points), self.n_cap, 4, _ptr(self.scene_offsets), B,
                 float_array(self.voxel_size), float_array(self.point_range), int3(self.grid),
                 self.max_points, self.max_voxels, _ptr(self.voxels), _ptr(lvl0.coords),
                 _ptr(self.num_points), _ptr(self.feat0), _ptr(self.n_voxels), _ptr(self.vox_ws),
                 self.vox_ws.numel(), st), "btc_voxelize")
        return {1: 7, 2: 2, 3: 9}[phase]

    def launch_index_step(self, s, st):
        """One step of the rulebook chain (hash build / neighbour table / mask sort) on stream `st`;
        returns the number of kernel launches it enqueued.  Must be called with `st` as torch's current stream
        (the bitmap clear of a strided level is a torch op)."""
        lib, B = self.lib, self.batch
        if s.kind == "hash_build":
            (lvl,) = s.args
            check(lib.btc_hash_build(_ptr(lvl.coords), lvl.cap, _ptr(lvl.n_dev), B, int3(lvl.shape),
                                     _ptr(lvl.hash_keys), _ptr(lvl.hash_vals), lvl.hash_keys.numel(), st),
                  "btc_hash_build")
            return 1
        if s.kind == "subm_rb":
            lvl, ksize, dil, nbr = s.args
            if lvl.hash_keys is not None:
                check(lib.btc_rulebook_subm_hash(_ptr(lvl.coords), lvl.cap, _ptr(lvl.n_dev), B, int3(lvl.hash_shape or lvl.shape),
                                                 int3(ksize), int3(dil), _ptr(lvl.hash_keys), _ptr(lvl.hash_vals),
                                                 lvl.hash_keys.numel(), _ptr(nbr), st), "btc_rulebook_subm_hash")
            else:
                check(lib.btc_rulebook_subm(_ptr(lvl.coords), lvl.cap, _ptr(lvl.n_dev), B, int3(lvl.shape),
                                            int3(ksize), int3(dil), _ptr(lvl.index), lvl.index.numel(),
                                            _ptr(lvl.perm), _ptr(nbr), st), "btc_rulebook_subm")
            K = nbr.shape[1]
            return 2 if (K % 2 == 1 and K // 2 <= 16) else 1   # -1 fill + symmetric half-probe kernel
        if s.kind == "conv_rb":
            lin, lout, ksize, stride, pad, dil, nbr = s.args
            # sparse two-level build: both bitmaps are all-zero here (allocation / the previous step's index_clear)
            ws = self._workspace(lib.btc_rulebook_conv_sparse_workspace_bytes(lout.index.numel()))
            check(lib.btc_rulebook_conv_sparse(_ptr(lin.coords), lin.cap, _ptr(lin.n_dev), B, int3(lin.shape),
                                               int3(lout.shape), int3(ksize), int3(stride), int3(pad), int3(dil), 0,
                                               _ptr(lout.index), lout.index.numel(), _ptr(lout.summary), _ptr(lout.coords),
                                               lout.cap, _ptr(lout.n_dev), _ptr(nbr), None, _ptr(ws), ws.numel(), st),
                  "btc_rulebook_conv_sparse")
            return 4
        if s.kind == "tile_meta":
            nbr, lvl, tmask, torder = s.args
            check(lib.btc_rulebook_tile_meta(_ptr(nbr), lvl.cap, _ptr(lvl.n_dev), nbr.shape[1], _ptr(tmask), _ptr(torder), st),
                  "btc_rulebook_tile_meta")
            return 1
        if s.kind == "index_clear":
            (lvl,) = s.args
            check(lib.btc_index_clear_sparse(_ptr(lvl.coords), lvl.cap, _ptr(lvl.n_dev), B, int3(lvl.shape), _ptr(lvl.index),
                                             lvl.index.numel(), _ptr(lvl.summary), st), "btc_index_clear_sparse")
            return 1
        if s.kind == "sort_rb":
            nbr, lvl, nbr_sorted, out_rows = s.args
            check(lib.btc_rulebook_sort_rows(_ptr(nbr), lvl.cap, _ptr(lvl.n_dev), nbr.shape[1], _ptr(nbr_sorted),
                                             _ptr(out_rows), st), "btc_rulebook_sort_rows")
            return 1
        raise ValueError(s.kind)

    def launch_conv(self, args, st):
        """One sparse-conv layer: tcgen05 tile when the weights were packed, fp32 FFMA tile otherwise."""
        fin, nbr, w, bias, scale, shift, relu, fout, lout, K, cin, cout, packed, rows, meta, fmt = args
        if packed is not None and rows is None:
            tm, to = (None, None) if meta is None else meta
            check(self.lib.btc_sparse_conv_fwd_tc_split(_ptr(fin), _ptr(nbr), _ptr(packed), _ptr(bias), _ptr(scale),
                                                        _ptr(shift), int(relu), _ptr(fout), lout.cap, _ptr(lout.n_dev), K,
                                                        cin, cout, int(fmt[0]), int(fmt[1]), _ptr(tm), _ptr(to), st),
                  "btc_sparse_conv_fwd_tc_split")
        elif packed is not None and rows is not None:
            check(self.lib.btc_sparse_conv_fwd_tc_rows(_ptr(fin), _ptr(rows[0]), _ptr(rows[1]), _ptr(packed), _ptr(bias),
                                                       _ptr(scale), _ptr(shift), int(relu), _ptr(fout), lout.cap,
                                                       _ptr(lout.n_dev), K, cin, cout, st), "btc_sparse_conv_fwd_tc_rows")
        elif packed is not None:
            check(self.lib.btc_sparse_conv_fwd_tc(_ptr(fin), _ptr(nbr), _ptr(packed), _ptr(bias), _ptr(scale),
                                                  _ptr(shift), int(relu), _ptr(fout), lout.cap, _ptr(lout.n_dev), K,
                                                  cin, cout, st), "btc_sparse_conv_fwd_tc")
        else:
            check(self.lib.btc_sparse_conv_fwd(_ptr(fin), _ptr(nbr), _ptr(w), _ptr(bias), _ptr(scale), _ptr(shift),
                                               int(relu), _ptr(fout), lout.cap, _ptr(lout.n_dev), K, cin, cout, 1, st),
                  "btc_sparse_conv_fwd")

    def capture(self):
        """Warm up once on a side stream, then capture the step into a CUDA graph."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()
        return self

    def load_points(self, points: torch.Tensor, scene_offsets: torch.Tensor):
        """Copy one batch (device or pinned-host tensors) into the static input buffers (async)."""
        n = points.shape[0]
        if n > self.n_cap:
            raise _lib.BtcError("batch of %d points exceeds the planned capacity %d" % (n, self.n_cap))
        self.points[:n].copy_(points, non_blocking=True)
        self.scene_offsets.copy_(scene_offsets, non_blocking=True)

    def step(self):
        """Replay the captured step (or enqueue it eagerly when graphs are disabled)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._run()

    def forward(self, points, scene_offsets):
        self.load_points(points, scene_offsets)
        self.step()
        return self.out_feat, self.out_lvl.coords, self.out_lvl.n_dev

    # ---- pipelined host-facing call: H2D, graph replay and D2H of consecutive batches overlap ------------
    def enable_pipeline(self, result_rows=None, slots=2):
        """Allocate the staging needed by submit()/retrieve(): per slot a compact device copy of the result rows
        (so the next replay may overwrite the static output buffer) and a pinned host buffer."""
        dev, C = self.device, self.out_feat.shape[1]
        rows = int(result_rows or min(self.out_lvl.cap, self.batch * 16384))
        self._pl = {
            "rows": rows, "slots": slots, "k": 0, "pending": [],
            "stage_feat": [torch.empty((rows, C), dtype=torch.float32, device=dev) for _ in range(slots)],
            "stage_coords": [torch.empty((rows, 4), dtype=torch.int32, device=dev) for _ in range(slots)],
            "stage_n": [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(slots)],
            "host_feat": [torch.empty((rows, C), dtype=torch.float32).pin_memory() for _ in range(slots)],
            "host_coords": [torch.empty((rows, 4), dtype=torch.int32).pin_memory() for _ in range(slots)],
            "host_n": [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(slots)],
            "host_levels": [torch.zeros(len(self.levels), dtype=torch.int32).pin_memory() for _ in range(slots)],
            "count_ready": [torch.cuda.Event() for _ in range(slots)],
            "d2h_done": [None] * slots,
            "copy_stream": torch.cuda.Stream(device=dev),
            # host -> device on its own stream: the points of batch i + 1 cross PCIe while batch i computes, the main stream
            # only pays a device-to-device copy into the graph's static input buffer
            "h2d_stream": torch.cuda.Stream(device=dev),
            "in_points": [torch.empty_like(self.points) for _ in range(slots)],
            "in_offsets": [torch.empty_like(self.scene_offsets) for _ in range(slots)],
            "in_ready": [torch.cuda.Event() for _ in range(slots)],
            "in_free": [None] * slots,
        }
        return self

    def submit(self, points, scene_offsets):
        """Enqueue one batch (pinned host or device tensors): H2D, the graph, staging of the live result rows and
        the tiny count read-back.  Returns immediately; call retrieve() for results in submission order."""
        pl = self._pl
        slot = pl["k"] % pl["slots"]
        main = torch.cuda.current_stream()
        if pl["d2h_done"][slot] is not None:
            main.wait_event(pl["d2h_done"][slot])      # the previous occupant of this slot has left the device
        if points.is_cuda:
            self.load_points(points, scene_offsets)
        else:
            n = points.shape[0]
            if n > self.n_cap:
                raise _lib.BtcError("batch of %d points exceeds the planned capacity %d" % (n, self.n_cap))
            hs = pl["h2d_stream"]
            if pl["in_free"][slot] is not None:
                hs.wait_event(pl["in_free"][slot])     # the staging buffers' previous batch has been consumed
            with torch.cuda.stream(hs):
                pl["in_points"][slot][:n].copy_(points, non_blocking=True)
                pl["in_offsets"][slot].copy_(scene_offsets, non_blocking=True)
                pl["in_ready"][slot].record(hs)
            main.wait_event(pl["in_ready"][slot])
            self.points[:n].copy_(pl["in_points"][slot][:n], non_blocking=True)
            self.scene_offsets.copy_(pl["in_offsets"][slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            pl["in_free"][slot] = ev
        self.step()
        st = ctypes.c_void_p(main.cuda_stream)
        C = self.out_feat.shape[1]
        check(self.lib.btc_copy_rows(_ptr(self.out_feat), _ptr(pl["stage_feat"][slot]), min(self.out_lvl.cap, pl["rows"]),
                                     _ptr(self.out_lvl.n_dev), C * 4, st), "btc_copy_rows")
        check(self.lib.btc_copy_rows(_ptr(self.out_lvl.coords), _ptr(pl["stage_coords"][slot]),
                                     min(self.out_lvl.cap, pl["rows"]), _ptr(self.out_lvl.n_dev), 16, st), "btc_copy_rows")
        pl["stage_n"][slot].copy_(self.out_lvl.n_dev, non_blocking=True)
        pl["host_n"][slot].copy_(pl["stage_n"][slot], non_blocking=True)
        pl["host_levels"][slot].copy_(self.dev_counts[:len(self.levels)], non_blocking=True)   # every level, for the overflow check
        pl["count_ready"][slot].record(main)
        pl["pending"].append(slot)
        pl["k"] += 1

    def retrieve(self, to_host=True):
        """Result of the oldest submitted batch: (features [n,C], coords [n,4]) as views of pinned host buffers,
        valid until two further submits.  Only this call waits, and only for that batch.
        to_host=False: the rows stay on the device — (features, coords) are views of the slot's DEVICE staging buffers for
        a consumer on the GPU (the reference's BEV backbone reads the backbone output on the device, btcnet.py:56-88),
        the host only receives the row counts of every level; the caller hands the slot back by recording its own event
        into `plan._pl["d2h_done"][slot]` or simply by finishing its work on the current stream before the slot's reuse
        (two submits later)."""
        pl = self._pl
        slot = pl["pending"].pop(0)
        pl["count_ready"][slot].synchronize()          # the row count of that batch (the GPU is already busy with the next)
        n = int(pl["host_n"][slot][0])
        # an intermediate level that overflowed its capacity clamps its rows silently inside the kernels: fail loudly here
        for c, l in zip(pl["host_levels"][slot].tolist(), self.levels):
            if c > l.cap:
                raise _lib.BtcError("level capacity exceeded in a pipelined batch: %d sites > capacity %d — raise level_growth"
                                    % (c, l.cap))
        if n > pl["rows"]:
            raise _lib.BtcError("result of %d rows exceeds the pipeline staging capacity %d" % (n, pl["rows"]))
        if not to_host:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())      # consumers on the current stream are ordered before the slot's reuse
            pl["d2h_done"][slot] = ev
            return pl["stage_feat"][slot][:n], pl["stage_coords"][slot][:n], ev
        cs = pl["copy_stream"]
        cs.wait_event(pl["count_ready"][slot])
        with torch.cuda.stream(cs):
            pl["host_feat"][slot][:n].copy_(pl["stage_feat"][slot][:n], non_blocking=True)
            pl["host_coords"][slot][:n].copy_(pl["stage_coords"][slot][:n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
        pl["d2h_done"][slot] = ev
        pl.setdefault("last_d2h", []).append(ev)
        return pl["host_feat"][slot][:n], pl["host_coords"][slot][:n], ev

    def read_counts(self):
        """One small D2H copy of every level's live count; raises if a capacity overflowed."""
        self.host_counts[:len(self.levels)].copy_(self.dev_counts[:len(self.levels)], non_blocking=False)
        counts = self.host_counts[:len(self.levels)].tolist()
        for c, l in zip(counts, self.levels):
            if c > l.cap:
                for lv in self.levels:   # rows past a capacity were never listed, so the sparse clear missed their bits
                    if getattr(lv, "index", None) is not None:
                        lv.index.zero_()
                    if getattr(lv, "summary", None) is not None:
                        lv.summary.zero_()
                raise _lib.BtcError("level capacity exceeded: %d sites > capacity %d — raise level_growth" % (c, l.cap))
        return counts

    def level_features(self):
        """(features, coords, n_dev) of the last layer output (capacity-sized buffers)."""
        return self.out_feat, self.out_lvl.coords, self.out_lvl.n_dev
