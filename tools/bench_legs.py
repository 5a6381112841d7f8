"""Additional measurement legs of bench.py (imported by it; each returns a small dict for the JSON line).

  native_gpu_leg   a GPU re-creation of spconv 1.2.1's Native algorithm: per layer
                   the rulebook in spconv's pair format with the per-offset counts copied to the HOST (as spconv sizes
                   its GEMMs), then per kernel offset {index_select gather, torch.mm (cuBLAS SGEMM), index_add_
                   scatter-add} — 27 x 3 launches per 3x3x3 layer — on the same scenes, same weights, fp32.  The rulebooks
                   themselves come from this library (spconv's own indice kernels are not available offline), which
                   favours the baseline.  This is the denominator of the north-star ">= 10x the spconv CUDA backbone".
  train_leg        BASELINE config 4: fwd + bwd + DDP gradient all-reduce (NCCL) + clip + Adam, 2 scenes per GPU.
  chain_leg        BASELINE config 3: btcdet_b200.chain.BtcHotPath on 2 scenes (occ targets -> occ backbone -> head ->
                   injection -> det backbone -> BEV features).
  stress_leg       BASELINE config 5: 100k-point clouds at voxel 0.025 m; rulebook-build / gather-GEMM bytes per second
                   against the measured HBM peak.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def native_gpu_leg(model, scenes, B, dev, steps=3, warmup=1):
    from btcdet_b200 import ops, synthetic as S
    model = model.to(dev).eval()
    pts, offs = S.batch_points(scenes[:B])
    v, c, n, mean, nv = ops.voxelize(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev), S.DET_VOXEL_SIZE,
                                     S.KITTI_RANGE, S.DET_MAX_POINTS, S.DET_MAX_VOXELS["train"], want_mean=True)
    m = int(nv[-1].item())
    feats0, coords0 = mean[:m].contiguous(), c[:m].contiguous()
    specs = model.layer_specs()
    launches = [0]

    def forward():
        x, coords, shape = feats0, coords0, list(model.sparse_shape)
        cache = {}
        launches[0] = 0
        for conv, bn in specs:
            K = conv.kernel_size[0] * conv.kernel_size[1] * conv.kernel_size[2]
            ent = cache.get(conv.indice_key)
            if ent is None:
                rb = ops.rulebook_subm(coords, B, shape, conv.kernel_size) if conv.subm else \
                    ops.rulebook_conv(coords, B, shape, conv.kernel_size, conv.stride, conv.padding, conv.dilation)
                pairs, pair_num = rb.pairs()
                ent = (rb, pairs.long(), pair_num.cpu().tolist())      # host-synchronised counts, as spconv 1.2.1
                cache[conv.indice_key] = ent
            rb, pairs, counts = ent
            W = conv.weight.reshape(K, conv.in_channels, conv.out_channels)
            if conv.subm:
                out = torch.mm(x, W[K // 2])                            # centre offset first (App. A.5)
                launches[0] += 1
            else:
                out = torch.zeros((rb.n_out, conv.out_channels), dtype=torch.float32, device=dev)
            for k in range(K):
                nk = counts[k]
                if nk == 0 or (conv.subm and k == K // 2):
                    continue
                buf = x.index_select(0, pairs[0, k, :nk])
                out.index_add_(0, pairs[1, k, :nk], torch.mm(buf, W[k]))
                launches[0] += 3
            if bn is not None:
                out = torch.relu(bn(out))
            x, coords, shape = out, rb.out_coords, rb.out_shape
        return x

    with torch.no_grad():
        for _ in range(warmup):
            ref_out = forward()
        torch.cuda.synchronize()
        e0, e1 = _events()
        e0.record()
        for _ in range(steps):
            forward()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": round(B / (ms * 1e-3), 2), "unit": "scenes/s", "ms_per_step": round(ms, 3), "scenes_per_step": B,
            "launches_per_step": launches[0], "rows_out": int(ref_out.shape[0]),
            "what": "GPU re-creation of spconv-1.2.1 Native: per offset index_select -> torch.mm (cuBLAS SGEMM, fp32) -> "
                    "index_add_, pair counts read on the host per rulebook; backbone only (voxel features already on the device), "
                    "rulebooks from this library"}, ref_out


class _TrainNet(torch.nn.Module):
    """VoxelBackBone8x + HeightCompression-style dense() + a 1x1 BEV head standing in for the RPN."""

    def __init__(self):
        super().__init__()
        from btcdet_b200 import backbones
        self.backbone = backbones.VoxelBackBone8x(4)
        self.head = torch.nn.Conv2d(256, 8, 1)

    def forward(self, feats, coords, batch):
        out = self.backbone({"voxel_features": feats, "voxel_coords": coords, "batch_size": batch})["encoded_spconv_tensor"]
        d = out.dense()
        n, c, dd, h, w = d.shape
        return self.head(d.view(n, c * dd, h, w))


def train_leg(rank, world, dev, steps=10, warmup=4, batch=2):
    """Runs on EVERY rank (DDP all-reduce over NCCL when world > 1); returns the dict on every rank."""
    from btcdet_b200 import ops, synthetic as S
    torch.manual_seed(0)
    net = _TrainNet().to(dev).train()
    model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[dev.index]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batches = []
    for i in range(3):
        sc = [S.lidar_like(20000, seed=7000 + 100 * rank + 10 * i + b) for b in range(batch)]
        pts, offs = S.batch_points(sc)
        batches.append((torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)))

    def step(i):
        pts, offs = batches[i % len(batches)]
        v, c, n, mean, nv = ops.voxelize(pts, offs, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000, want_mean=True)
        m = int(nv[-1].item())
        loss = model(mean[:m], c[:m], batch).square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    n_param = sum(p.numel() for p in net.parameters())
    return {"value": round(world * batch * steps / (ms * 1e-3), 2), "unit": "scenes/s", "ms_per_step": round(ms / steps, 3),
            "scenes_per_gpu": batch, "steps": steps, "allreduce_bytes_per_step": 4 * n_param if world > 1 else 0,
            "loss": float(loss.detach()),
            "what": "config 4 shape: fwd + bwd (CUDA dX / dW / db kernels) + DDP bucketed NCCL all-reduce overlapped with "
                    "backward + grad clip + Adam, VoxelBackBone8x + BEV 1x1 head, train-mode BatchNorm, eager spconv shim"}


def batch_leg(model, scenes, dev, batch=32, steps=20, warmup=3):
    """The headline workload at another batch size (same plan, same timing rules: HBM-resident input, L2 flushed between
    timed steps, CUDA events per step): how far the latency-bound parts of the step amortise with more scenes per step."""
    from btcdet_b200 import engine, synthetic as S
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, batch, batch * 20000, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=S.DET_MAX_POINTS, max_voxels=S.DET_MAX_VOXELS["train"], device=dev).capture()
    pts, offs = S.batch_points([scenes[j % len(scenes)] for j in range(batch)])
    p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    ms = []
    for i in range(warmup + steps):
        flush.fill_(float(i))
        e0, e1 = _events()
        e0.record()
        plan.load_points(p, o)       # D2D into the graph's static input buffer, inside the timed region as in bench.py
        plan.step()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ms.append(e0.elapsed_time(e1))
    plan.read_counts()
    med = float(np.median(ms))
    del plan
    torch.cuda.empty_cache()
    return {"scenes_per_step": batch, "value": round(batch / (med * 1e-3), 2), "unit": "scenes/s", "ms_per_step": round(med, 4),
            "steps": steps, "what": "the headline step at batch %d instead of 16 (one distinct batch, HBM-resident, L2 flushed "
                                    "between timed steps): the rulebook chain and the per-launch costs amortise" % batch}


def chain_leg(dev, steps=10, warmup=3, batch=2):
    from btcdet_b200 import backbones, chain
    torch.manual_seed(0)
    model = chain.BtcHotPath()
    backbones.randomize_bn_(model, 0)
    model = model.to(dev).eval()
    bds = [chain.synthetic_batch([900 + 10 * i + b for b in range(batch)], n_points=20000, device=dev, with_rot=True, mode="test")
           for i in range(2)]
    chain.calibrate_occ_head_bias(model, bds[0], 0.03)   # random-init head: let ~3 % of the candidate cells pass the threshold

    def run(i):
        bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in bds[i % len(bds)].items()}
        with torch.no_grad():
            return model(bd)

    for i in range(warmup):
        out = run(i)
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for i in range(steps):
        out = run(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # the same chain in static mode, captured once and replayed as one CUDA graph (chain.PlannedHotPath): inputs of every
    # step are copied into the graph's static buffers inside the timed region; counts verified after the region
    planned = None
    try:
        plan = chain.PlannedHotPath(model, batch, occ_vox_cap=batch * 20000, det_vox_cap=batch * 40000).capture()
        for i in range(warmup):
            plan(bds[i % len(bds)])
        plan.verify()
        torch.cuda.synchronize()
        p_steps = max(steps, 20)
        e0, e1 = _events()
        e0.record()
        for i in range(p_steps):
            pout = plan(bds[i % len(bds)])
        e1.record()
        torch.cuda.synchronize()
        plan.verify()
        p_ms = e0.elapsed_time(e1) / p_steps
        with torch.no_grad():
            out = run((p_steps - 1) % len(bds))          # the eager result of the batch the last replay saw
        n_ref, n_got = int(out["encoded_spconv_tensor"].features.shape[0]), int(pout["encoded_n_dev"].item())
        planned = {"value": round(batch / (p_ms * 1e-3), 2), "unit": "scenes/s", "ms_per_step": round(p_ms, 3), "steps": p_steps,
                   "bev_rows": n_got, "bev_rows_match_eager": n_ref == n_got,
                   "what": "same chain, static mode (capacity-sized buffers, device-side counts, BatchNorm + ReLU in the conv "
                           "epilogues), ONE CUDA graph per step, no host synchronisation; inference masks only (a5-a8)"}
    except Exception as e:      # noqa: BLE001 — a benchmark leg must not take the headline down
        planned = {"error": repr(e)[:300]}
    return {"value": round(batch / (ms * 1e-3), 2), "unit": "scenes/s", "ms_per_step": round(ms, 3), "scenes_per_step": batch,
            "planned": planned,
            "occ_sites": int(out["general_cls_loss_mask"].sum()), "det_voxels": int(out["voxel_coords"].shape[0]),
            "bev_rows": int(out["encoded_spconv_tensor"].features.shape[0]),
            "what": "config 3 shape: occ targets -> MeanVFE -> VoxelBackBoneDeconv -> OccHead -> PassOccVox -> OccVFE -> "
                    "VoxelBackBone8xOcc -> HeightCompression on 2 scenes of 20k points (eval mode, random-init weights, eager "
                    "spconv shim: one host read per new rulebook)"}


def stress_leg(dev, peak_gbs, reps=5):
    """Config 5: 100k-point dense clouds, voxel [0.025, 0.025, 0.05] (721 M cells per scene): voxelise + level-1 SubM rulebook +
    strided rulebook + one 16->16 SubM conv; algorithmic bytes per second against the measured HBM peak."""
    from btcdet_b200 import ops, synthetic as S
    vs, B = [0.025, 0.025, 0.05], 2
    scenes = [S.lidar_like(100000, seed=50 + b, az_density=4.0) for b in range(B)]
    pts, offs = S.batch_points(scenes)
    p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    grid = ops.voxel_grid_size(vs, S.KITTI_RANGE)
    shape = [grid[2] + 1, grid[1], grid[0]]

    def timed(fn):
        out = fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            e0, e1 = _events()
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms)), out

    ms_v, vox = timed(lambda: ops.voxelize(p, o, vs, S.KITTI_RANGE, 5, 150000, want_mean=True))
    m = int(vox[4][-1].item())
    coords, feats = vox[1][:m].contiguous(), vox[3][:m].contiguous()
    ms_s, rb_s = timed(lambda: ops.rulebook_subm(coords, B, shape, 3))
    ms_c, rb_c = timed(lambda: ops.rulebook_conv(coords, B, shape, 3, 2, 1))
    w = torch.randn(27, 16, 16, device=dev) * 0.1
    f16 = torch.randn(m, 16, device=dev)
    packed = ops.tc_pack_weight(w)
    ms_g, _ = timed(lambda: ops.sparse_conv_fwd_tc(f16, rb_s.nbr_out, packed, 16, 16))
    pairs_s = int((rb_s.nbr_out >= 0).sum())
    pairs_c = int((rb_c.nbr_out >= 0).sum())
    n_pts = int(p.shape[0])

    def row(ms, nbytes):
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"ms": round(ms, 4), "alg_bytes": int(nbytes), "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak_gbs, 4)}

    return {"points": n_pts, "l1_sites": m, "strided_sites": int(rb_c.n_out), "subm_pairs": pairs_s, "strided_pairs": pairs_c,
            "voxelize": row(ms_v, 16 * n_pts + (4 * 5 * 4 + 16) * m),
            "rulebook_subm": row(ms_s, 16 * m + 8 * pairs_s + 16 * m),
            "rulebook_strided": row(ms_c, 16 * m + 8 * pairs_c + 16 * rb_c.n_out),
            "gather_gemm_16x16": row(ms_g, 4 * m * 16 * 2 + 8 * pairs_s + 4 * 27 * 16 * 16),
            "what": "config 5 (stress): 2 x 100k-point clouds, voxel [0.025,0.025,0.05]; eager calls incl. their allocations and "
                    "the one host read of a strided rulebook; bytes = SURVEY §8(d) algorithmic bytes"}


def roi_pool_leg(dev, steps=10, warmup=3, batch=2, n_rois=128):
    """SURVEY §8(f) N1 — the native ops under `ConvHead.roi_conv_pool` (conv_head.py:247-379) at the yaml's training
    shape (2 scenes x 128 RoIs x 27 grid points, 96-cell mini grids): both stacked ball queries (4 + 3 radii, one launch
    each), the fused reverse trilinear gather from the stride-8 sparse tensor, and the three SparseConv3d(128 -> 128) of
    the mini grids + dense().  Product code only; reference-side numbers are in profiles/r2_roi_pool_report.json."""
    import spconv
    from btcdet_b200 import pointnet2_stack_cuda as p2, roi_pool, synthetic as S
    case = S.roi_head_case(batch=batch, n_points=20000, n_rois=n_rois)
    rng = np.random.default_rng(7)
    rois = case["rois"]
    g = np.stack(np.meshgrid(np.arange(3), np.arange(3), np.arange(3), indexing="ij"), -1).reshape(-1, 3)
    grid = (rois[:, :, None, :3] + ((g + 0.5) / 3.0 - 0.5)[None, None] * rois[:, :, None, 3:6]).astype(np.float32)   # un-rotated 3x3x3
    q = torch.from_numpy(np.ascontiguousarray(grid.reshape(-1, 3))).to(dev)
    qcnt = torch.full((batch,), n_rois * 27, dtype=torch.int32, device=dev)
    pts = case["points"]
    xyz = torch.from_numpy(np.ascontiguousarray(pts[:, 1:4])).to(dev)
    cnt = torch.from_numpy(np.bincount(pts[:, 0].astype(np.int64), minlength=batch).astype(np.int32)).to(dev)
    occ = torch.from_numpy(np.ascontiguousarray(case["occ_pnts"][:, :3])).to(dev)
    occ_cnt = torch.from_numpy(np.bincount(case["added_occ_b_ind"], minlength=batch).astype(np.int32)).to(dev)
    M = int(q.shape[0])
    raw_r, raw_n, occ_r, occ_n = [0.4, 0.8, 1.2, 2.4], [16, 16, 32, 64], [0.8, 1.2, 2.4], [16, 16, 32]
    raw_idx = [torch.zeros((M, n), dtype=torch.int32, device=dev) for n in raw_n]
    occ_idx = [torch.zeros((M, n), dtype=torch.int32, device=dev) for n in occ_n]
    sp = spconv.SparseConvTensor(torch.from_numpy(case["x_features"]).to(dev), torch.from_numpy(case["x_coords"]).to(dev),
                                 case["x_shape"], batch)
    lz, ly, lx = np.meshgrid(np.arange(2), np.arange(4), np.arange(12), indexing="ij")
    cell = np.stack([(lx.ravel() + 0.5) * 0.4 - 2.4, (ly.ravel() + 0.5) * 0.4 - 0.8, (lz.ravel() + 0.5) * 0.8 - 0.8], axis=1)
    cpts = torch.from_numpy((grid[:, :, :, None, :] + cell[None, None, None]).reshape(batch, -1, 3).astype(np.float32)).to(dev)
    zyx = roi_pool.target_indices(cpts, S.KITTI_RANGE, S.DET_VOXEL_SIZE, [8, 8, 8])
    per_scene = int(cpts.shape[1])
    torch.manual_seed(0)
    convs = spconv.SparseSequential(
        spconv.SparseConv3d(128, 128, [3, 3, 3], stride=[1, 1, 2], padding=[1, 1, 1], bias=False, indice_key="n1_0"),
        spconv.SparseConv3d(128, 128, [3, 3, 3], stride=[1, 2, 2], padding=[1, 1, 1], bias=False, indice_key="n1_1"),
        spconv.SparseConv3d(128, 128, [2, 2, 3], stride=[2, 2, 3], padding=[0, 0, 0], bias=False, indice_key="n1_2")).to(dev).eval()
    state = {}

    def ball():
        p2.ball_query_multi(raw_r, raw_n, q, qcnt, xyz, cnt, raw_idx)
        p2.ball_query_multi(occ_r, occ_n, q, qcnt, occ, occ_cnt, occ_idx)

    def gather():
        state["rows"] = roi_pool.trilinear_gather_rows(sp, zyx, per_scene, [2, 4, 12])

    def conv():
        c, r = state["rows"]
        t = spconv.SparseConvTensor(r, c, [2, 4, 12], batch * n_rois * 27)
        state["out"] = convs(t).dense()

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(steps):
            e0, e1 = _events()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))

    with torch.no_grad():
        ms_b, ms_g, ms_c = timed(ball), timed(gather), timed(conv)
    total = ms_b + ms_g + ms_c
    n_rows = int(state["rows"][1].shape[0])
    alg = int(zyx.shape[0]) * (12 + 32) + n_rows * (128 * 4 + 16) + int(sp.features.shape[0]) * 128 * 4
    return {"value": round(batch / (total * 1e-3), 1), "unit": "scenes/s", "ms": round(total, 4),
            "ball_query_ms": round(ms_b, 4), "trilinear_gather_ms": round(ms_g, 4), "mini_grid_convs_ms": round(ms_c, 4),
            "queries": M, "targets": int(zyx.shape[0]), "gathered_rows": n_rows, "out_shape": list(state["out"].shape),
            "trilinear_alg_bytes": alg, "trilinear_gbs": round(alg / (ms_g * 1e-3) / 1e9, 1),
            "what": "N1 ops at the yaml's training shape (2 scenes x 128 RoIs x 27 grid points): 4 + 3 radii stacked ball "
                    "queries in two launches, fused reverse trilinear gather (exact mode: one count read), three "
                    "SparseConv3d(128->128) on 6912 mini grids of [2,4,12] + dense(); eager calls incl. allocations"}
