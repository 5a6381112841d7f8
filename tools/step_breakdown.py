#!/usr/bin/env python
"""Per-stage CUDA-event breakdown of one bench.py step (configs[1], same plan as bench.py).

Times, eagerly on one stream and warm (the step's working set is L2-resident in the real step as well):
every rulebook-chain step, every conv layer, then the two chains as wholes and the captured graph.
Answers "which chain is the critical path of the two-stream graph".  Prints one JSON object.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from btcdet_b200 import backbones, engine, synthetic as S  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--sort", action="store_true")
    ap.add_argument("--algo", type=int, default=0)
    ap.add_argument("--side-priority", type=int, default=0, help="priority of the rulebook stream (-1 high, 0 default)")
    ap.add_argument("--diag", action="store_true", help="also time the conv chain with parts of the tcgen05 tile skipped "
                    "(btc_sparse_conv_tc_diag masks; wrong results, timing only)")
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--no-tile-meta", action="store_true", help="A/B: without the per-rulebook tile masks / heaviest-first order")
    ap.add_argument("--grids", default="", help="comma-separated caps on the conv grid: only the captured graph is timed per cap")
    ap.add_argument("--variants", default="16,0,1", help="semicolon-separated tcgen05 tile variants npw,cat,dyn "
                    "(one JSON line each), e.g. '16,1,1;16,0,1;16,1,0'")
    ap.add_argument("--early", default="", help="comma-separated L:G pairs: first L convolutions capped at G CTAs (captured graph only)")
    args = ap.parse_args()
    if args.early:
        early_grid_sweep(args)
        return
    if args.grids:
        grid_sweep(args)
        return
    for v in args.variants.split(";"):
        npw, cat, dyn = (int(x) for x in v.split(","))
        run(args, npw, cat, dyn)


def early_grid_sweep(args):
    """Captured-graph time with the first L convolutions capped at G CTAs (engine.BackbonePlan(early_conv_grid=(L, G)))."""
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N = args.batch, 20000
    pts, offs = S.batch_points([S.lidar_like(N, seed=i) for i in range(B)])
    pts_d, offs_d = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    out = []
    for spec in [None] + [tuple(int(x) for x in v.split(":")) for v in args.early.split(",")]:
        torch.manual_seed(0)
        model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
        plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * N, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                                   max_points=S.DET_MAX_POINTS, max_voxels=S.DET_MAX_VOXELS["train"], device=dev,
                                   early_conv_grid=spec).capture()
        plan.load_points(pts_d, offs_d)
        ms = []
        for r in range(args.reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.step()
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ms.append(e0.elapsed_time(e1))
        out.append({"early_conv_grid": spec, "graph_us": round(float(np.median(ms)) * 1e3, 1)})
        del plan
        torch.cuda.empty_cache()
    from btcdet_b200 import _lib
    _lib.load().btc_sparse_conv_tc_grid(148)
    print(json.dumps({"batch": B, "early_grid_sweep": out}), flush=True)


def grid_sweep(args):
    """Captured-graph time of the whole step for several caps on the persistent conv grid."""
    from btcdet_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N = args.batch, 20000
    pts, offs = S.batch_points([S.lidar_like(N, seed=i) for i in range(B)])
    pts_d, offs_d = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    out = []
    for g in [int(x) for x in args.grids.split(",")]:
        lib.btc_sparse_conv_tc_grid(g)
        torch.manual_seed(0)
        model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
        plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * N, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                                   max_points=S.DET_MAX_POINTS, max_voxels=S.DET_MAX_VOXELS["train"], algo=args.algo,
                                   device=dev, use_graph=True, sort_rows=args.sort, side_priority=args.side_priority,
                               tile_meta=not args.no_tile_meta, split_format=not args.no_split).capture()
        plan.load_points(pts_d, offs_d)
        ms = []
        for r in range(args.reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.step()
            e1.record()
            torch.cuda.synchronize()
            if r >= 3:
                ms.append(e0.elapsed_time(e1))
        out.append({"conv_grid": g, "graph_us": round(float(np.median(ms)) * 1e3, 1)})
        del plan
        torch.cuda.empty_cache()
    lib.btc_sparse_conv_tc_grid(148)
    print(json.dumps({"batch": B, "grid_sweep": out}), flush=True)


def run(args, npw, cat, dyn):
    from btcdet_b200 import ops
    ops.tc_config(npw, cat, dyn)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N = args.batch, 20000
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * N, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=S.DET_MAX_POINTS, max_voxels=S.DET_MAX_VOXELS["train"], algo=args.algo,
                               device=dev, use_graph=True, sort_rows=args.sort, side_priority=args.side_priority,
                               tile_meta=not args.no_tile_meta, split_format=not args.no_split).capture()
    pts, offs = S.batch_points([S.lidar_like(N, seed=i) for i in range(B)])
    plan.load_points(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev))
    plan.step()
    torch.cuda.synchronize()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(fn):
        ms = []
        for r in range(args.reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ms.append(e0.elapsed_time(e1))
        return round(float(np.median(ms)) * 1e3, 2)   # microseconds

    rows = [{"step": "voxelize(+MeanVFE)", "us": timed(lambda: plan.launch_voxelize(st))}]
    for s in plan.steps:
        if s.kind == "conv":
            a = s.args
            name = "conv %d->%d K=%d" % (a[10], a[11], a[9])
            rows.append({"step": name, "us": timed(lambda: plan.launch_conv(s.args, st))})
        else:
            name = s.kind
            if s.kind == "conv_rb":
                name += " -> %s" % (s.args[1].shape,)
            elif s.kind in ("subm_rb", "hash_build"):
                name += " @ %s" % (s.args[0].shape,)
            else:
                name += " K=%d cap=%d" % (s.args[0].shape[1], s.args[0].shape[0])
            rows.append({"step": name, "us": timed(lambda: plan.launch_index_step(s, st))})

    def index_chain():
        plan.launch_voxelize(st)
        for s in plan.steps:
            if s.kind != "conv":
                plan.launch_index_step(s, st)

    def conv_chain():
        for s in plan.steps:
            if s.kind == "conv":
                plan.launch_conv(s.args, st)

    res = {"batch": B, "sorted_rows": args.sort, "tile_meta": not args.no_tile_meta, "variant": {"producer_warps": npw, "concat_b": cat, "dynamic_tiles": dyn},
           "steps": rows,
           "sum_index_us": round(sum(r["us"] for r in rows if not r["step"].startswith("conv ")), 1),
           "sum_conv_us": round(sum(r["us"] for r in rows if r["step"].startswith("conv ")), 1),
           "index_chain_us": timed(index_chain), "conv_chain_us": timed(conv_chain),
           "graph_us": timed(plan.step), "counts": plan.read_counts()}
    if args.diag:
        from btcdet_b200 import _lib
        lib = _lib.load()
        names = {0: "full", 1: "no gather", 2: "no readback/split/tmem-store", 3: "no gather, no split", 4: "1 of 3 MMAs",
                 5: "no gather, 1 MMA", 6: "no split, 1 MMA", 7: "barriers + 1 MMA only"}
        convs = [s for s in plan.steps if s.kind == "conv"]
        diag = []
        names.update({8: "no weight tiles", 15: "barriers + 1 MMA, no weight tiles", 16: "full, one stage per issue trip",
                      31: "skeleton, one stage per issue trip"})
        for mask in [0, 16, 15, 31, 7, 3, 4]:
            lib.btc_sparse_conv_tc_diag(mask)
            diag.append({"mask": mask, "what": names[mask], "conv_chain_us": timed(conv_chain),
                         "conv32_us": timed(lambda: plan.launch_conv(convs[3].args, st)),
                         "conv64_us": timed(lambda: plan.launch_conv(convs[6].args, st))})
        lib.btc_sparse_conv_tc_diag(0)
        res["diag"] = diag
    print(json.dumps(res), flush=True)
    del plan
    torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
