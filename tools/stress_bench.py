#!/usr/bin/env python
"""BASELINE.json configs[4] (stress): dense 100k-point clouds at voxel [0.025,0.025,0.05] — HBM-side view of the
indexing kernels (voxelize, hash/bitmap index, rulebooks) and of one gather-GEMM layer per level.

Prints one JSON object: per-stage CUDA-event time, algorithmic bytes and achieved GB/s against the
measured HBM peak.  Scenes are batched (--batch) so the working set exceeds L2 (126 MB) and DRAM is actually exercised.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from btcdet_b200 import ops, synthetic as S  # noqa: E402


def timed(fn, reps, flush):
    out = None
    ms = []
    for r in range(reps + 1):
        flush.fill_(float(r))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        if r:
            ms.append(e0.elapsed_time(e1))
    return out, float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    dev = torch.device("cuda")
    B = args.batch
    vs, rng = [0.025, 0.025, 0.05], S.KITTI_RANGE
    grid = ops.voxel_grid_size(vs, rng)
    sparse_shape = [grid[2] + 1, grid[1], grid[0]]
    scenes = [S.lidar_like(args.points, seed=50 + b, az_density=1.5, point_range=[0, -20, -3, 35.2, 20, 1]) for b in range(B)]
    pts, offs = S.batch_points(scenes)
    pts_d, offs_d = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    res = {"config": {"workload": "configs[4] stress: %d x %d-pt dense clouds, voxel %s" % (B, args.points, vs),
                      "sparse_shape": sparse_shape}, "hbm_peak_gbs": peak, "stages": []}

    def add(name, ms, alg_bytes, extra=None):
        row = {"stage": name, "ms": round(ms, 4), "alg_MB": round(alg_bytes / 1e6, 2),
               "GBps": round(alg_bytes / (ms * 1e-3) / 1e9, 1), "frac_of_hbm_peak": round(alg_bytes / (ms * 1e-3) / 1e9 / peak, 4)}
        if extra:
            row.update(extra)
        res["stages"].append(row)

    max_vox = 150000
    (v, c, npnt, mean, nv), ms = timed(lambda: ops.voxelize(pts_d, offs_d, vs, rng, 5, max_vox, want_mean=True, grid=grid),
                                       args.reps, flush)
    m = int(nv[-1].item())
    add("voxelize (+MeanVFE)", ms, 16 * pts.shape[0] + (4 * 5 * 4 + 16 + 16) * m, {"voxels": m})
    coords = c[:m].contiguous()
    idx, ms = timed(lambda: ops.build_hash(coords, B, sparse_shape), args.reps, flush)
    add("coordinate hash build", ms, 16 * m + 12 * idx.hash_keys.numel())
    rb1, ms = timed(lambda: ops.rulebook_subm(coords, B, sparse_shape, 3, index=idx), args.reps, flush)
    p1 = int((rb1.nbr_out >= 0).sum())
    add("subm rulebook L1 (hash probes)", ms, 16 * m + 4 * 27 * m, {"pairs": p1})
    rb2, ms = timed(lambda: ops.rulebook_conv(coords, B, sparse_shape, 3, 2, 1), args.reps, flush)
    p2 = int((rb2.nbr_out >= 0).sum())
    add("strided-conv rulebook L1->L2 (bitmap mark+scan+emit+tables)", ms,
        16 * m + 8 * p2 + 16 * rb2.n_out + 2 * 8 * rb2.out_index.entries.numel(), {"out_sites": rb2.n_out, "pairs": p2,
                                                                                 "bitmap_MB": round(8 * rb2.out_index.entries.numel() / 1e6, 1)})
    rb3, ms = timed(lambda: ops.rulebook_subm(rb2.out_coords, B, rb2.out_shape, 3, index=rb2.out_index), args.reps, flush)
    p3 = int((rb3.nbr_out >= 0).sum())
    add("subm rulebook L2 (bitmap probes)", ms, 16 * rb2.n_out + 4 * 27 * rb2.n_out, {"pairs": p3})
    res["total_active_sites"] = m + rb2.n_out
    g = torch.Generator(device="cpu").manual_seed(0)
    for name, rb, cin, cout, n_in in (("conv L1 subm 16->16 (ffma)", rb1, 16, 16, m), ("conv L1->L2 16->32 (ffma)", rb2, 16, 32, m),
                                      ("conv L2 subm 32->32 (tcgen05)", rb3, 32, 32, rb2.n_out)):
        f = torch.randn(n_in, cin, generator=g).to(dev)
        w = (torch.randn(27, cin, cout, generator=g) * 0.1).to(dev)
        pairs = int((rb.nbr_out >= 0).sum())
        n_out = rb.nbr_out.shape[0]
        if "tcgen05" in name:
            pk = ops.tc_pack_weight(w)
            _, ms = timed(lambda: ops.sparse_conv_fwd_tc(f, rb.nbr_out, pk, cin, cout), args.reps, flush)
        else:
            _, ms = timed(lambda: ops.sparse_conv_fwd(f, rb.nbr_out, w, algo=1), args.reps, flush)
        add(name, ms, 4 * n_in * cin + 4 * n_out * cout + 8 * pairs + 4 * 27 * cin * cout,
            {"pairs": pairs, "tflops": round(2 * pairs * cin * cout / (ms * 1e-3) / 1e12, 2)})
    print(json.dumps(res))


if __name__ == "__main__":
    main()
