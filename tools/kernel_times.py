"""Per-kernel device time of ONE eager backbone step (CUPTI through torch.profiler): name, launches, total us, mean us.
    python tools/kernel_times.py [--batch 16] [--no-tile-meta] > gpurun_out/kernel_times.json"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--plain", action="store_true", help="no profiler: just run eager steps (target for ncu)")
    ap.add_argument("--no-tile-meta", action="store_true")
    args = ap.parse_args()
    from btcdet_b200 import backbones, engine, synthetic as S
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    B = args.batch
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * 20000, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=5, max_voxels=16000, device=dev, use_graph=False,
                               tile_meta=not args.no_tile_meta, split_format=not args.no_split).capture()
    pts, offs = S.batch_points([S.lidar_like(20000, seed=1000 + i) for i in range(B)])
    p, o = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)
    for _ in range(3):
        plan.forward(p, o)
    torch.cuda.synchronize()
    if args.plain:
        plan.forward(p, o)
        torch.cuda.synchronize()
        print("plain run done")
        return
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        plan.forward(p, o)
        torch.cuda.synchronize()
    rows = {}
    order = []
    for ev in prof.events():
        if ev.device_type is None or "cuda" not in str(ev.device_type).lower():
            continue
        name = ev.name.split("(")[0][:70]
        if name not in rows:
            rows[name] = [0, 0.0]
            order.append(name)
        rows[name][0] += 1
        rows[name][1] += float(ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total)
    out = [{"kernel": n, "launches": rows[n][0], "total_us": round(rows[n][1], 1), "mean_us": round(rows[n][1] / rows[n][0], 2)}
           for n in order]
    out.sort(key=lambda r: -r["total_us"])
    print(json.dumps({"batch": B, "tile_meta": not args.no_tile_meta, "sum_us": round(sum(r["total_us"] for r in out), 1),
                      "kernels": out}, indent=1))


if __name__ == "__main__":
    main()
