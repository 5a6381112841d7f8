"""Per-CTA timeline of the tcgen05 gather-GEMM launches of one backbone step (btc_sparse_conv_tc_trace).

For every conv layer of the planned VoxelBackBone8x step (batch 16, eager launches, L2 flushed before each) prints where a
launch's time goes: kernel entry -> set-up done -> first gather issued -> first operand stage full -> first MMA issued ->
last accumulator committed -> last epilogue done -> exit, as medians / maxima over the 148 CTAs, plus the cycles the MMA
issuer spent waiting for operand stages, for chunk lists and for a free accumulator.

    python tools/tc_timeline.py [--batch 16] > gpurun_out/tc_timeline.json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--no-split", action="store_true")
    ap.add_argument("--no-tile-meta", action="store_true")
    args = ap.parse_args()
    from btcdet_b200 import _lib, backbones, engine, synthetic as S
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    B = args.batch
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * 20000, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=5, max_voxels=16000, device=dev, use_graph=False, tile_meta=not args.no_tile_meta, split_format=not args.no_split).capture()
    pts, offs = S.batch_points([S.lidar_like(20000, seed=1000 + i) for i in range(B)])
    plan.forward(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev))
    torch.cuda.synchronize()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    trace = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    clk_ghz = None
    out = []
    for s in plan.steps:
        if s.kind != "conv":
            continue
        fin, nbr, w, bias, scale, shift, relu, fout, lout, K, cin, cout, packed, rows, meta, fmt = s.args
        if packed is None:
            continue
        rec = {"layer": "%d->%d K=%d" % (cin, cout, K), "rows": int(lout.n_dev.item())}
        for warm in (True, False):
            flush.fill_(1.0)
            trace.zero_()
            _lib.check(lib.btc_sparse_conv_tc_trace(ctypes.c_void_p(trace.data_ptr())), "trace on")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.launch_conv(s.args, st)
            e1.record()
            torch.cuda.synchronize()
            _lib.check(lib.btc_sparse_conv_tc_trace(None), "trace off")
        t = trace.cpu().numpy().reshape(148, 32).astype(np.float64)
        live = t[:, 1] > 0
        t = t[live]
        wall_ns = t[:, 10].max() - t[:, 0].min()
        cyc = t[:, 9] - t[:, 1]
        ghz = float(np.median(cyc / np.maximum(t[:, 10] - t[:, 0], 1)))
        us = lambda c: c / ghz / 1e3   # noqa: E731
        rel = lambda i: us(t[:, i] - t[:, 1])   # noqa: E731
        rec.update({
            "event_us": round(e0.elapsed_time(e1) * 1e3, 1), "ctas": int(live.sum()), "sm_ghz": round(ghz, 3),
            "first_entry_to_last_exit_us": round(wall_ns / 1e3, 1),
            "entry_skew_us": round((t[:, 0].max() - t[:, 0].min()) / 1e3, 1),
            "exit_skew_us": round((t[:, 10].max() - t[:, 10].min()) / 1e3, 1),
            "cta_lifetime_us_med_max": [round(float(np.median(us(cyc))), 1), round(float(us(cyc).max()), 1)],
            "setup_done_us_med": round(float(np.median(rel(2))), 2),
            "first_gather_us_med": round(float(np.median(rel(4))), 2),
            "first_stage_full_us_med": round(float(np.median(rel(5))), 2),
            "first_mma_us_med": round(float(np.median(rel(6))), 2),
            "last_commit_us_med": round(float(np.median(rel(7))), 2),
            "last_epilogue_us_med": round(float(np.median(rel(8))), 2),
            "tiles_per_cta_med_max": [float(np.median(t[:, 11])), float(t[:, 11].max())],
            "stages_per_cta_med": float(np.median(t[:, 13])),
            "issuer_wait_stage_us_med": round(float(np.median(us(t[:, 12]))), 1),
            "issuer_wait_list_us_med": round(float(np.median(us(t[:, 3]))), 1),
            "issuer_wait_accumulator_us_med": round(float(np.median(us(t[:, 15]))), 1),
            "producer0_wait_empty_us_med": round(float(np.median(us(t[:, 14]))), 1),
            "index_loader_us_med": {"wait_free_buffer": round(float(np.median(us(t[:, 16]))), 1),
                                    "tile_fetch": round(float(np.median(us(t[:, 17]))), 1),
                                    "index_copy_and_mask": round(float(np.median(us(t[:, 18]))), 1),
                                    "chunk_list": round(float(np.median(us(t[:, 19]))), 1)},
            "producer0_us_med": {"locate_and_issue": round(float(np.median(us(t[:, 22]))), 1),
                                 "cp_async_wait_readback": round(float(np.median(us(t[:, 23]))), 1),
                                 "slot_wait_tmem_store": round(float(np.median(us(t[:, 24]))), 1)},
            "epilogue_us_med": {"wait_accumulator": round(float(np.median(us(t[:, 20]))), 1),
                                "busy": round(float(np.median(us(t[:, 21]))), 1)},
            "cycles_per_stage_med": round(float(np.median((t[:, 7] - t[:, 6]) / np.maximum(t[:, 13], 1))), 1),
        })
        out.append(rec)
    print(json.dumps({"batch": B, "tile_meta": not args.no_tile_meta, "layers": out}, indent=1))


if __name__ == "__main__":
    main()
