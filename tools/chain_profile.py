"""Where the time of one config-3 chain step goes (btcdet_b200.chain.BtcHotPath, eager spconv shim, 2 scenes).

Runs the step under torch.profiler (CPU + CUDA activities) and prints: wall time per step (CUDA events), summed device
time of all kernels / copies, number of launches, number of host synchronisations (cudaStreamSynchronize,
cudaDeviceSynchronize, blocking device->host copies) and the top kernels — i.e. how much of the step is host-side
launch / synchronisation overhead rather than device work.

    python tools/chain_profile.py [--batch 2] > gpurun_out/chain_profile.json
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--planned", action="store_true", help="profile chain.PlannedHotPath (sync-free, one CUDA graph) instead")
    args = ap.parse_args()
    from btcdet_b200 import backbones, chain
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = chain.BtcHotPath()
    backbones.randomize_bn_(model, 0)
    model = model.to(dev).eval()
    bds = [chain.synthetic_batch([900 + 10 * i + b for b in range(args.batch)], n_points=20000, device=dev, with_rot=True,
                                 mode="test") for i in range(2)]
    chain.calibrate_occ_head_bias(model, bds[0], 0.03)

    def run(i):
        bd = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in bds[i % 2].items()}
        with torch.no_grad():
            return model(bd)

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1) / 5
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        run(0)
        torch.cuda.synchronize()
    kern, order, syncs, launches = {}, [], {}, 0
    for ev in prof.events():
        dt = str(ev.device_type).lower()
        if "cuda" in dt:
            name = ev.name.split("(")[0][:72]
            if name not in kern:
                kern[name] = [0, 0.0]
                order.append(name)
            kern[name][0] += 1
            kern[name][1] += float(ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total)
        else:
            n = ev.name
            if n in ("cudaLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernel", "cuLaunchKernelEx"):
                launches += 1
            if n in ("cudaStreamSynchronize", "cudaDeviceSynchronize", "cudaEventSynchronize", "aten::item",
                     "aten::_local_scalar_dense", "cudaMemcpyAsync", "cudaMemcpy", "aten::nonzero", "cudaMalloc"):
                syncs[n] = syncs.get(n, 0) + 1
    rows = [{"kernel": n, "launches": kern[n][0], "total_us": round(kern[n][1], 1)} for n in order]
    rows.sort(key=lambda r: -r["total_us"])
    dev_us = sum(r["total_us"] for r in rows)
    print(json.dumps({"batch": args.batch, "wall_ms_per_step": round(wall_ms, 3), "device_ms_per_step": round(dev_us / 1e3, 3),
                      "device_fraction_of_wall": round(dev_us / 1e3 / wall_ms, 3), "kernel_launches": launches,
                      "device_ops": sum(r["launches"] for r in rows), "host_calls": syncs, "top": rows[:40]}, indent=1))


if __name__ == "__main__":
    main()
