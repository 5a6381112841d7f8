"""Where the time of one config-4-shaped training step goes (tools/bench_legs.train_leg's step on one GPU): wall time per
step, summed device time, launches, host synchronisations and the top kernels (torch.profiler).

    python tools/train_profile.py > gpurun_out/train_profile.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402


def main():
    import bench_legs
    from btcdet_b200 import ops, synthetic as S
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = bench_legs._TrainNet().to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    batch = 2
    sc = [S.lidar_like(20000, seed=7000 + b) for b in range(batch)]
    pts, offs = S.batch_points(sc)
    pts, offs = torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)

    def step():
        v, c, n, mean, nv = ops.voxelize(pts, offs, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000, want_mean=True)
        m = int(nv[-1].item())
        loss = net(mean[:m], c[:m], batch).square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        return loss

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall_ms = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    kern, order, host, launches = {}, [], {}, 0
    for ev in prof.events():
        if "cuda" in str(ev.device_type).lower():
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
            if n in ("cudaStreamSynchronize", "cudaDeviceSynchronize", "aten::item", "cudaMemcpyAsync", "aten::nonzero"):
                host[n] = host.get(n, 0) + 1
    rows = [{"kernel": n, "launches": kern[n][0], "total_us": round(kern[n][1], 1)} for n in order]
    rows.sort(key=lambda r: -r["total_us"])
    dev_us = sum(r["total_us"] for r in rows)
    print(json.dumps({"wall_ms_per_step": round(wall_ms, 3), "device_ms_per_step": round(dev_us / 1e3, 3),
                      "kernel_launches": launches, "device_ops": sum(r["launches"] for r in rows), "host_calls": host,
                      "top": rows[:30]}, indent=1))


if __name__ == "__main__":
    main()
