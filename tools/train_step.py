#!/usr/bin/env python
"""BASELINE.json configs[3] shape: data-parallel TRAINING step of the sparse backbone — forward + backward through the
CUDA library (dX, dW, db kernels), train-mode BatchNorm, Adam, DDP gradient all-reduce over NCCL — on synthetic
20k-point scenes, 2 scenes per GPU as in the reference config (tools/cfgs/model_configs/btcdet_kitti_car.yaml:332).

  python tools/train_step.py --steps 10                                  # 1 GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py --steps 10
Prints one JSON line on rank 0 (scenes/s = all ranks' scenes / max-over-ranks device time).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from btcdet_b200 import backbones, dist as bd, ops, synthetic as S  # noqa: E402


class Net(nn.Module):
    """VoxelBackBone8x + HeightCompression-style dense() + a 1x1 BEV head standing in for the RPN."""

    def __init__(self):
        super().__init__()
        self.backbone = backbones.VoxelBackBone8x(4)
        self.head = nn.Conv2d(256, 8, 1)

    def forward(self, feats, coords, batch):
        out = self.backbone({"voxel_features": feats, "voxel_coords": coords, "batch_size": batch})["encoded_spconv_tensor"]
        d = out.dense()                                  # [B, 128, 2, 200, 176]
        n, c, dd, h, w = d.shape
        return self.head(d.view(n, c * dd, h, w))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    rank, world = bd.init()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    net = Net().to(dev).train()
    model = nn.parallel.DistributedDataParallel(net, device_ids=[dev.index]) if world > 1 else net
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    batches = []
    for i in range(4):
        scenes = [S.lidar_like(20000, seed=7000 + 100 * rank + 10 * i + b) for b in range(args.batch)]
        pts, offs = S.batch_points(scenes)
        batches.append((torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev)))

    def step(i):
        pts, offs = batches[i % len(batches)]
        v, c, n, mean, nv = ops.voxelize(pts, offs, S.DET_VOXEL_SIZE, S.KITTI_RANGE, 5, 16000, want_mean=True)
        m = int(nv[-1].item())
        pred = model(mean[:m], c[:m], args.batch)
        loss = pred.square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()                                   # DDP: bucketed NCCL all-reduce overlaps backward
        torch.nn.utils.clip_grad_norm_(model.parameters(), 10.0)
        opt.step()
        return loss

    for i in range(args.warmup):
        step(i)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = bd.max_over_ranks(e0.elapsed_time(e1), device=dev)
    if rank == 0:
        print(json.dumps({"metric": "training scenes/sec (fwd+bwd+allreduce+Adam, VoxelBackBone8x + BEV head)",
                          "value": round(world * args.batch * args.steps / (ms * 1e-3), 2), "unit": "scenes/s", "n_gpus": world,
                          "steps": args.steps, "ms_per_step": round(ms / args.steps, 3), "scenes_per_gpu": args.batch,
                          "loss": float(loss.detach()), "path": "eager spconv shim (autograd Functions over the C ABI)"}))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
