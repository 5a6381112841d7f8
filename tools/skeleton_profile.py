#!/usr/bin/env python
"""Target for one `ncu --set full --import-source on` capture of the tcgen05 tile in a timing-diagnostic mode
(default mask 15: no gather, no split / TMEM stores, one MMA of three, no weight copies): the warp-state samples of the
remaining skeleton show which hand-off the issuing warp spends its time on.

  ncu --set full --clock-control none --import-source on -k regex:conv_fwd_tc --launch-skip 25 --launch-count 1 \\
      -o gpurun_out/prof_skeleton python tools/skeleton_profile.py
(24 tcgen05 launches precede the three diagnostic ones: two eager passes over the 12 layers.)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from btcdet_b200 import _lib, backbones, engine, synthetic as S  # noqa: E402


def main():
    mask = int(sys.argv[1]) if len(sys.argv) > 1 else 15
    layer = int(sys.argv[2]) if len(sys.argv) > 2 else 6          # 6 = first SubM 64->64 of level 3
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B, N = 16, 20000
    torch.manual_seed(0)
    model = backbones.randomize_bn_(backbones.VoxelBackBone8x(4)).eval()
    plan = engine.BackbonePlan(model.layer_specs(), model.sparse_shape, B, B * N, S.DET_VOXEL_SIZE, S.KITTI_RANGE,
                               max_points=S.DET_MAX_POINTS, max_voxels=S.DET_MAX_VOXELS["train"], device=dev,
                               use_graph=False).capture()
    pts, offs = S.batch_points([S.lidar_like(N, seed=i) for i in range(B)])
    plan.load_points(torch.from_numpy(pts).to(dev), torch.from_numpy(offs).to(dev))
    plan.step()
    torch.cuda.synchronize()
    lib = _lib.load()
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    conv = [s for s in plan.steps if s.kind == "conv"][layer]
    lib.btc_sparse_conv_tc_diag(mask)
    for _ in range(3):
        plan.launch_conv(conv.args, st)
    torch.cuda.synchronize()
    lib.btc_sparse_conv_tc_diag(0)
    print("done: mask %d, layer %d (%d->%d)" % (mask, layer, conv.args[10], conv.args[11]))


if __name__ == "__main__":
    main()
