"""Generates tests/golden/roi_pool_ref.npz by running the REFERENCE's own RoI-head code on the CPU:
`ConvHead` built from the reference's btcdet_kitti_car.yaml (btcdet/models/roi_heads/conv_head.py, unchanged, executed
through tests/golden/ref_loader.py) -> get_global_grid_points_of_roi -> create_local_conv_grid ->
interpolate_from_3d_features (which calls common_utils.reverse_sparse_trilinear_interpolate_torch) on a seeded
synthetic case.  The fixture carries the inputs (RoIs, sparse source rows), the reference's intermediate grid points and
its outputs, so that the oracle — and the kernel source under the warp emulation — stay pinned to the reference where
no checkout exists.  Run from the repo root in the build container:  python tests/golden/make_roi_pool_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_loader  # noqa: E402
from btcdet_b200 import synthetic as S  # noqa: E402


def main():
    mods = ref_loader.load_roi_head_modules()
    head = ref_loader.build_conv_head(mods, S.DET_VOXEL_SIZE, S.KITTI_RANGE)
    case = S.roi_head_case(batch=2, n_points=6000, n_rois=4, n_occ=50, channels=8, seed=3)
    rois = torch.from_numpy(case["rois"])
    feats, coords, shape = torch.from_numpy(case["x_features"]), torch.from_numpy(case["x_coords"]), case["x_shape"]

    class Feat(object):                       # what the reference touches of a SparseConvTensor: dense() = index assignment
        spatial_shape = shape
        indices = coords

        @staticmethod
        def dense():
            im = torch.zeros((2, shape[0], shape[1], shape[2], feats.shape[1]))
            c = coords.long()
            im[c[:, 0], c[:, 1], c[:, 2], c[:, 3]] = feats
            return im.permute(0, 4, 1, 2, 3).contiguous()

    with ref_loader.cuda_as_cpu():
        grid_pts, _ = head.get_global_grid_points_of_roi(rois, grid_size=head.grid_size, e2e=False, dim_times=head.dim_times)
        sm = head.size_map["x_combine"]
        conv_pts, dense_idx = head.create_local_conv_grid(grid_pts.view(-1, 3), rois, sm["local_grid_size"], sm["dims"],
                                                          sm["scene_times"])
        out_c, out_f = head.interpolate_from_3d_features(conv_pts, dense_idx, Feat, head.downsample_times_map["x_combine"])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roi_pool_ref.npz")
    np.savez_compressed(out, rois=case["rois"], x_features=case["x_features"], x_coords=case["x_coords"],
                        x_shape=np.array(shape), grid_points=grid_pts.numpy(), conv_grid_points=conv_pts.numpy(),
                        local_grid_size=np.array(sm["local_grid_size"]), stride=np.array(head.downsample_times_map["x_combine"]),
                        out_coords=out_c.numpy(), out_features=out_f.numpy())
    print("wrote %s: %d targets -> %d rows (%d KB)" % (out, conv_pts.shape[0] * conv_pts.shape[1], out_f.shape[0],
                                                       os.path.getsize(out) // 1024))


if __name__ == "__main__":
    main()
