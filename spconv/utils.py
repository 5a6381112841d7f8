"""spconv.utils.VoxelGeneratorV2 / VoxelGenerator (spconv 1.2.1 `spconv/utils/__init__.py`).

Reference call sites: btcdet/datasets/processor/data_processor.py:64-73,85 / :108-117,136 /
:161-170,177 (`VoxelGenerator(voxel_size=, point_cloud_range=, max_num_points=, max_voxels=)`,
`.generate(points)` -> dict with `voxels`, `coordinates`, `num_points_per_voxel`).

The grouping runs on the GPU (libbtcdet_b200.so `btc_voxelize`), bit-exact with the sequential
first-come CPU loop of spconv.  numpy in -> numpy out (H2D/D2H around the kernel), CUDA tensor in
-> CUDA tensors out (no host traffic).  There is deliberately no CPU implementation here: a
DataLoader that calls this from worker processes must use the `spawn` start method, or — better —
voxelise on the training process' GPU after collation.
"""
import numpy as np
import torch

from btcdet_b200 import ops as _ops


class VoxelGeneratorV2:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, full_mean=False,
                 block_filtering=False, block_factor=8, block_size=3, height_threshold=0.1,
                 height_high_threshold=2.0):
        assert full_mean is False, "full_mean is not supported"
        assert block_filtering is False, "block_filtering is not supported"
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = int(max_num_points)
        self._max_voxels = int(max_voxels)
        self._grid_size = grid_size
        self._full_mean = full_mean

    def generate(self, points, max_voxels=None):
        """points: [N, C>=3] float32 numpy array or CUDA tensor."""
        max_voxels = int(max_voxels or self._max_voxels)
        is_numpy = isinstance(points, np.ndarray)
        if is_numpy:
            pts = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float32)).cuda(non_blocking=True)
        else:
            pts = points.to(dtype=torch.float32).contiguous()
            if not pts.is_cuda:
                raise RuntimeError("VoxelGeneratorV2 (btcdet_b200) needs a CUDA tensor or a numpy array")
        n = pts.shape[0]
        offsets = torch.tensor([0, n], dtype=torch.int32, device=pts.device)
        voxels, coords, num_points, _, n_voxels = _ops.voxelize(
            pts, offsets, self._voxel_size.tolist(), self._point_cloud_range.tolist(), self._max_num_points,
            max_voxels, want_mean=False, grid=[int(g) for g in self._grid_size])
        m = int(n_voxels[1].item())
        voxels, coords, num_points = voxels[:m], coords[:m, 1:], num_points[:m]
        mask = (torch.arange(self._max_num_points, device=pts.device).view(1, -1) < num_points.view(-1, 1))
        res = {
            "voxels": voxels,
            "coordinates": coords.contiguous(),
            "num_points_per_voxel": num_points,
            "voxel_point_mask": mask.view(m, self._max_num_points, 1).to(voxels.dtype),
        }
        if is_numpy:
            res = {k: v.cpu().numpy() for k, v in res.items()}
        res["voxel_num"] = m
        return res

    def generate_multi_gpu(self, points, max_voxels=None):
        return self.generate(points, max_voxels)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size


class VoxelGenerator(VoxelGeneratorV2):
    """spconv's older generator returns a 3-tuple (data_processor.py:90 handles both)."""

    def generate(self, points, max_voxels=None):
        res = super().generate(points, max_voxels)
        return res["voxels"], res["coordinates"], res["num_points_per_voxel"]
