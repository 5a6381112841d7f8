"""Dataset-side rows of the hot path through the REFERENCE'S OWN DataProcessor and collate_batch.

* a1 / a2: `DataProcessor.transform_points_to_sphere_voxels` and `.det_transform_points_to_voxels`
  (btcdet/datasets/processor/data_processor.py:105-190) run unchanged on the drop-in `spconv.utils.VoxelGeneratorV2`
  (GPU grouping behind numpy in / numpy out) — in the main process and inside DataLoader WORKER processes started with
  the `spawn` method (the DataLoader story of INTEGRATION.md: forked workers cannot use CUDA, spawned ones can) — and
  match the sequential CPU oracle.
* a3 / a4: `DatasetTemplate.collate_batch` (btcdet/datasets/dataset.py:168-223) + `load_data_to_gpu`
  (btcdet/models/__init__.py:16-22: every array -> float32 CUDA tensor) produce exactly what the GPU-resident input
  pipeline (ops.voxelize_occ_and_det on the concatenated raw points, batch index from scene_offsets) produces.
Needs the staged reference sources (oracle/stage_reference.py)."""
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import ref_loader  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not staged (oracle/stage_reference.py)")


def _reference_dataset_modules():
    mods = ref_loader.load_reference_modules()
    R = ref_loader.REF
    ref_loader._ns("btcdet.datasets", os.path.join(R, "btcdet/datasets"))
    ref_loader._ns("btcdet.datasets.processor", os.path.join(R, "btcdet/datasets/processor"))
    ref_loader._ns("btcdet.datasets.augmentor", os.path.join(R, "btcdet/datasets/augmentor"))
    ref_loader._load("btcdet.utils.box_utils", "btcdet/utils/box_utils.py")
    dp = ref_loader._load("btcdet.datasets.processor.data_processor", "btcdet/datasets/processor/data_processor.py")
    # dataset.py imports the augmentors (database sampler -> compiled iou3d extension): stub the module, collate_batch
    # itself is plain numpy
    aug = types.ModuleType("btcdet.datasets.augmentor.data_augmentor")
    aug.DataAugmentor = object
    sys.modules[aug.__name__] = aug
    ref_loader._load("btcdet.datasets.processor.point_feature_encoder", "btcdet/datasets/processor/point_feature_encoder.py")
    ds = ref_loader._load("btcdet.datasets.dataset", "btcdet/datasets/dataset.py")
    return dp, ds


def make_processor(training=False):
    from btcdet_b200 import synthetic as S
    dp, _ = _reference_dataset_modules()
    Cfg = ref_loader.Cfg
    queue = [Cfg.wrap({"NAME": "transform_points_to_sphere_voxels", "VOXEL_SIZE": S.OCC_VOXEL_SIZE,
                       "MAX_POINTS_PER_VOXEL": S.OCC_MAX_POINTS, "MAX_NUMBER_OF_VOXELS": S.OCC_MAX_VOXELS}),
             Cfg.wrap({"NAME": "det_transform_points_to_voxels", "VOXEL_SIZE": S.DET_VOXEL_SIZE,
                       "MAX_POINTS_PER_VOXEL": S.DET_MAX_POINTS, "MAX_NUMBER_OF_VOXELS": S.DET_MAX_VOXELS})]
    return dp.DataProcessor(queue, np.array(S.OCC_RANGE, np.float32), training,
                            occ_config=Cfg.wrap({"COORD_TYPE": "cylinder"}),
                            det_point_cloud_range=np.array(S.KITTI_RANGE, np.float32))


class _Scenes(torch.utils.data.Dataset):
    """__getitem__ = the reference's data_processor.forward on a synthetic scene (what KittiDataset.prepare_data ends with)."""

    def __init__(self, seeds):
        self.seeds = seeds
        self.proc = None

    def __len__(self):
        return len(self.seeds)

    def __getitem__(self, i):
        from btcdet_b200 import synthetic as S
        if self.proc is None:
            self.proc = make_processor()
        pts = S.lidar_like(20000, seed=self.seeds[i])
        d = self.proc.forward({"points": pts, "use_lead_xyz": True})
        return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in d.items() if isinstance(v, np.ndarray)}


def _first(batch):
    return batch[0]


def _phi_ambiguous(sc, ulps=4):
    from btcdet_b200 import synthetic as S
    from oracle import coords
    phi = coords.absxyz_2_cylinxyz(sc)[:, 1]
    lo = np.floor((phi - ulps * np.spacing(phi) - np.float32(S.OCC_RANGE[1])) / np.float32(S.OCC_VOXEL_SIZE[1]))
    hi = np.floor((phi + ulps * np.spacing(phi) - np.float32(S.OCC_RANGE[1])) / np.float32(S.OCC_VOXEL_SIZE[1]))
    return int((lo != hi).sum())


@needs_ref
def test_reference_data_processor_on_the_shim_main_and_spawned_workers(cuda, oracle):
    from btcdet_b200 import synthetic as S
    from oracle import coords
    seeds = [60, 61, 62, 63]
    ds = _Scenes(seeds)
    main = [ds[i] for i in range(len(seeds))]
    gen_det = oracle.VoxelGeneratorV2(S.DET_VOXEL_SIZE, S.KITTI_RANGE, S.DET_MAX_POINTS, S.DET_MAX_VOXELS["test"])
    gen_occ = oracle.VoxelGeneratorV2(S.OCC_VOXEL_SIZE, S.OCC_RANGE, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["test"])
    for i, seed in enumerate(seeds):
        sc = S.lidar_like(20000, seed=seed)
        want = gen_det.generate(sc)
        np.testing.assert_array_equal(main[i]["det_voxels"].numpy(), want["voxels"])
        np.testing.assert_array_equal(main[i]["det_voxel_coords"].numpy(), want["coordinates"])
        np.testing.assert_array_equal(main[i]["det_voxel_num_points"].numpy(), want["num_points_per_voxel"])
        # the cylindrical branch: the processor's numpy a2 (coords_utils.py:282-292) feeds the shim's GPU voxeliser, so
        # it is bit-identical to numpy a2 + the CPU oracle (no CUDA libm involved on this route)
        want = gen_occ.generate(coords.absxyz_2_cylinxyz(sc))
        np.testing.assert_array_equal(main[i]["voxels"].numpy(), want["voxels"])
        np.testing.assert_array_equal(main[i]["voxel_coords"].numpy(), want["coordinates"])
        np.testing.assert_array_equal(main[i]["voxel_num_points"].numpy(), want["num_points_per_voxel"])
    # the same Dataset inside spawned DataLoader workers (CUDA cannot be used in forked workers; spawn is the supported mode)
    loader = torch.utils.data.DataLoader(_Scenes(seeds), batch_size=1, shuffle=False, num_workers=2,
                                         multiprocessing_context="spawn", collate_fn=_first)
    got = list(loader)
    assert len(got) == len(seeds)
    for a, b in zip(main, got):
        assert a.keys() == b.keys()
        for k in a:
            assert torch.equal(a[k], b[k]), k


@needs_ref
def test_reference_collate_and_load_to_gpu_equal_the_gpu_resident_pipeline(cuda, oracle):
    from btcdet_b200 import ops, synthetic as S
    _, dsmod = _reference_dataset_modules()
    seeds = [70, 71, 72]
    scenes = [S.lidar_like(20000 - 1500 * i, seed=s) for i, s in enumerate(seeds)]
    proc = make_processor()
    per_scene = []
    n_boxes = [3, 5, 2]
    for i, sc in enumerate(scenes):
        d = proc.forward({"points": sc.copy(), "use_lead_xyz": True})
        d["gt_boxes"] = np.full((n_boxes[i], 8), float(i + 1), np.float32)
        d["box_mirr_flag"] = np.ones(n_boxes[i], np.float32)
        d["is_train"] = False
        per_scene.append(d)
    batch = dsmod.DatasetTemplate.collate_batch(per_scene)                  # a3: dataset.py:168-223
    gpu = {k: torch.from_numpy(v).float().cuda() for k, v in batch.items()   # a4: models/__init__.py:16-22
           if isinstance(v, np.ndarray) and v.dtype != object and v.dtype != bool and k not in ("frame_id", "metadata", "calib")}
    pts, offs = S.batch_points(scenes)
    occ, det = ops.voxelize_occ_and_det(torch.from_numpy(pts).cuda(), torch.from_numpy(offs).cuda(), S.OCC_VOXEL_SIZE,
                                        S.OCC_RANGE, S.OCC_MAX_POINTS, S.OCC_MAX_VOXELS["test"], S.DET_VOXEL_SIZE,
                                        S.KITTI_RANGE, S.DET_MAX_POINTS, S.DET_MAX_VOXELS["test"], want_mean=False)
    m_det = int(det[4][-1].item())
    # detection branch: identical rows, batch index column included; the float -> int round trip of the reference is exact
    assert gpu["det_voxel_coords"].shape == (m_det, 4)
    assert torch.equal(gpu["det_voxel_coords"].int(), det[1][:m_det])
    assert torch.equal(gpu["det_voxels"], det[0][:m_det])
    assert torch.equal(gpu["det_voxel_num_points"].int(), det[2][:m_det])
    # points: batch index prepended as column 0 (dataset.py:187-192)
    assert torch.equal(gpu["points"][:, 1:], torch.from_numpy(pts).cuda())
    assert torch.equal(gpu["points"][:, 0].int(), torch.repeat_interleave(torch.arange(3), torch.from_numpy(np.diff(offs))).int().cuda())
    # occupancy branch: rho / z / intensity and the grouping identical unless a point's phi sits within 4 ulp of a bin edge
    # (numpy's arctan2 on the reference route, CUDA atan2f on the device route; tests/test_points_transform_gpu.py)
    if sum(_phi_ambiguous(sc) for sc in scenes) == 0:
        m_occ = int(occ[4][-1].item())
        assert torch.equal(gpu["voxel_coords"].int(), occ[1][:m_occ])
        assert torch.equal(gpu["voxel_num_points"].int(), occ[2][:m_occ])
        assert torch.equal(gpu["voxels"][..., [0, 2, 3]], occ[0][:m_occ][..., [0, 2, 3]])
        assert float((gpu["voxels"][..., 1] - occ[0][:m_occ][..., 1]).abs().max()) < 1e-4
    assert int(batch["batch_size"]) == 3
    # gt boxes: zero padded to the largest count, counts kept as a python list (dataset.py:193-200, 207-212)
    assert batch["gt_boxes"].shape == (3, 5, 8) and batch["gt_boxes_num"] == n_boxes and batch["box_mirr_flag"].shape == (3, 5)
    for i, k in enumerate(n_boxes):
        assert (batch["gt_boxes"][i, :k] == i + 1).all() and (batch["gt_boxes"][i, k:] == 0).all()
        assert batch["box_mirr_flag"][i, :k].sum() == k and batch["box_mirr_flag"][i, k:].sum() == 0
