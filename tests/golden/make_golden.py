"""Generates tests/golden/config1.npz — BASELINE.json configs[0] ("CPU-only synthetic: 1 scene of 2k
random pts -> points_to_voxel + one SubMConv3d(16->16,k3) ... dump indices+features").

The reference's spconv 1.2.1 is not installable here, so the vectors come from the CPU
oracle (oracle/), which tests/test_oracle_cpu.py pins against dense stock-torch convolutions.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from btcdet_b200 import synthetic as S  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    pts = S.uniform(2000, seed=0)
    gen = O.VoxelGeneratorV2(S.DET_VOXEL_SIZE, S.KITTI_RANGE, max_num_points=5, max_voxels=16000)
    r = gen.generate(pts)
    coords = np.pad(r["coordinates"], ((0, 0), (1, 0))).astype(np.int32)
    m = coords.shape[0]
    feat = np.random.default_rng(1).standard_normal((m, 16)).astype(np.float32)
    weight = (np.random.default_rng(2).standard_normal((3, 3, 3, 16, 16)) * 0.1).astype(np.float32)
    outids, pairs, pair_num, _ = O.get_indice_pairs(coords, 1, [41, 1600, 1408], 3, subm=True)
    out = O.indice_conv(feat, weight.reshape(27, 16, 16), pairs, pair_num, m, subm=True, use_c=True)
    out_mm = O.indice_conv(feat, weight.reshape(27, 16, 16), pairs, pair_num, m, subm=True)
    assert np.abs(out - out_mm).max() < 1e-5
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config1.npz")
    np.savez_compressed(path, coordinates=r["coordinates"], num_points_per_voxel=r["num_points_per_voxel"],
                        features=feat, weight=weight, indice_pairs=pairs, indice_pair_num=pair_num, out_features=out)
    print(path, m, int(pair_num.sum()), os.path.getsize(path))


if __name__ == "__main__":
    main()
