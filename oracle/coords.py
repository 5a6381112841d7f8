"""numpy restatement of the reference's dataset-side coordinate transforms — TEST INFRASTRUCTURE ONLY.

Follows btcdet/utils/coords_utils.py:268-292 line by line (np.linalg.norm over the first 2 / 3 columns, arctan2 * 180. / pi
in float32 = two roundings).  Pinned bit-for-bit against the reference's own functions in tests/test_points_transform_cpu.py
(they are plain numpy and import without CUDA).
"""
import numpy as np


def absxyz_2_cylinxyz(points):
    """coords_utils.py:282-292."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    rho = np.linalg.norm(points[:, :2], axis=1)
    phi = np.arctan2(-y, x) * 180. / np.pi
    xyz = np.stack([rho, phi, z], axis=-1)
    return np.concatenate([xyz, points[:, 3:]], axis=-1) if points.shape[1] > 3 else xyz


def absxyz_2_spherexyz(points):
    """coords_utils.py:268-279."""
    x, y, z = points[:, 0], points[:, 1], points[:, 2]
    dist = np.linalg.norm(points[:, :3], axis=1)
    xydist = np.linalg.norm(points[:, :2], axis=1)
    az = np.arctan2(-y, x) * 180. / np.pi
    el = np.arctan2(z, xydist) * 180. / np.pi
    xyz = np.stack([dist, az, el], axis=-1)
    return np.concatenate([xyz, points[:, 3:]], axis=-1) if points.shape[1] > 3 else xyz


def ulp_distance(a, b):
    """Distance in float32 units in the last place between two float32 arrays (same sign regions)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)
