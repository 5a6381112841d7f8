"""Seeded synthetic scenes: KITTI-range point clouds with no dataset on disk.

uniform(N, seed)     x~U[0,70.4) y~U[-40,40) z~U[-3,1), intensity~U[0,1)  — capacity / indexing tests.
lidar_like(N, seed)  64-beam spinning-LiDAR ray cast against a ground plane, ~25 boxes and two
                     walls, 1 cm range noise, cropped to the KITTI range and subsampled to N —
                     realistic submanifold neighbourhood statistics (~3 neighbours / voxel).
Geometry constants are the reference's: tools/cfgs/dataset_configs/kitti_dataset.yaml:4,
tools/cfgs/model_configs/btcdet_kitti_car.yaml:23-37.
"""
import numpy as np

KITTI_RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
DET_VOXEL_SIZE = [0.05, 0.05, 0.1]
DET_MAX_POINTS = 5
DET_MAX_VOXELS = {"train": 16000, "test": 40000}
OCC_VOXEL_SIZE = [0.32, 0.5184, 0.36]
OCC_RANGE = [2.24, -40.6944, -2.6, 69.12, 40.6944, 0.64]
OCC_MAX_POINTS = 12
OCC_MAX_VOXELS = {"train": 20000, "test": 40000}


def uniform(n, seed=0, point_range=KITTI_RANGE):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(point_range[:3]), np.array(point_range[3:])
    xyz = rng.uniform(lo, hi, size=(n, 3))
    inten = rng.uniform(0.0, 1.0, size=(n, 1))
    return np.concatenate([xyz, inten], axis=1).astype(np.float32)


def _ray_box(origin, dirs, center, dims, yaw):
    """Slab test of rays against one yawed box; returns t (inf where missed)."""
    c, s = np.cos(-yaw), np.sin(-yaw)
    rot = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    o = (origin - center) @ rot.T
    d = dirs @ rot.T
    half = dims / 2.0
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (-half - o) / d
        t2 = (half - o) / d
    tmin = np.nanmax(np.minimum(t1, t2), axis=1)
    tmax = np.nanmin(np.maximum(t1, t2), axis=1)
    hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.0)
    return np.where(hit, tmin, np.inf)


def make_boxes(rng, n_boxes=25):
    """gt_boxes rows [x, y, z, dx, dy, dz, yaw, class=1] standing on the ground plane z=-1.73."""
    x = rng.uniform(5.0, 65.0, n_boxes)
    y = rng.uniform(-30.0, 30.0, n_boxes)
    dx = rng.uniform(3.4, 4.6, n_boxes)
    dy = rng.uniform(1.5, 1.9, n_boxes)
    dz = rng.uniform(1.4, 1.8, n_boxes)
    z = -1.73 + dz / 2.0
    yaw = rng.uniform(-np.pi, np.pi, n_boxes)
    return np.stack([x, y, z, dx, dy, dz, yaw, np.ones(n_boxes)], axis=1).astype(np.float32)


def lidar_like(n, seed=0, n_boxes=25, az_density=1.0, point_range=KITTI_RANGE, return_boxes=False):
    rng = np.random.default_rng(seed)
    n_beams = 64
    n_az = int(max(700, np.ceil(2.6 * n / n_beams)) * az_density)
    elev = np.deg2rad(np.linspace(-24.9, 2.0, n_beams))
    az = np.deg2rad(np.linspace(-45.0, 45.0, n_az, endpoint=False))
    ee, aa = np.meshgrid(elev, az, indexing="ij")
    dirs = np.stack([np.cos(ee) * np.cos(aa), np.cos(ee) * np.sin(aa), np.sin(ee)], axis=-1).reshape(-1, 3)
    origin = np.zeros(3)
    t = np.full(dirs.shape[0], np.inf)
    # ground plane z = -1.73
    with np.errstate(divide="ignore"):
        tg = np.where(dirs[:, 2] < -1e-6, -1.73 / dirs[:, 2], np.inf)
    t = np.minimum(t, tg)
    # two walls y = +-wall_y, 3 m tall
    for wall_y in (rng.uniform(18.0, 34.0), -rng.uniform(18.0, 34.0)):
        with np.errstate(divide="ignore", invalid="ignore"):
            tw = wall_y / dirs[:, 1]
        zw = tw * dirs[:, 2]
        tw = np.where((tw > 0) & (zw > -1.73) & (zw < 1.3), tw, np.inf)
        t = np.minimum(t, tw)
    boxes = make_boxes(rng, n_boxes)
    for b in boxes:
        t = np.minimum(t, _ray_box(origin, dirs, b[:3].astype(np.float64), b[3:6].astype(np.float64), float(b[6])))
    keep = np.isfinite(t) & (t < 120.0)
    t = t[keep] + rng.normal(0.0, 0.01, keep.sum())
    xyz = dirs[keep] * t[:, None]
    lo, hi = np.array(point_range[:3]), np.array(point_range[3:])
    inside = np.all((xyz >= lo) & (xyz < hi), axis=1)
    xyz = xyz[inside]
    if xyz.shape[0] >= n:
        sel = rng.permutation(xyz.shape[0])[:n]
    else:  # not enough returns: top up with jittered copies
        extra = rng.integers(0, xyz.shape[0], n - xyz.shape[0])
        xyz = np.concatenate([xyz, xyz[extra] + rng.normal(0.0, 0.02, (extra.size, 3))], axis=0)
        xyz = np.clip(xyz, lo, np.nextafter(hi, lo))
        sel = rng.permutation(xyz.shape[0])
    xyz = xyz[sel]
    inten = rng.uniform(0.0, 1.0, size=(xyz.shape[0], 1))
    pts = np.concatenate([xyz, inten], axis=1).astype(np.float32)
    return (pts, boxes) if return_boxes else pts


def batch_points(scenes):
    """Concatenate scenes; returns (points [sum N, C] f32, scene_offsets [B+1] i32)."""
    offs = np.zeros(len(scenes) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([s.shape[0] for s in scenes])
    return np.ascontiguousarray(np.concatenate(scenes, axis=0), dtype=np.float32), offs


def roi_head_case(batch=2, n_points=20000, n_rois=128, n_occ=3000, channels=128, seed=0, stride=8):
    """Synthetic inputs of the RoI head (`ConvHead.roi_conv_pool`, conv_head.py:247-379), numpy:
    rois [B, n_rois, 7] (the scene's boxes, jittered and repeated), points [sum N, 5] (b, x, y, z, intensity),
    occ_pnts [sum n_occ, 4] (x, y, z, probability) + their scene index, and the rows of a `x_combine`-like sparse
    tensor at `stride` (coords [n, 4] b z y x over [2, 200, 176], features [n, channels] >= 0 with exact zeros): the
    occupied cells are those that contain points."""
    rng = np.random.default_rng(1000 + seed)
    pts, rois, occ, occ_b, coords = [], [], [], [], []
    vs = np.array(DET_VOXEL_SIZE) * stride
    shape = [2, 200, 176]
    for b in range(batch):
        p, boxes = lidar_like(n_points, seed=seed * 16 + b, return_boxes=True)
        pts.append(np.concatenate([np.full((p.shape[0], 1), b, np.float32), p], axis=1))
        pick = rng.integers(0, boxes.shape[0], n_rois)
        r = boxes[pick, :7].copy()
        r[:, :3] += rng.normal(0.0, 0.3, (n_rois, 3)).astype(np.float32)
        r[:, 3:6] *= rng.uniform(0.9, 1.1, (n_rois, 3)).astype(np.float32)
        r[:, 6] += rng.normal(0.0, 0.1, n_rois).astype(np.float32)
        rois.append(r)
        o = boxes[rng.integers(0, boxes.shape[0], n_occ), :3] + rng.normal(0.0, 0.8, (n_occ, 3))
        occ.append(np.concatenate([o, rng.uniform(0.3, 1.0, (n_occ, 1))], axis=1).astype(np.float32))
        occ_b.append(np.full(n_occ, b, np.int64))
        x = np.floor((p[:, 0] - KITTI_RANGE[0]) / vs[0]).astype(np.int64)
        y = np.floor((p[:, 1] - KITTI_RANGE[1]) / vs[1]).astype(np.int64)
        z = np.clip(np.floor((p[:, 2] - KITTI_RANGE[2]) / (vs[2] * 2.5)).astype(np.int64), 0, shape[0] - 1)
        ok = (x >= 0) & (x < shape[2]) & (y >= 0) & (y < shape[1])
        cells = np.unique(np.stack([z[ok], y[ok], x[ok]], axis=1), axis=0)
        coords.append(np.concatenate([np.full((cells.shape[0], 1), b, np.int64), cells], axis=1))
    coords = np.concatenate(coords, axis=0).astype(np.int32)
    feats = np.maximum(rng.standard_normal((coords.shape[0], channels)), 0.0).astype(np.float32)
    return {"batch_size": batch, "rois": np.stack(rois).astype(np.float32), "points": np.concatenate(pts).astype(np.float32),
            "occ_pnts": np.concatenate(occ), "added_occ_b_ind": np.concatenate(occ_b), "x_coords": coords,
            "x_features": feats, "x_shape": shape}
