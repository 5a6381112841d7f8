"""ORACLE (test infrastructure only — never imported by btcdet_b200/ or spconv/): rotated BEV overlap / IoU and greedy NMS.

Restates what the reference's iou3d_nms extension computes (btcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:128-230 box_overlap /
iou_bev; iou3d_nms.cpp:103-140 + iou3d_nms_kernel.cu:269-340 greedy suppression of score-sorted boxes) in float64 numpy:
intersection polygon of two rotated rectangles by half-plane clipping, shoelace area, IoU = overlap / max(sa + sb -
overlap, 1e-8).  Pinned against the reference's own CPU code compiled from the checkout (`oracle/_ref/libiou3d_ref.so`,
built by `make -C oracle ref`) in tests/test_iou3d_cpu.py; the GPU tests use that library directly when it is present.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libiou3d_ref.so")


def _corners(box):
    x, y, dx, dy, a = box[0], box[1], box[3], box[4], box[6]
    c, s = np.cos(a), np.sin(a)
    pts = np.array([[-dx / 2, -dy / 2], [dx / 2, -dy / 2], [dx / 2, dy / 2], [-dx / 2, dy / 2]], np.float64)
    return pts @ np.array([[c, s], [-s, c]]) + np.array([x, y])


def _clip(poly, p0, p1):
    """Keep the part of `poly` on the left of the directed line p0 -> p1."""
    out = []
    d = p1 - p0
    side = lambda p: d[0] * (p[1] - p0[1]) - d[1] * (p[0] - p0[0])   # noqa: E731
    for i in range(len(poly)):
        a, b = poly[i], poly[(i + 1) % len(poly)]
        sa, sb = side(a), side(b)
        if sa >= 0:
            out.append(a)
        if (sa >= 0) != (sb >= 0):
            t = sa / (sa - sb)
            out.append(a + t * (b - a))
    return out


def box_overlap(box_a, box_b):
    poly = list(_corners(np.asarray(box_a, np.float64)))
    cb = _corners(np.asarray(box_b, np.float64))
    for i in range(4):
        if len(poly) < 3:
            return 0.0
        poly = _clip(poly, cb[i], cb[(i + 1) % 4])
    if len(poly) < 3:
        return 0.0
    p = np.array(poly)
    return 0.5 * abs(float(np.sum(p[:, 0] * np.roll(p[:, 1], -1) - np.roll(p[:, 0], -1) * p[:, 1])))


def iou_bev(box_a, box_b):
    ov = box_overlap(box_a, box_b)
    return ov / max(float(box_a[3]) * float(box_a[4]) + float(box_b[3]) * float(box_b[4]) - ov, 1e-8)


def boxes_bev(a, b, overlap=False):
    out = np.zeros((len(a), len(b)), np.float64)
    for i in range(len(a)):
        for j in range(len(b)):
            out[i, j] = box_overlap(a[i], b[j]) if overlap else iou_bev(a[i], b[j])
    return out


def iou_normal(a, b):
    l, r = max(a[0] - a[3] / 2, b[0] - b[3] / 2), min(a[0] + a[3] / 2, b[0] + b[3] / 2)
    t, d = max(a[1] - a[4] / 2, b[1] - b[4] / 2), min(a[1] + a[4] / 2, b[1] + b[4] / 2)
    ov = max(r - l, 0.0) * max(d - t, 0.0)
    return ov / max(a[3] * a[4] + b[3] * b[4] - ov, 1e-8)


def greedy_nms(iou, thresh):
    """Kept indices of score-sorted boxes given their pairwise IoU matrix (what the reference's mask + host scan does)."""
    n = iou.shape[0]
    removed = np.zeros(n, bool)
    keep = []
    for i in range(n):
        if removed[i]:
            continue
        keep.append(i)
        removed[i + 1:] |= iou[i, i + 1:] > thresh
    return np.array(keep, np.int64)


def reference_lib():
    """ctypes handle of the reference's own CPU code (oracle/_ref/libiou3d_ref.so) or None when it was not built."""
    if not os.path.exists(REF_LIB):
        return None
    import torch  # noqa: F401  (the library links against libtorch for the symbols of the reference's tensor entry point)
    lib = ctypes.CDLL(REF_LIB)
    lib.ref_boxes_bev.restype = None
    lib.ref_boxes_bev.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def reference_boxes_bev(a, b, overlap=False):
    """[n,7], [m,7] float32 -> [n,m] float32 through the reference's iou_bev / box_overlap."""
    lib = reference_lib()
    assert lib is not None
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.zeros((len(a), len(b)), np.float32)
    lib.ref_boxes_bev(a.ctypes.data, len(a), b.ctypes.data, len(b), int(overlap), out.ctypes.data)
    return out
