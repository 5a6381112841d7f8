"""SURVEY 8(f) N2 oracle: the float64 restatement (oracle/iou3d.py) against hand-derived answers and against the
REFERENCE'S OWN CPU code compiled from the checkout (oracle/_ref/libiou3d_ref.so, `make -C oracle ref`)."""
import numpy as np
import pytest

from oracle import iou3d


def box(x, y, dx, dy, a):
    return np.array([x, y, 0.0, dx, dy, 1.5, a], np.float32)


def random_boxes(n, seed):
    rng = np.random.default_rng(seed)
    b = np.zeros((n, 7), np.float32)
    b[:, 0], b[:, 1] = rng.uniform(0, 20, n), rng.uniform(-10, 10, n)
    b[:, 3], b[:, 4], b[:, 5] = rng.uniform(1, 5, n), rng.uniform(1, 3, n), 1.5
    b[:, 6] = rng.uniform(-3.2, 3.2, n)
    return b


def test_known_answers():
    a = box(0, 0, 4, 2, 0.0)
    assert iou3d.iou_bev(a, a) == pytest.approx(1.0, abs=1e-12)
    assert iou3d.box_overlap(a, box(10, 0, 4, 2, 0.3)) == 0.0                       # disjoint
    assert iou3d.box_overlap(a, box(2, 0, 4, 2, 0.0)) == pytest.approx(4.0, abs=1e-9)   # half of each
    assert iou3d.iou_bev(a, box(2, 0, 4, 2, 0.0)) == pytest.approx(4.0 / 12.0, abs=1e-9)
    assert iou3d.box_overlap(box(0, 0, 2, 2, 0.0), box(0, 0, 2, 2, np.pi / 2)) == pytest.approx(4.0, abs=1e-6)
    # a unit square rotated by 45 degrees inside a big box: its own area; heading sign convention = counter-clockwise
    assert iou3d.box_overlap(box(0, 0, 10, 10, 0.0), box(1, 1, 1, 1, np.pi / 4)) == pytest.approx(1.0, abs=1e-6)
    long_a, long_b = box(0, 0, 6, 1, np.pi / 6), box(0, 0, 6, 1, -np.pi / 6)
    assert iou3d.box_overlap(long_a, long_b) == pytest.approx(1.0 / np.sin(np.pi / 3), abs=1e-6)   # rhombus of two unit strips
    assert iou3d.iou_normal(box(0, 0, 4, 2, 1.0), box(2, 0, 4, 2, -2.0)) == pytest.approx(4.0 / 12.0, abs=1e-9)


def test_greedy_nms_known_answer():
    iou = np.array([[1, .8, .1, .0], [.8, 1, .6, .0], [.1, .6, 1, .9], [.0, .0, .9, 1]])
    assert iou3d.greedy_nms(iou, 0.5).tolist() == [0, 2]
    assert iou3d.greedy_nms(iou, 0.95).tolist() == [0, 1, 2, 3]


@pytest.mark.skipif(iou3d.reference_lib() is None, reason="oracle/_ref/libiou3d_ref.so not built (needs the reference checkout)")
def test_oracle_against_the_reference_cpu_code():
    """The reference's fp32 routine adds a box corner to the intersection polygon when it lies up to 1e-2 m OUTSIDE the
    other box (its in-box test carries a margin), so it over-estimates the exact area by up to margin x edge length in
    those configurations; everywhere else the two agree to fp32 rounding."""
    a, b = random_boxes(200, 1), random_boxes(180, 2)
    ref_ov, ora_ov = iou3d.reference_boxes_bev(a, b, True), iou3d.boxes_bev(a, b, True)
    hit = ora_ov > 0
    assert hit.sum() > 1000
    d = np.abs(ref_ov - ora_ov)
    assert d.max() < 0.03                                         # 1e-2 m margin x (at most ~5 m of edge) / 2, with slack
    assert np.quantile(d[hit], 0.9) < 2e-4 and np.median(d[hit]) < 2e-5
    assert (ref_ov[~hit] < 1e-3).all()
    ref_iou, ora_iou = iou3d.reference_boxes_bev(a, b), iou3d.boxes_bev(a, b)
    assert np.abs(ref_iou - ora_iou).max() < 3e-3
    # identical boxes and symmetric pairs
    assert np.allclose(np.diag(iou3d.reference_boxes_bev(a[:20], a[:20])), 1.0, atol=1e-5)
