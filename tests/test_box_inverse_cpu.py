"""The LU inverse and GEMM chains that csrc/box_masks.cu follows, restated in numpy and pinned to what torch.inverse /
torch.einsum produced ON A B200 (tests/golden/box_inverse_cuda.npz, written by tools/o3_reference_cuda.py from the
reference's own transform construction, point_box_utils.py:272-288,310-329).  This is the evidence that rows a9-a11 can be
bit-exact: 256 / 256 inverses and 18 432 / 18 432 box-frame coordinates identical."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32


def fma(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))     # exact product, one rounding (double rounding: negligible)


def lu_inverse(A):
    """partial pivoting (first max), l = a * (1 / pivot), FMA updates, column-oriented substitution, division by U_kk."""
    n = A.shape[0]
    A = A.astype(np.float32).copy()
    piv = list(range(n))
    for k in range(n):
        p = k + int(np.argmax(np.abs(A[k:, k])))
        if p != k:
            A[[k, p]] = A[[p, k]]
            piv[k], piv[p] = piv[p], piv[k]
        r = f32(1) / A[k, k]
        for i in range(k + 1, n):
            A[i, k] = f32(A[i, k] * r)
        for i in range(k + 1, n):
            for j in range(k + 1, n):
                A[i, j] = fma(-A[i, k], A[k, j], A[i, j])
    X = np.eye(n, dtype=np.float32)[piv]
    for k in range(n):
        for i in range(k + 1, n):
            for j in range(n):
                X[i, j] = fma(-A[i, k], X[k, j], X[i, j])
    for k in range(n - 1, -1, -1):
        for j in range(n):
            X[k, j] = f32(X[k, j] / A[k, k])
        for i in range(k):
            for j in range(n):
                X[i, j] = fma(-A[i, k], X[k, j], X[i, j])
    return X


def test_lu_inverse_reproduces_torch_inverse_on_cuda():
    d = np.load(os.path.join(HERE, "golden", "box_inverse_cuda.npz"))
    T, want = d["T"], d["inv_all"]
    assert T.shape == (256, 4, 4)
    for m in range(T.shape[0]):
        np.testing.assert_array_equal(lu_inverse(T[m]), want[m], err_msg=str(m))
    if "T2" in d.files:                                  # 3x3 transforms of the 2-D pre-filter (later dumps)
        for m in range(d["T2"].shape[0]):
            np.testing.assert_array_equal(lu_inverse(d["T2"][m]), d["inv2_all"][m], err_msg="2d %d" % m)


def test_box_frame_coordinates_follow_the_gemm_fma_chain():
    d = np.load(os.path.join(HERE, "golden", "box_inverse_cuda.npz"))
    pts, inv, q = d["pts"], d["inv_12"], d["q"]
    v = np.vectorize(fma, otypes=[np.float32])
    for m in range(inv.shape[0]):
        for i in range(3):
            a = inv[m, i]
            acc = (pts[:, 0] * a[0]).astype(np.float32)
            acc = v(pts[:, 1], a[1], acc)
            acc = v(pts[:, 2], a[2], acc)
            np.testing.assert_array_equal((acc + a[3]).astype(np.float32), q[:, m, i])
