"""GPU tests of the dense solver of the reduced camera system (mageslam_b200/csrc/dense_ldlt.cuh): blocked LDL^T whose trailing
update runs on the tcgen05 tensor cores as exact int8 slice products (Ozaki scheme) + the two grid-wide substitutions, through the
test entry mage_dense_debug_solve. Reference: numpy's FP64 solve / the reconstruction L D L^T = A (the reference itself calls
Eigen::LDLT, ref Dependencies/g2o/g2o/solvers/dense/linear_solver_dense.h:65-113). Tolerances are FP64-level: the slices carry 56
bits below each row's largest entry."""
import ctypes as C

import numpy as np
import pytest

from mageslam_b200 import _lib

pytestmark = pytest.mark.gpu


def solve(A, b, want_factor=True):
    L = _lib.lib()
    n = len(b)
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n); F = np.zeros((n, n)) if want_factor else None
    ok = C.c_int(0); ns = (C.c_longlong * 16)()
    rc = L.mage_dense_debug_solve(n, A.ctypes.data, b.ctypes.data, x.ctypes.data, F.ctypes.data if want_factor else None, C.byref(ok), ns)
    assert rc == 0, L.mage_last_error()
    return x, F, ok.value, list(ns)


def spd(n, seed, spread=0.0):
    """random SPD matrix with a condition number of a few thousand; spread > 0 scales rows / columns by 10^U(-spread, spread)
    (entries of very different magnitude, as the 6 x 6 pose blocks of a reduced camera system have)"""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, max(n // 2, 8)))
    A = M @ M.T / M.shape[1] + 0.05 * np.eye(n)
    if spread > 0:
        d = 10.0 ** rng.uniform(-spread, spread, n)
        A = A * d[:, None] * d[None, :]
    return (A + A.T) / 2


@pytest.mark.parametrize("n", [6, 90, 128, 130, 200, 256, 300, 514, 1000, 1666, 2604, 2976])
def test_solution_and_factor_match_fp64(n):
    A = spd(n, n, spread=2.0)
    rng = np.random.default_rng(n + 1)
    b = rng.standard_normal(n)
    x, F, ok, _ = solve(A, b)
    assert ok == 1
    Lm = np.tril(F, -1) + np.eye(n); D = np.diag(F).copy()
    assert (D > 0).all()
    rec = (Lm * D[None, :]) @ Lm.T
    scale = np.sqrt(np.outer(np.diag(A), np.diag(A)))                  # element-wise natural scale of an SPD matrix
    assert np.max(np.abs(rec - A) / scale) < 1e-12, "L D L^T differs from A by %.3g (scaled)" % np.max(np.abs(rec - A) / scale)
    xr = np.linalg.solve(A, b)
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    res_ref = np.linalg.norm(A @ xr - b) / np.linalg.norm(b)
    assert res < 50 * max(res_ref, 1e-15), "residual %.3g vs numpy's %.3g" % (res, res_ref)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9


def test_tier_size_reduced_system():
    """n = 2988 (500 key frames, 2 fixed): the size of BASELINE config 4"""
    n = 2988
    A = spd(n, 7, spread=1.5)
    b = np.random.default_rng(8).standard_normal(n)
    x, _, ok, ns = solve(A, b, want_factor=False)
    assert ok == 1
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
    print("dense solve n=2988 phases (us): diag %.0f panel %.0f update %.0f barriers %.0f back-substitution %.0f; inside diag: sub-factor %.0f rows %.0f update %.0f; update epilogue warp: wait %.0f (-) %.0f store %.0f; whole kernel %.0f" % tuple(v / 1e3 for v in list(ns[:11]) + [ns[15]]))


def test_not_positive_definite_is_reported():
    """a non-positive pivot => 'not positive' (ref linear_solver_dense.h:104-112), in the first panel and in a later one"""
    for n, bad in ((200, 17), (300, 250)):
        A = spd(n, 3)
        A[bad, bad] = -1.0
        _, _, ok, _ = solve(A, np.ones(n))
        assert ok == 0


def test_rows_of_zeros_and_tiny_entries():
    """decoupled unknowns (zero off-diagonal rows) and entries far below the row maximum survive the slicing"""
    n = 400
    A = spd(n, 11)
    A[150, :] = 0; A[:, 150] = 0; A[150, 150] = 3.0
    A[:, 300] *= 1e-9; A[300, :] *= 1e-9
    b = np.random.default_rng(12).standard_normal(n)
    x, _, ok, _ = solve(A, b)
    assert ok == 1
    xr = np.linalg.solve(A, b)
    assert np.linalg.norm(x - xr) / np.linalg.norm(xr) < 1e-9
    assert abs(x[150] - b[150] / 3.0) < 1e-14
