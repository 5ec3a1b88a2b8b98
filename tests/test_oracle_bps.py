"""CPU tier: the BPS oracle (oracle/bps_oracle.c) against the reference's own checks re-instantiated with seeds:
factors and tau equal LAPACK's unblocked QR of Matrix(A) (test_qr.jl:13-30), lmul!(Q',b) equals the explicit
reflector product (:37-48), ldiv!(R,b) equals triu(F)\\b (:50-53), A*b and UpperTriangular(A)\\b
(test_bandedplussemiseparable.jl:11-17)."""
import numpy as np
import pytest

import oracle
from helpers import bps_randn, explicit_qt, relerr

SHAPES = [(20, 4, 5, 2, 3),      # test_qr.jl:6-7, test_bandedplussemiseparable.jl:4-5
          (64, 3, 3, 2, 2), (30, 3, 0, 1, 1), (30, 1, 3, 2, 2), (25, 2, 2, 3, 1), (30, 0, 3, 2, 2), (30, 0, 0, 1, 1),
          (6, 4, 5, 2, 3), (3, 3, 3, 2, 2), (2, 1, 1, 1, 1), (1, 1, 1, 1, 1)]


@pytest.mark.parametrize("n,l,m,r,p", SHAPES)
def test_qr_factors_match_dense(n, l, m, r, p):
    rng = np.random.default_rng(1000 + n + l)
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p)
    A = oracle.bps_dense(B, U, V, W, S, l, m)
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, m)
    F = oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, l + m)
    qr_, t = oracle.dense_qr_unblocked(A)
    cond = np.linalg.cond(A)
    assert np.abs(F - qr_).max() <= 1e-14 * cond * np.abs(qr_).max()
    assert np.abs(tau - t).max() <= 1e-14 * cond
    c = oracle.bps_lmul_qt(Bf, U, Vf, tau, b, l, l + m)
    assert np.abs(c - explicit_qt(qr_, t, b)).max() <= 1e-13 * cond
    x = oracle.bps_upper_ldiv(Bf, Wf, Sf, c, l, l + m)
    assert relerr(x, np.linalg.solve(A, b)) <= 1e-14 * cond * 10


def test_mul_and_upper_ldiv():
    """test_bandedplussemiseparable.jl:11,16,17."""
    import scipy.linalg as sl
    rng = np.random.default_rng(7)
    n, l, m, r, p = 20, 4, 5, 2, 3
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p)
    A = oracle.bps_dense(B, U, V, W, S, l, m)
    Bd = np.zeros((n, n))
    for j in range(n):
        for k in range(max(0, j - m), min(n, j + l + 1)):
            Bd[k, j] = B[j, m + k - j]
    assert np.allclose(A, Bd + np.tril(U.T @ V, -1) + np.triu(W.T @ S, 1))
    assert np.allclose(oracle.bps_mul(B, U, V, W, S, b, l, m), A @ b, atol=1e-13)
    x = oracle.bps_upper_ldiv(B, W, S, b, l, m)
    assert relerr(x, sl.solve_triangular(np.triu(A), b)) <= 1e-8


def test_generator_conditioning():
    B, U, V, W, S, b = (a[0] for a in oracle.bps_generate(256, 3, 3, 2, 2, 1, seed=20261021))
    A = oracle.bps_dense(B, U, V, W, S, 3, 3)
    assert np.linalg.cond(A) < 5
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, 3, 3)
    qr_, t = oracle.dense_qr_unblocked(A)
    assert np.abs(oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, 3, 6) - qr_).max() <= 1e-14 * np.abs(qr_).max()
