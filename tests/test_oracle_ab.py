"""CPU tier: the AlmostBanded oracle (oracle/ab_oracle.c) against the reference's own checks —
the README known-answer vector (bitwise) and dense LAPACK on re-instantiations of test_almostbanded.jl."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

import oracle
from helpers import (R_from_factors, degenerate_system, kat_n80, randn_system, readme_expz, relerr, vcat_system)

GOLDEN = Path(__file__).parent / "golden"


def test_readme_known_answer_bitwise():
    """README.md:73-94: the oracle reproduces all 20 printed values bit for bit (and 1/k! to 2.3e-16)."""
    g = json.loads((GOLDEN / "readme_expz.json").read_text())
    (l, u, r), bands, U, V, b = readme_expz(g["n"])
    x = oracle.ab_solve(bands, U, V, b, l, u)
    assert np.array_equal(x, np.array(g["x"]))
    assert np.abs(x - np.array([1 / math.factorial(k) for k in range(g["n"])])).max() < 3e-16


def test_readme_large_n():
    """README.md:96-97: the same system at n = 1_000_000 (config 1) solves in O(n) and stays accurate."""
    (l, u, r), bands, U, V, b = readme_expz(1_000_000)
    x = oracle.ab_solve(bands, U, V, b, l, u)
    assert np.abs(x[:20] - np.array([1 / math.factorial(k) for k in range(20)])).max() < 1e-15
    assert oracle.ab_relres(bands, U, V, x, b, l, u) < 1e-15


@pytest.mark.parametrize("n,l,u,r", [(20, 1, 0, 1), (80, 1, 1, 1), (64, 2, 2, 2), (50, 4, 4, 2), (33, 2, 0, 2), (7, 1, 0, 2),
                                      (40, 3, 1, 3), (5, 3, 3, 2), (2, 1, 1, 1), (1, 1, 1, 1)])
def test_factors_match_dense_geqrf(n, l, u, r):
    """F.Q'A == F.R (test_almostbanded.jl:145), sharpened: factors and tau equal LAPACK's unblocked QR."""
    rng = np.random.default_rng(100 * n + l)
    bands, U, V, b = randn_system(rng, n, l, u, r)
    A = oracle.ab_dense(bands, U, V, l, u)
    F, Uf, tau = oracle.ab_qr(bands, U, V, l, u)
    qr_, t = oracle.dense_qr_unblocked(A)
    R = R_from_factors(F, Uf, V, l, l + u)
    assert np.abs(R - np.triu(qr_)).max() <= 1e-13 * np.abs(qr_).max()
    assert np.abs(np.tril(oracle.storage_to_dense(F, l, l + u), -1) - np.tril(qr_, -1)).max() <= 1e-13
    assert np.abs(tau - t).max() <= 1e-13
    x = oracle.ab_solve(bands, U, V, b, l, u)
    assert relerr(x, np.linalg.solve(A, b)) <= 1e-10 * max(1.0, np.linalg.cond(A) * 1e-6)


def test_kat_n80_qr_and_solve():
    """test_almostbanded.jl:135-158."""
    (l, u, r), bands, U, V = kat_n80()
    n = 80
    rng = np.random.default_rng(0)
    b = rng.standard_normal(n)
    A = oracle.ab_dense(bands, U, V, l, u)
    F, Uf, tau = oracle.ab_qr(bands, U, V, l, u)
    # widening keeps values (:140-141); input untouched (:146) is by construction (const pointers)
    qr_, t = oracle.dense_qr_unblocked(A)
    assert np.abs(R_from_factors(F, Uf, V, l, l + u) - np.triu(qr_)).max() < 1e-12
    x = oracle.ab_solve(bands, U, V, b, l, u)
    assert relerr(x, np.linalg.solve(A, b)) < 1e-13                                   # :153
    x2 = oracle.ab_ldiv(F, Uf, V, tau, b, l, l + u)                                   # :154 (bitwise)
    assert np.array_equal(x, x2)
    Qtb = oracle.ab_lmul_qt(F, tau, b, l, l + u)
    assert np.allclose(oracle.ab_lmul_qt(F, tau, Qtb, l, l + u, adjoint=False), b, atol=1e-13)   # :157-158


def test_triangular_solves():
    """test_almostbanded.jl:123-132."""
    n = 80
    A = np.full((n, n), 2.0)
    bands = oracle.banded_to_storage(A, 1, 1)
    U = np.full((1, n), 3.0); V = np.ones((n, 1))
    Ad = oracle.ab_dense(bands, U, V, 1, 1)
    b = np.random.default_rng(1).standard_normal(n)
    import scipy.linalg as sl
    for upper in (True, False):
        for unit in (False, True):
            T = np.triu(Ad) if upper else np.tril(Ad)
            ref = sl.solve_triangular(T, b, lower=not upper, unit_diagonal=unit)
            got = oracle.ab_tri_ldiv(bands, U, V, b, 1, 1, upper=upper, unit=unit)
            assert relerr(got, ref) < 1e-9     # the reference's own tolerance is sqrt(eps); the system is ill-conditioned


def test_vcat_and_degenerate():
    """test_almostbanded.jl:160-163,179-185."""
    rng = np.random.default_rng(2)
    (l, u, r), bands, U, V, A = vcat_system(rng)
    assert np.array_equal(oracle.ab_dense(bands, U, V, l, u), A)
    b = rng.standard_normal(80)
    assert relerr(oracle.ab_solve(bands, U, V, b, l, u), np.linalg.solve(A, b)) < 1e-9
    (l, u, r), bands, U, V, A = degenerate_system(rng)
    assert np.array_equal(oracle.ab_dense(bands, U, V, l, u), A)
    b = rng.standard_normal(7)
    assert relerr(oracle.ab_solve(bands, U, V, b, l, u), np.linalg.solve(A, b)) < 1e-11


def test_one_column_qr():
    """test_almostbanded.jl:169-177: _almostbanded_qr!(A,[1.0],1) on a 2x2 equals dense qr(A[:,1]).Q' * A."""
    rng = np.random.default_rng(3)
    bands, U, V, _ = randn_system(rng, 2, 1, 1, 1)
    A = oracle.ab_dense(bands, U, V, 1, 1)
    F, Uf, tau = oracle.ab_qr(bands, U, V, 1, 1, ncols=1)
    x = A[:, 0].copy()
    nu = math.copysign(np.linalg.norm(x), x[0]); v = np.array([1.0, x[1] / (x[0] + nu)]); t = (x[0] + nu) / nu
    QtA = A - t * np.outer(v, v @ A)
    assert np.allclose(np.triu(QtA), np.triu(oracle.storage_to_dense(F, 1, 2)), atol=1e-14)


def _dense_partial_householder(A, ncols):
    """H_ncols ... H_1 A with LinearAlgebra.reflector! reflectors, and their tau."""
    M = A.copy(); n = M.shape[0]; tau = np.zeros(n)
    for k in range(ncols):
        x = M[k:, k].copy()
        nrm = np.linalg.norm(x)
        if nrm == 0.0:
            continue
        nu = math.copysign(nrm, x[0]); v = x / (x[0] + nu); v[0] = 1.0; t = (x[0] + nu) / nu
        M[k:, :] -= t * np.outer(v, v @ M[k:, :])
        tau[k] = t
    return M, tau


def partial_factors_dense(F, Uf, V, l, u, ncols):
    """The matrix a partially factored AlmostBandedMatrix stands for: R rows / trailing block from the widened band storage, fill Uf*V
    beyond it, zeros below the diagonal of the ncols factored columns (the storage holds the reflector tails there)."""
    n = F.shape[0]; ub = l + u
    M = oracle.storage_to_dense(F, l, ub)
    fill = Uf.T @ V.T if Uf.size else np.zeros((n, n))      # U is stored (r, n), V (n, r): column-major n x r and r x n
    for k in range(n):
        M[k, k + ub + 1:] = fill[k, k + ub + 1:]
    for j in range(min(ncols, n)):
        M[j + 1:, j] = 0.0
    return M


def test_partial_qr_matches_dense_householder():
    """_almostbanded_qr!(A, tau, ncols) (:139-172) for every kind of ncols: inside a block, at a block edge, n-1 (the default), n."""
    rng = np.random.default_rng(17)
    for (n, l, u, r) in [(24, 2, 2, 2), (17, 1, 0, 1), (30, 4, 4, 2), (21, 3, 1, 3), (12, 0, 2, 1)]:
        bands, U, V, _ = randn_system(rng, n, l, u, r)
        A = oracle.ab_dense(bands, U, V, l, u)
        for ncols in sorted({1, 2, l + u + 1, l + u + 2, n // 2, n - 1, n}):
            F, Uf, tau = oracle.ab_qr(bands, U, V, l, u, ncols=ncols)
            M, t = _dense_partial_householder(A, ncols)
            if l == 0:            # reference quirk: length-1 reflectors give tau = 2 and a negated row
                assert np.allclose(tau[:ncols], 2.0) and np.all(tau[ncols:] == 0.0)
                continue
            assert np.abs(partial_factors_dense(F, Uf, V, l, u, ncols) - M).max() <= 1e-12 * np.abs(A).max(), (n, l, u, r, ncols)
            assert np.abs(tau - t).max() <= 1e-13 and np.all(tau[ncols:] == 0.0)


def test_generator_is_well_conditioned_and_deterministic():
    bands, U, V, b = oracle.ab_generate(64, 2, 2, 2, 3, seed=20261020)
    b2 = oracle.ab_generate(64, 2, 2, 2, 1, seed=20261020, s0=2)
    assert np.array_equal(bands[2], b2[0][0]) and np.array_equal(b[2], b2[3][0])
    A = oracle.ab_dense(bands[0], U[0], V[0], 2, 2)
    assert np.linalg.cond(A) < 5
    x = oracle.ab_solve(bands, U, V, b, 2, 2)
    assert relerr(x[0], np.linalg.solve(A, b[0])) < 1e-14
