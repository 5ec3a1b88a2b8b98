"""CPU tier: independent pins of the BandedPlusSemiseparable oracle (oracle/bps_oracle.c) beyond LAPACK:

(i)   after EVERY step j of the sweep, the state (Q, K, Es, Xs, Ys, Zs) satisfies the invariant the reference states at
      semiseparableqr.jl:1-12 and checks in test/test_getindex.jl:32-38 — the trailing block rebuilt from the state equals the trailing block of
      the dense partial Householder factorisation;
(ii)  factors and tau against dense Householder QR in 50-digit mpmath arithmetic at the shape of the reference's own test (test_qr.jl:6-12) and
      in numpy longdouble at n = 256 on randn data (tests/golden/make_bps_golden.py wrote both, inputs included);
(iii) the drift SURVEY H4 predicted for ill-conditioned randn input grows with cond(A), not with n: oracle vs LAPACK at n = 1024."""
import json
from pathlib import Path

import numpy as np
import pytest

import oracle
from helpers import bps_randn

GOLD = Path(__file__).parent / "golden"


def partial_householder(A, ncols):
    """H_ncols ... H_1 A (LinearAlgebra.reflector! convention)"""
    M = A.copy()
    n = A.shape[0]
    for k in range(ncols):
        x = M[k:, k].copy()
        nrm = np.linalg.norm(x)
        if nrm == 0.0:
            continue
        nu = np.copysign(nrm, x[0])
        xi = x[0] + nu
        v = x / xi; v[0] = 1.0
        M[k:, :] -= np.outer(v, (xi / nu) * (v @ M[k:, :]))
    return M


@pytest.mark.parametrize("n,l,m,r,p", [(20, 4, 5, 2, 3), (10, 2, 3, 4, 5), (24, 3, 3, 2, 2), (16, 1, 2, 1, 1), (12, 0, 2, 2, 1), (9, 2, 0, 1, 2), (5, 4, 5, 2, 3)])
def test_invariant_after_every_step(n, l, m, r, p):
    rng = np.random.default_rng(40 + n + l)
    B, U, V, W, S, _ = bps_randn(rng, n, l, m, r, p, dominant=2.0)
    A0 = oracle.bps_dense(B, U, V, W, S, l, m)
    Ud, Sd = U.T, S.T                                            # n x r, n x p
    for j in range(0, n):                                        # j columns done
        _, _, _, _, tau, st = oracle.bps_qr_steps(B, U, V, W, S, l, m, j)
        lm, lx = min(l + m, n), min(l, n)
        nt = n - j                                               # trailing block size
        Ut, St, At = Ud[j:], Sd[j:], A0[j:, j:]
        E = np.zeros((r, nt)); c = min(lm, nt); E[:, :c] = st["E"][:, :c]
        X = np.zeros((nt, p)); rws = min(lx, nt); X[:rws] = st["X"][:rws]
        Y = np.zeros((nt, r)); Y[:rws] = st["Y"][:rws]
        Z = np.zeros((nt, nt)); Z[:rws, :c] = st["Z"][:rws, :c]
        T = At + Ut @ st["Q"] @ St.T + Ut @ st["K"] @ Ut.T @ At + Ut @ E + X @ St.T + Y @ Ut.T @ At + Z
        ref = partial_householder(A0, j)[j:, j:]
        assert np.abs(T - ref).max() <= 1e-12 * np.abs(A0).max(), (j, np.abs(T - ref).max())


def test_against_50_digit_householder_at_the_reference_test_shape():
    g = json.loads((GOLD / "bps_qr_mp_n20.json").read_text())
    l, m = g["l"], g["m"]
    B, U, V, W, S = (np.array(g[k]) for k in "BUVWS")
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, m)
    F = oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, l + m)
    ref, t = np.array(g["factors"]), np.array(g["tau"])
    cond = np.linalg.cond(oracle.bps_dense(B, U, V, W, S, l, m))
    err = np.abs(F - ref).max() / np.abs(ref).max()
    lap, tl = oracle.dense_qr_unblocked(oracle.bps_dense(B, U, V, W, S, l, m))
    err_lapack = np.abs(lap - ref).max() / np.abs(ref).max()
    # the O(n) sweep is as close to the exact factors as LAPACK's dense QR is (both backward stable: error ~ eps * cond)
    assert err <= 2e-16 * cond * 20 and err <= 50 * max(err_lapack, 1e-16), (err, err_lapack, cond)
    assert np.abs(tau - t).max() <= 2e-16 * cond * 20


def test_against_extended_precision_householder_n256_randn():
    g = np.load(GOLD / "bps_qr_ld_n256.npz")
    n, l, m, r, p = (int(v) for v in g["shape"])
    B, U, V, W, S = (g[k] for k in "BUVWS")
    A = oracle.bps_dense(B, U, V, W, S, l, m)
    cond = np.linalg.cond(A)
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, m)
    F = oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, l + m)
    ref, t = g["factors"], g["tau"]
    err = np.abs(F - ref).max() / np.abs(ref).max()
    lap, _ = oracle.dense_qr_unblocked(A)
    err_lapack = np.abs(lap - ref).max() / np.abs(ref).max()
    assert err <= 1e-15 * cond and err <= 100 * max(err_lapack, 1e-16), (err, err_lapack, cond)
    assert np.abs(tau - t).max() <= 1e-15 * cond
    # the solve built on the factors: relative error against the dense solve stays at the eps * cond level
    b = g["b"]
    x = oracle.bps_upper_ldiv(Bf, Wf, Sf, oracle.bps_lmul_qt(Bf, U, Vf, tau, b, l, l + m), l, l + m)
    x0 = np.linalg.solve(A, b)
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) <= 1e-15 * cond


def test_drift_scales_with_conditioning_not_with_n():
    """on randn input the O(n) sweep drifts from LAPACK by ~eps * cond; on the diagonally dominant generator it stays at ~1e-15."""
    n, l, m, r, p = 1024, 3, 3, 2, 2
    rng = np.random.default_rng(9)
    out = {}
    for name, dom in (("randn", 0.0), ("dominant", 8.0)):
        B, U, V, W, S, _ = bps_randn(rng, n, l, m, r, p, dominant=dom)
        if dom:
            U /= np.sqrt(r * n); V /= np.sqrt(r * n); W /= np.sqrt(p * n); S /= np.sqrt(p * n)
        A = oracle.bps_dense(B, U, V, W, S, l, m)
        Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, m)
        F = oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, l + m)
        lap, _ = oracle.dense_qr_unblocked(A)
        out[name] = (np.abs(np.triu(F) - np.triu(lap)).max() / np.abs(lap).max(), np.linalg.cond(A))
    assert out["dominant"][0] <= 1e-13 and out["dominant"][1] < 10
    assert out["randn"][0] <= 1e-14 * out["randn"][1], out
