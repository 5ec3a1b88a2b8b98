"""CPU tier: the oracles over the whole supported shape space (property tests).  Every drawn shape — including systems smaller than
their bandwidths and rank-0 fill — must reproduce LAPACK's unblocked Householder QR (factors, tau) and the dense solve, which is how
the reference's own tests pin the path (test_almostbanded.jl:145,153; test_qr.jl:16-30)."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from helpers import R_from_factors, randn_system, relerr


@settings(max_examples=300, deadline=None, derandomize=True)
@given(n=st.integers(1, 48), l=st.integers(0, 8), u=st.integers(0, 8), r=st.integers(0, 4), seed=st.integers(0, 2**16))
def test_almostbanded_oracle_equals_dense_householder(n, l, u, r, seed):
    rng = np.random.default_rng(seed)
    bands, U, V, b = randn_system(rng, n, l, u, r)
    bands[:, u] += 3.0 * (l + u + 1) * np.sign(bands[:, u] + 0.5)        # keep the drawn systems well conditioned
    A = oracle.ab_dense(bands, U, V, l, u)
    F, Uf, tau = oracle.ab_qr(bands, U, V, l, u)
    x = oracle.ab_solve(bands, U, V, b, l, u)
    if l + u == 0 and r > 0:
        return      # reference quirk: with ubar = 0 the `jr1[1] < n` guard of _almostbanded_upper_ldiv! (:229,233) drops the last fill coupling
    assert relerr(x, np.linalg.solve(A, b)) <= 1e-11
    if l == 0:
        return                                                           # reference quirk: tau = 2 and a negated R row, solutions agree
    qr_, t = oracle.dense_qr_unblocked(A)
    R = R_from_factors(F, Uf, V, l, l + u)
    scale = np.abs(qr_).max()
    assert np.abs(R - np.triu(qr_)).max() <= 1e-12 * scale
    assert np.abs(np.tril(oracle.storage_to_dense(F, l, l + u), -1) - np.tril(qr_, -1)).max() <= 1e-12
    assert np.abs(tau - t).max() <= 1e-12


@settings(max_examples=150, deadline=None, derandomize=True)
@given(n=st.integers(2, 40), l=st.integers(1, 4), m=st.integers(1, 4), r=st.integers(1, 3), p=st.integers(1, 3), seed=st.integers(0, 2**16))
def test_bps_oracle_equals_dense_householder(n, l, m, r, p, seed):
    B, U, V, W, S, b = (a[0] for a in oracle.bps_generate(n, l, m, r, p, 1, seed))
    A = oracle.bps_dense(B, U, V, W, S, l, m)
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, m)
    qr_, t = oracle.dense_qr_unblocked(A)
    Fd = oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, l + m)
    assert np.abs(Fd - qr_).max() <= 1e-12 * np.abs(qr_).max()
    assert np.abs(tau - t).max() <= 1e-12
    c = oracle.bps_lmul_qt(Bf, U, Vf, tau, b.copy(), l, l + m)
    x = oracle.bps_upper_ldiv(Bf, Wf, Sf, c, l, l + m)
    assert relerr(x, np.linalg.solve(A, b)) <= 1e-11
