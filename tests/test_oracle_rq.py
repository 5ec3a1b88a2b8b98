"""CPU tier: the numpy restatement of the reference's symmetric R*Q step (oracle/rq_oracle.py, semiseparableqr.jl:170-311,642-845)
pinned the way the reference's own test pins it (test/test_rq.jl:20-34): the result equals dense R0 * Q."""
import numpy as np
import pytest

import oracle
from oracle import rq_oracle


def sym_bps(rng, n, l, r, scale=0.3, nb=None):
    """symmetric BPS in the reference layout: B (n, 2l+1) symmetric band, W = V, S = U (test_rq.jl:6-19)"""
    def one():
        Bd = np.zeros((n, n))
        for i in range(n):
            for j in range(max(0, i - l), i + 1):
                v = rng.standard_normal(); Bd[i, j] = v; Bd[j, i] = v
            Bd[i, i] += (2 * l + 2) * np.sign(Bd[i, i])
        B = oracle.banded_to_storage(Bd, l, l)
        U = rng.standard_normal((r, n)) * scale; V = rng.standard_normal((r, n)) * scale
        return B, U, V, V.copy(), U.copy()
    if nb is None:
        return one()
    parts = [one() for _ in range(nb)]
    return tuple(np.stack([p[k] for p in parts]) for k in range(5))


def rq_true(B, U, V, W, S, l):
    A = oracle.bps_dense(B, U, V, W, S, l, l)
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, l)
    R0 = np.triu(oracle.bps_factors_dense(Bf, U, Vf, Wf, Sf, l, 2 * l))
    return R0 @ (A @ np.linalg.inv(R0)), (Bf, Vf, Wf, Sf, tau)


@pytest.mark.parametrize("n,l,r", [(20, 4, 3), (30, 2, 2), (17, 1, 1), (25, 3, 2), (12, 3, 3), (40, 4, 5)])
def test_rq_oracle_matches_dense_R0_times_Q(n, l, r):
    rng = np.random.default_rng(n + l + r)
    B, U, V, W, S = sym_bps(rng, n, l, r)
    want, (Bf, Vf, Wf, Sf, tau) = rq_true(B, U, V, W, S, l)
    Bs, Un, Vn = rq_oracle.rq_mul(Bf, U, Vf, Wf, Sf, tau, l, 2 * l)
    got = rq_oracle.rq_dense(Bs, Un, Vn, l)
    assert np.allclose(got, got.T)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()


def test_rq_oracle_refuses_unsymmetric_generators():
    rng = np.random.default_rng(0)
    B, U, V, W, S = sym_bps(rng, 12, 2, 2)
    Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, 2, 2)
    Sf[0, 3] += 1.0
    with pytest.raises(AssertionError):
        rq_oracle.rq_mul(Bf, U, Vf, Wf, Sf, tau, 2, 4)


def test_rq_steps_are_similarity_transforms():
    """Independent of the QR oracle's factors: R*Q = Q'AQ, so three steps of the iteration A -> qr -> R*Q keep the spectrum of the symmetric
    matrix (LAPACK eigvalsh of the dense matrices), the trace and the Frobenius norm, and the result stays symmetric."""
    rng = np.random.default_rng(77)
    n, l, r = 96, 3, 2
    B, U, V, W, S = sym_bps(rng, n, l, r)
    A0 = oracle.bps_dense(B, U, V, W, S, l, l)
    ev0 = np.linalg.eigvalsh((A0 + A0.T) / 2)
    assert np.abs(A0 - A0.T).max() == 0.0
    for step in range(3):
        Bf, Vf, Wf, Sf, tau = oracle.bps_qr(B, U, V, W, S, l, l)
        B, Un, Vn = rq_oracle.rq_mul(Bf, U, Vf, Wf, Sf, tau, l, 2 * l)
        U, V, W, S = Un, Vn, Vn.copy(), Un.copy()
        A = rq_oracle.rq_dense(B, U, V, l)
        assert np.abs(A - A.T).max() <= 1e-13 * np.abs(A).max()
        ev = np.linalg.eigvalsh((A + A.T) / 2)
        assert np.abs(ev - ev0).max() <= 1e-12 * np.abs(ev0).max(), (step, np.abs(ev - ev0).max())
        assert abs(np.trace(A) - np.trace(A0)) <= 1e-12 * np.abs(ev0).sum()
        assert abs(np.linalg.norm(A) - np.linalg.norm(A0)) <= 1e-12 * np.linalg.norm(A0)
