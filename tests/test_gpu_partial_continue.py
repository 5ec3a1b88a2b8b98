"""GPU tier: continuing a partial AlmostBanded factorisation in place (ssm_abqr_continue), the step adaptive QR repeats as it needs more
columns of an operand it has begun to factor (_almostbanded_qr!(A, tau, ncols), AlmostBandedMatrix.jl:139-172, on the trailing view)."""
import numpy as np
import pytest

import oracle
from helpers import randn_system, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ssm():
    import ssm_b200
    return ssm_b200


@pytest.fixture(autouse=True)
def _reset_variant(ctx):
    yield
    ctx.set_option("ab_variant", 0)


@pytest.mark.parametrize("variant", [0, 9])
def test_continue_partial_qr(ssm, ctx, variant):
    """ssm_abqr_continue: qr_cols(A, n0) continued in place to n1 reflectors is qr_cols(A, n1) (to rounding; the compiled shapes re-enter
    their register-resident sweep, the others the runtime-shape kernel) and matches the oracle's _almostbanded_qr!(A, tau, n1)
    (:139-172)."""
    ctx.set_option("ab_variant", variant)
    rng = np.random.default_rng(29)
    for (n, l, u, r) in [(40, 2, 2, 2), (33, 1, 0, 1), (37, 4, 4, 2), (29, 3, 1, 3), (20, 0, 2, 1), (6, 4, 4, 2)]:
        nb = 37
        bands, U, V, b = randn_system(rng, n, l, u, r, nb)
        A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
        for n0, n1 in [(1, 2), (min(l + u + 1, n - 1), n // 2 + 1), (n // 2, n - 1), (3, n), (0, n - 1), (n - 1, n - 1)]:
            n1 = max(n0, n1)
            F = ssm.qr(A, ncols=n0).continue_to(n1)
            gF, gU, gtau = F.get()
            dF, dU, dtau = ssm.qr(A, ncols=n1).get()
            tag = (n, l, u, r, n0, n1)
            scale = max(1.0, np.abs(dF).max())
            assert np.abs(gF - dF).max() <= 1e-12 * scale and np.abs(gU - dU).max() <= 1e-12 * scale and np.abs(gtau - dtau).max() <= 1e-13, tag
            for s in range(0, nb, 9):
                oF, oU, otau = oracle.ab_qr(bands[s], U[s], V[s], l, u, ncols=n1)
                assert np.abs(gF[s] - oF).max() <= 1e-12 * max(1.0, np.abs(oF).max()), tag + (s,)
                assert np.abs(gU[s] - oU).max() <= 1e-12 * max(1.0, np.abs(oU).max()), tag + (s,)
                assert np.abs(gtau[s] - otau).max() <= 1e-13 and np.all(gtau[s][n1:] == 0.0), tag + (s,)
        # in steps, then solve with the completed factorisation
        F = ssm.qr(A, ncols=2)
        for n1 in sorted({min(5, n - 1), n // 2, n - 1}):
            F.continue_to(n1)
        x = F.ldiv(b.copy())
        x0 = ssm.qr(A).ldiv(b.copy())
        assert max(relerr(x[s], x0[s]) for s in range(nb)) <= 1e-8, (n, l, u, r)
        with pytest.raises(Exception):
            ssm.qr(A, ncols=5).continue_to(4)


def _leading(bands, U, V, n0, l, u):
    """operands of the leading n0 x n0 blocks of a batch (what the cached operand held before it grew)"""
    b0 = bands[:, :n0].copy()
    for j in range(n0):
        for d in range(l + u + 1):
            if j + d - u >= n0:
                b0[:, j, d] = 0.0
    return b0, U[:, :, :n0].copy(), V[:, :n0].copy()


@pytest.mark.parametrize("variant", [0, 9])
def test_grow_in_three_steps_equals_one_shot_qr(ssm, ctx, variant):
    """Adaptive QR across growth steps (AlmostBandedMatrix.jl:357-394 + :139-172): factor a few columns, let the operand grow (ssm_ab_resize +
    ssm_ab_set_rows), grow the factor object with it (ssm_abqr_resize re-derives what reaches into the new columns from the stored reflectors),
    continue — three times.  The result equals the one-shot qr of the final matrix and the oracle within 1e-13; so do the solutions."""
    ctx.set_option("ab_variant", variant)
    rng = np.random.default_rng(31)
    for (l, u, r) in [(2, 2, 2), (1, 0, 1), (4, 4, 2), (3, 1, 3), (3, 3, 2), (0, 2, 1)]:
        nb = 35
        sizes = [3 * (l + u) + 8, 5 * (l + u) + 17, 8 * (l + u) + 30, 9 * (l + u) + 41]
        n3 = sizes[-1]
        bands, U, V, b = randn_system(rng, n3, l, u, r, nb)
        bands[:, :, u] += 4.0
        # columns factored at each size: some well inside, some as close to the edge as growth allows (n - l)
        cols = [sizes[0] - l, sizes[1] - 2 * l - u - 1, sizes[2] - l - 1, n3 - 1]
        A = F = None
        for n, k in zip(sizes, cols):
            b0, U0, V0 = _leading(bands, U, V, n, l, u)
            if A is None:
                A = ssm.AlmostBandedMatrix(b0, (U0, V0), (l, u), ctx=ctx)
                F = ssm.qr(A, ncols=k)
            else:
                n_old = A.n
                A = A.resize(n).set_rows(max(0, n_old - u - l), n, b0, (U0, V0))
                F.resize(A)
                gF, gU, gtau = F.get()                                  # == qr_cols(A, previous k)
                dF, dU, dtau = ssm.qr(A, ncols=F.ncols).get()
                scale = max(1.0, np.abs(dF).max())
                assert np.abs(gF - dF).max() <= 1e-13 * scale and np.abs(gU - dU).max() <= 1e-13 * scale and np.abs(gtau - dtau).max() <= 1e-13, (l, u, r, n)
                F.continue_to(k)
        gF, gU, gtau = F.get()
        dF, dU, dtau = ssm.qr(A).get()
        scale = max(1.0, np.abs(dF).max())
        assert np.abs(gF - dF).max() <= 1e-13 * scale and np.abs(gtau - dtau).max() <= 1e-13, (l, u, r)
        for s in range(0, nb, 7):
            oF, oU, otau = oracle.ab_qr(bands[s], U[s], V[s], l, u)
            assert np.abs(gF[s] - oF).max() <= 1e-13 * max(1.0, np.abs(oF).max()), (l, u, r, s)
            assert np.abs(gtau[s] - otau).max() <= 1e-13
        x = F.ldiv(b.copy())
        ox = oracle.ab_solve(bands, U, V, b, l, u)
        assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12, (l, u, r)


def test_grow_refuses_reflectors_cut_off_by_the_old_size(ssm, ctx):
    n, l, u, r, nb = 30, 2, 2, 2, 4
    rng = np.random.default_rng(32)
    bands, U, V, b = randn_system(rng, n, l, u, r, nb)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    F = ssm.qr(A)                                   # n - 1 reflectors: the last l of them saw a truncated column
    with pytest.raises(ssm.SSMError, match="cut off"):
        F.resize(A.resize(n + 5))
    with pytest.raises(ssm.DimensionMismatch):
        ssm.qr(A.resize(n + 5), ncols=3).resize(A)  # cannot shrink
