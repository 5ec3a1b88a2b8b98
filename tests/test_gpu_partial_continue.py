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
    """ssm_abqr_continue: qr_cols(A, n0) continued in place to n1 reflectors is qr_cols(A, n1) (to rounding: the continuation always runs the
    runtime-shape kernel, the first part may have run a register-resident instance) and matches the oracle's _almostbanded_qr!(A, tau, n1)
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
