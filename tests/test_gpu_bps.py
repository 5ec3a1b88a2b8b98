"""GPU tier: the CUDA BandedPlusSemiseparable path through the C ABI against the CPU oracle and dense LAPACK.
Tolerances: factors / tau <= 1e-13 and solution relerr <= 1e-12 on the well-conditioned 'dd' inputs; scaled by
cond(A) on randn inputs (the reference's own tests use sqrt(eps))."""
import numpy as np
import pytest

import oracle
from helpers import bps_randn, explicit_qt, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture()
def bps():
    from ssm_b200 import bps as mod
    return mod


@pytest.fixture(autouse=True)
def _reset(ctx):
    yield
    ctx.set_option("ab_variant", 0)
    ctx.set_option("pipeline_stages", 0)


SHAPES = [(3, 3, 2, 2), (4, 5, 2, 3), (2, 2, 2, 2), (1, 1, 1, 1), (3, 0, 1, 1), (0, 3, 2, 2), (0, 0, 1, 1), (2, 1, 1, 2), (2, 2, 3, 1)]


def test_generate_matches_oracle_bitwise(bps, ctx):
    n, l, m, r, p, nb = 50, 3, 3, 2, 2, 40
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, seed=20261021, ctx=ctx)
    B, U, V, W, S, b = oracle.bps_generate(n, l, m, r, p, nb, seed=20261021)
    import ssm_b200
    x = ssm_b200.Vec(ctx, n, nb).generate(20261021, array_id=5).get()
    assert np.array_equal(x, b)
    A2 = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
    y1, y2 = A.mul(b), A2.mul(b)
    assert np.array_equal(y1, y2)


@pytest.mark.parametrize("variant", [1, 2, 3])
@pytest.mark.parametrize("l,m,r,p", SHAPES)
def test_qr_and_solves_match_oracle(bps, ctx, variant, l, m, r, p):
    ctx.set_option("ab_variant", variant)
    for n, nb in [(96, 70), (9, 33), (2, 3), (1, 2)]:
        B, U, V, W, S, b = oracle.bps_generate(n, l, m, r, p, nb, seed=31 + n)
        A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
        F = bps.qr(A)
        gB, gU, gV, gW, gS, gt = F.get()
        oB, oV, oW, oS, ot = oracle.bps_qr_batch(B, U, V, W, S, l, m)
        e = lambda a, r_: np.abs(a - r_).max() / max(1.0, np.abs(r_).max()) if a.size else 0.0
        assert e(gB, oB) <= 1e-13 and e(gV, oV) <= 1e-13 and e(gW, oW) <= 1e-13 and e(gS, oS) <= 1e-13, (n, nb)
        assert np.array_equal(gU, U)
        assert np.abs(gt - ot).max() <= 1e-13
        c = F.Q.T.lmul(b)
        oc = np.stack([oracle.bps_lmul_qt(oB[s], U[s], oV[s], ot[s], b[s], l, l + m) for s in range(nb)])
        assert e(c, oc) <= 1e-13
        x = F.R.solve(c)
        ox = oracle.bps_solve_batch(oB, U, oV, oW, oS, ot, b, l, l + m)
        assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12
        assert max(relerr(F.ldiv(b)[s], ox[s]) for s in range(nb)) <= 1e-12
        rr = A.residual(x, b)
        assert rr.max() <= 1e-13
        # A*b and UpperTriangular(A)\b on the unfactored matrix
        y = A.mul(b)
        oy = np.stack([oracle.bps_mul(B[s], U[s], V[s], W[s], S[s], b[s], l, m) for s in range(nb)])
        assert e(y, oy) <= 1e-13
        xu = A.upper_ldiv(b)
        oxu = np.stack([oracle.bps_upper_ldiv(B[s], W[s], S[s], b[s], l, m) for s in range(nb)])
        assert max(relerr(xu[s], oxu[s]) for s in range(nb)) <= 1e-12


def test_reference_test_qr_shape_vs_dense(bps, ctx):
    """test_qr.jl:6-53 re-instantiated: n=20, (l,m,r,p)=(4,5,2,3), randn data, against dense LAPACK."""
    rng = np.random.default_rng(11)
    n, l, m, r, p, nb = 20, 4, 5, 2, 3, 8
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p, nb)
    A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
    F = bps.qr(A)
    gB, gU, gV, gW, gS, gt = F.get()
    c = F.Q.T.lmul(b)
    x = F.R.solve(c)
    for s in range(nb):
        Ad = oracle.bps_dense(B[s], U[s], V[s], W[s], S[s], l, m)
        qr_, t = oracle.dense_qr_unblocked(Ad)
        cond = np.linalg.cond(Ad)
        Fd = oracle.bps_factors_dense(gB[s], gU[s], gV[s], gW[s], gS[s], l, l + m)
        assert np.abs(Fd - qr_).max() <= 1e-13 * cond * np.abs(qr_).max()      # fact_true.factors ≈ fact.factors (:17,29)
        assert np.abs(gt[s] - t).max() <= 1e-13 * cond                          # fact_true.τ ≈ fact.τ (:18,30)
        assert np.abs(c[s] - explicit_qt(qr_, t, b[s])).max() <= 1e-12 * cond   # lmul!(Q',b) (:37-48)
        assert relerr(x[s], np.linalg.solve(Ad, b[s])) <= 1e-13 * cond          # ldiv!(R,b) (:50-53)


def test_errors(bps, ctx):
    import ssm_b200
    rng = np.random.default_rng(3)
    B, U, V, W, S, b = bps_randn(rng, 12, 2, 2, 1, 1, 3)
    with pytest.raises(ssm_b200.DimensionMismatch):
        bps.BandedPlusSemiseparableMatrix(B, (U, V[:, :, :11]), (W, S), (2, 2), ctx=ctx)   # "Dimensions are not compatible."
    with pytest.raises(ssm_b200.SSMError):
        bps.BandedPlusSemiseparableMatrix.generate(12, 6, 2, 1, 1, 3, seed=1, ctx=ctx)       # no compiled shape


def test_full_size_config3_properties(bps, ctx):
    """BASELINE.json configs[2]: BPS QR, n=4096, l=m=3, r=p=2, 16,384 systems: residual of every system after
    Q'b and the R solve; sampled factors vs the oracle."""
    import ssm_b200
    n, l, m, r, p, nb, seed = 4096, 3, 3, 2, 2, 16384, 20261019 + 2
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, seed, ctx=ctx)
    b = ssm_b200.Vec(ctx, n, nb).generate(seed, array_id=5)
    x = ssm_b200.Vec(ctx, n, nb)
    F = bps.qr(A)
    F.solve(b, out=x)
    rr = A.residual(x, b)
    assert np.isfinite(rr).all() and rr.max() <= 1e-13, rr.max()
    xh = x.get()
    for s in (0, 1, 8191, 16383):
        oB_, oU, oV_, oW_, oS_, ob = (a[0] for a in oracle.bps_generate(n, l, m, r, p, 1, seed, s0=s))
        fB, fV, fW, fS, ft = oracle.bps_qr(oB_, oU, oV_, oW_, oS_, l, m)
        ox = oracle.bps_upper_ldiv(fB, fW, fS, oracle.bps_lmul_qt(fB, oU, fV, ft, ob, l, l + m), l, l + m)
        assert relerr(xh[s], ox) <= 1e-12


@pytest.mark.parametrize("n,l,r", [(20, 3, 2), (33, 2, 2), (17, 1, 1), (64, 4, 3), (25, 2, 1), (12, 3, 2)])
def test_rq_mul_matches_oracle_and_dense(bps, ctx, n, l, r):
    """rq_mul(F.factors, F.tau) (semiseparableqr.jl:282-287; test_rq.jl:26-34): R*Q of a symmetric BPS batch == the numpy
    restatement of the reference step and == dense R0*Q; the result is again a symmetric BPS (tested through its operands)."""
    from oracle import rq_oracle
    from test_oracle_rq import rq_true, sym_bps
    rng = np.random.default_rng(n * 7 + l)
    nb = 37
    B, U, V, W, S = sym_bps(rng, n, l, r, nb=nb)
    A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, l), ctx=ctx)
    F = bps.qr(A)
    RQ = F.rq_mul()
    assert (RQ.l, RQ.m, RQ.r, RQ.p) == (l, l, r, r)
    gB, gU, gV, gW, gS = RQ.get()
    assert np.array_equal(gW, gV) and np.array_equal(gS, gU)
    for s in range(nb):
        want, (Bf, Vf, Wf, Sf, tau) = rq_true(B[s], U[s], V[s], W[s], S[s], l)
        oB, oU, oV = rq_oracle.rq_mul(Bf, U[s], Vf, Wf, Sf, tau, l, 2 * l)
        scale = np.abs(want).max()
        assert np.abs(gB[s] - oB).max() <= 1e-13 * scale and np.abs(gU[s] - oU).max() <= 1e-13 * scale and np.abs(gV[s] - oV).max() <= 1e-13 * scale
        assert np.abs(rq_oracle.rq_dense(gB[s], gU[s], gV[s], l) - want).max() <= 1e-12 * scale


def test_rq_mul_errors(bps, ctx):
    import ssm_b200
    B, U, V, W, S, _ = oracle.bps_generate(16, 2, 2, 2, 2, 3, seed=5)          # not symmetric: S != U
    F = bps.qr(bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (2, 2), ctx=ctx))
    with pytest.raises(ssm_b200.SSMError):
        F.rq_mul()
    B, U, V, W, S, _ = oracle.bps_generate(16, 2, 3, 2, 2, 3, seed=5)          # m != l
    F = bps.qr(bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (2, 3), ctx=ctx))
    with pytest.raises(ssm_b200.DimensionMismatch):
        F.rq_mul()


def test_random_shapes_match_oracle(bps, ctx):
    """Sweep of 40 drawn shapes (l <= 4, m <= 5, r, p <= 3: everything the compiled instances cover by zero padding) through the
    default kernel selection: factors, tau, Q'b and solutions against the oracle."""
    ctx.set_option("ab_variant", 0)
    rng = np.random.default_rng(2027)
    for it in range(40):
        l, m, r, p = int(rng.integers(0, 5)), int(rng.integers(0, 6)), int(rng.integers(1, 4)), int(rng.integers(1, 4))
        n = max(1, int(rng.choice([1, 2, 3, l + m, l + m + 1, l + m + 2, 17, 64, 150])))
        nb = int(rng.choice([1, 5, 32, 33, 70]))
        B, U, V, W, S, b = oracle.bps_generate(n, l, m, r, p, nb, seed=2000 + it)
        A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
        F = bps.qr(A)
        gB, gU, gV, gW, gS, gt = F.get()
        oB, oV, oW, oS, ot = oracle.bps_qr_batch(B, U, V, W, S, l, m)
        e = lambda a, r_: np.abs(a - r_).max() / max(1.0, np.abs(r_).max()) if a.size else 0.0
        tag = (it, n, l, m, r, p, nb)
        assert e(gB, oB) <= 1e-13 and e(gV, oV) <= 1e-13 and e(gW, oW) <= 1e-13 and e(gS, oS) <= 1e-13, tag
        assert np.abs(gt - ot).max() <= 1e-13, tag
        x = F.R.solve(F.Q.T.lmul(b))
        ox = oracle.bps_solve_batch(oB, U, oV, oW, oS, ot, b, l, l + m)
        assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12, tag
