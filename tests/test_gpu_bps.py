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


def _gpu_dense_factors(bps, ctx, B, U, V, W, S, l, m):
    A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
    F = bps.qr(A)
    gB, gU, gV, gW, gS, gt = (a[0] for a in F.get())           # a batch of one
    return A, F, oracle.bps_factors_dense(gB, gU, gV, gW, gS, l, l + m), gt


def test_gpu_factors_against_high_precision_householder(bps, ctx):
    """The device factors against references that owe nothing to the oracle: dense Householder QR in 50-digit arithmetic at the reference's test
    shape (test_qr.jl:6-12) and in extended precision at n = 256 on randn data (tests/golden/make_bps_golden.py)."""
    import json
    from pathlib import Path
    gold = Path(__file__).parent / "golden"
    g = json.loads((gold / "bps_qr_mp_n20.json").read_text())
    l, m = g["l"], g["m"]
    B, U, V, W, S = (np.array(g[k]) for k in "BUVWS")
    _, _, F, tau = _gpu_dense_factors(bps, ctx, B, U, V, W, S, l, m)
    ref, t = np.array(g["factors"]), np.array(g["tau"])
    cond = np.linalg.cond(oracle.bps_dense(B, U, V, W, S, l, m))
    assert np.abs(F - ref).max() / np.abs(ref).max() <= 4e-15 * cond and np.abs(tau - t).max() <= 4e-15 * cond
    g = np.load(gold / "bps_qr_ld_n256.npz")
    n, l, m, r, p = (int(v) for v in g["shape"])
    B, U, V, W, S, b = (g[k] for k in ("B", "U", "V", "W", "S", "b"))
    Amat = oracle.bps_dense(B, U, V, W, S, l, m)
    cond = np.linalg.cond(Amat)
    A, Fq, F, tau = _gpu_dense_factors(bps, ctx, B, U, V, W, S, l, m)
    err = np.abs(F - g["factors"]).max() / np.abs(g["factors"]).max()
    assert err <= 1e-15 * cond and np.abs(tau - g["tau"]).max() <= 1e-15 * cond, (err, cond)
    x = Fq.ldiv(b.copy())
    assert relerr(x, np.linalg.solve(Amat, b)) <= 1e-15 * cond


def test_randn_n4096_drift_scales_with_cond(bps, ctx, record_property):
    """randn (ill-conditioned) BPS input at the config-3 size.  Device factors against the oracle and against LAPACK's dense QR, the
    tolerance scaled by the measured condition number; the drift is recorded (printed with -rA / in the junit properties)."""
    n, l, m, r, p, nb = 4096, 3, 3, 2, 2, 3
    rng = np.random.default_rng(4096)
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p, nb)
    A = bps.BandedPlusSemiseparableMatrix(B, (U, V), (W, S), (l, m), ctx=ctx)
    F = bps.qr(A)
    gB, gU, gV, gW, gS, gt = F.get()
    x = F.ldiv(b.copy())
    oB, oV, oW, oS, ot = oracle.bps_qr_batch(B, U, V, W, S, l, m)
    ox = oracle.bps_solve_batch(oB, U, oV, oW, oS, ot, b, l, l + m)
    s = 0
    Amat = oracle.bps_dense(B[s], U[s], V[s], W[s], S[s], l, m)
    cond = float(np.linalg.cond(Amat))
    lap, tl = oracle.dense_qr_unblocked(Amat)
    Fg = oracle.bps_factors_dense(gB[s], gU[s], gV[s], gW[s], gS[s], l, l + m)
    Fo = oracle.bps_factors_dense(oB[s], U[s], oV[s], oW[s], oS[s], l, l + m)
    d_gpu_lapack = float(np.abs(np.triu(Fg) - np.triu(lap)).max() / np.abs(lap).max())
    d_oracle_lapack = float(np.abs(np.triu(Fo) - np.triu(lap)).max() / np.abs(lap).max())
    d_gpu_oracle = float(max(np.abs(gB[k] - oB[k]).max() / np.abs(oB[k]).max() for k in range(nb)))
    x0 = np.linalg.solve(Amat, b[s])
    e_sol = float(relerr(x[s], x0)); e_sol_oracle = float(relerr(ox[s], x0))
    for k, v in dict(cond=cond, gpu_vs_lapack=d_gpu_lapack, oracle_vs_lapack=d_oracle_lapack, gpu_vs_oracle=d_gpu_oracle, relerr_gpu=e_sol, relerr_oracle=e_sol_oracle).items():
        record_property(k, v)
    print(f"randn n=4096: cond {cond:.3e}; R factor drift vs LAPACK: gpu {d_gpu_lapack:.2e}, oracle {d_oracle_lapack:.2e}; gpu vs oracle {d_gpu_oracle:.2e}; "
          f"solution relerr vs dense: gpu {e_sol:.2e}, oracle {e_sol_oracle:.2e}")
    assert d_gpu_lapack <= 1e-14 * cond and d_gpu_oracle <= 1e-14 * cond
    assert e_sol <= 1e-14 * cond and e_sol <= 10 * max(e_sol_oracle, 1e-15)
