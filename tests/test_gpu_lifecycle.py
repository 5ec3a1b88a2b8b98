"""GPU tier: object / context lifetime and stream ordering at the C-ABI boundary (round-1 advisor findings).

* device blocks touched by ANOTHER context's stream never go back to the free list (their owner's destroy-time synchronisation does
  not cover that use);
* a context destroyed before its objects is only retired — the objects stay valid to destroy (Julia runs atexit hooks before finalizers);
* solve entry points refuse a partial factorisation (SSM_E_STATE) while lmul!(Q', .) applies the reflectors formed so far;
* a factor object keeps the matrix it factored when the operand batch is rewritten afterwards (qr(A) copies, AlmostBandedMatrix.jl:129-133);
* torch CUDA tensors handed to the library are ordered against torch's current stream.
"""
import ctypes as C

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def test_blocks_used_by_another_context_are_not_recycled():
    import ssm_b200 as ssm
    c1, c2 = ssm.Context(0), ssm.Context(0)
    n, l, u, r, nb = 96, 2, 2, 2, 40
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=11)
    ox = oracle.ab_solve(bands, U, V, b, l, u)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=c1)
    F = ssm.qr(A)
    ssm.pool_trim(0)
    v = ssm.Vec(c2, n, nb, b)              # lives on c2, used on c1's stream
    F.ldiv(v)
    x = v.get()
    assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12
    # the owner's stream is ordered behind the other context's work (device-side events), so v.get() above needs no host synchronisation:
    # repeat with a large batch, where the solve is long enough for an unordered read to see stale data
    n2, nb2 = 512, 8192
    A2 = ssm.AlmostBandedMatrix.generate(n2, l, u, r, nb2, 21, ctx=c1)
    F2 = ssm.qr(A2)
    b2 = oracle.ab_generate(n2, l, u, r, nb2, 21)[3]
    for _ in range(3):
        v2 = ssm.Vec(c2, n2, nb2, b2)
        F2.ldiv(v2)
        x2 = v2.get()
        assert A2.residual(x2, b2).max() <= 1e-13
        del v2
    del F2, A2
    ssm.pool_trim(0)
    before = ssm.pool_stats(0)["blocks"]
    del v                                   # foreign-used: must be cudaFree'd, not listed
    assert ssm.pool_stats(0)["blocks"] == before
    w = ssm.Vec(c1, n, nb, b)               # same-context use: listed as before
    F.ldiv(w)
    del w
    assert ssm.pool_stats(0)["blocks"] == before + 1
    del F, A
    c1.close(); c2.close()


def test_context_destroyed_before_its_objects():
    import ssm_b200 as ssm
    from ssm_b200 import _lib
    c = ssm.Context(0)
    n, l, u, r, nb = 64, 1, 1, 1, 8
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=12)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=c)
    F = ssm.qr(A)
    v = ssm.Vec(c, n, nb, b)
    x = F.ldiv(v).get()
    c.close()                               # retire the context first ...
    del v, F, A                             # ... then the objects: no use-after-free, blocks go back to the list or the driver
    c2 = ssm.Context(0)
    A2 = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=c2)
    x2 = ssm.qr(A2).ldiv(b.copy())
    assert np.array_equal(x, x2)
    assert _lib.lib().ssm_ctx_destroy(None) == 0


def test_solves_refuse_a_partial_factorisation(ctx):
    import ssm_b200 as ssm
    n, l, u, r, nb = 48, 2, 1, 2, 6
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=13)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    F = ssm.qr(A, ncols=10)
    for call in (lambda: F.ldiv(b.copy()), lambda: F.R.solve(b.copy()), lambda: F.ldiv_mat(np.stack([b, b], axis=1))):
        with pytest.raises(ssm.SSMError, match="reflectors"):
            call()
    # Q' of the 10 reflectors formed so far against the oracle's partial factors
    y = F.Q.T.lmul(b.copy())
    for s in range(nb):
        oF, oU, otau = oracle.ab_qr(bands[s], U[s], V[s], l, u, ncols=10)
        assert relerr(y[s], oracle.ab_lmul_qt(oF, otau, b[s], l, l + u)) <= 1e-13
    F.continue_to(n - 1)
    x = F.ldiv(b.copy())
    ox = oracle.ab_solve(bands, U, V, b, l, u)
    assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12


def test_factor_object_survives_rewriting_the_operand_batch(ctx):
    import ssm_b200 as ssm
    from ssm_b200 import bps
    n, l, u, r, nb = 80, 2, 2, 2, 33
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=14)
    ox = oracle.ab_solve(bands, U, V, b, l, u)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    F = ssm.qr(A)
    b2, U2, V2, _ = oracle.ab_generate(n, l, u, r, nb, seed=15)
    A.set_rows(0, n, b2, (U2, V2))          # rewrites V (shared with F until now)
    x = F.ldiv(b.copy())
    assert max(relerr(x[s], ox[s]) for s in range(nb)) <= 1e-12
    x2 = ssm.qr(A).ldiv(b.copy())           # and A really is the new matrix
    ox2 = oracle.ab_solve(b2, U2, V2, b, l, u)
    assert max(relerr(x2[s], ox2[s]) for s in range(nb)) <= 1e-12
    # BPS twin: the factor object shares the operand records
    n, l, m, r, p, nb = 64, 2, 2, 2, 2, 20
    B, Ub, Vb, W, S, rhs = oracle.bps_generate(n, l, m, r, p, nb, seed=16)
    Ab = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, 16, ctx=ctx)
    Fb = bps.qr(Ab)
    oB, oV, oW, oS, ot = oracle.bps_qr_batch(B, Ub, Vb, W, S, l, m)
    ox = oracle.bps_solve_batch(oB, Ub, oV, oW, oS, ot, rhs, l, l + m)
    from ssm_b200._lib import check, lib
    check(lib().ssm_bps_generate(Ab._h, 17), "ssm_bps_generate")
    xb = Fb.ldiv(rhs.copy())
    assert max(relerr(xb[s], ox[s]) for s in range(nb)) <= 1e-12


def test_torch_tensors_are_ordered_against_the_current_stream():
    import torch
    import ssm_b200 as ssm
    c = ssm.Context(0)                      # private non-blocking stream: NOT torch's current stream
    n, nb = 512, 4096
    big = torch.randn(4096, 4096, device="cuda")
    for trial in range(4):
        for _ in range(6):
            big = big @ big * 1e-3          # keep torch's stream busy so that a missing wait would read stale data
        src = torch.full((nb, n), float(trial), dtype=torch.float64, device="cuda") + torch.arange(n, device="cuda", dtype=torch.float64)
        v = ssm.Vec(c, n, nb).set(src.t().contiguous().t())        # non-contiguous view: a temporary crosses the boundary
        out = torch.empty((nb, n), dtype=torch.float64, device="cuda")
        v.get(out)
        got = out.cpu()                     # torch's stream must see the unpack kernel's result
        assert torch.equal(got, src.cpu()), trial
    c.close()
