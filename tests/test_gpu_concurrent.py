"""GPU tier: several library contexts (= CUDA streams) working on ONE GPU at the same time.

Regression for a cross-proxy hazard in the bulk-async rings (csrc/stream_loader.cuh): a stage was re-armed (async-proxy write by
the bulk copy) without a `fence.proxy.async` after the warp's generic-proxy reads of it.  Alone on the GPU the copy never landed
early enough to matter and every sequential test passed; with ANY other stream keeping the GPU busy (even a plain device copy)
the sweeps returned wrong rows for a third of the tiles.  The checks are bit-for-bit against the same calls run alone."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ssm():
    import ssm_b200
    return ssm_b200


def _ctx_on_new_stream(ssm, torch):
    st = torch.cuda.Stream()
    return st, ssm.Context(torch.cuda.current_device(), stream=st.cuda_stream)


@pytest.mark.parametrize("variant", [0, 2, 4])
@pytest.mark.parametrize("other", ["copy", "qr_direct", "solve_group"])
def test_sweeps_are_exact_while_another_stream_is_busy(ssm, variant, other):
    import torch
    n, nb = 256, 65536
    st0, c0 = _ctx_on_new_stream(ssm, torch)
    if variant:
        c0.set_option("ab_variant", variant)
    A = ssm.AlmostBandedMatrix.generate(n, 2, 2, 2, nb, 20261023, ctx=c0)
    F = ssm.qr(A)
    b = ssm.Vec(c0, n, nb).generate(20261023)
    x = ssm.Vec(c0, n, nb)

    def work():
        F.refactor(A)
        F.solve(b, out=x)      # Q'b out of place, back-substitution in place
        F.Q.T.lmul(x)          # in place
        F.R.solve(x)           # in place

    for _ in range(3):
        work()
    c0.sync()
    ref = x.get().copy()

    st1, c1 = _ctx_on_new_stream(ssm, torch)
    t1 = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
    t2 = torch.empty_like(t1)
    if other != "copy":
        c1.set_option("ab_variant", 1 if other == "qr_direct" else 4)
        n1, nb1 = (512, 32768) if other == "qr_direct" else (2048, 4096)
        A1 = ssm.AlmostBandedMatrix.generate(n1, 2, 2, 2, nb1, 5, ctx=c1)
        F1 = ssm.qr(A1)
        b1 = ssm.Vec(c1, n1, nb1).generate(5)
        x1 = ssm.Vec(c1, n1, nb1)
        c1.sync()

    def noise():
        if other == "copy":
            with torch.cuda.stream(st1):
                t2.copy_(t1)
        elif other == "qr_direct":
            F1.refactor(A1)
        else:
            F1.solve(b1, out=x1)

    for rep in range(2):
        x.generate(1)          # scribble, so a stale result cannot pass
        c0.sync()
        for _ in range(3):
            noise(); work(); noise()
        torch.cuda.synchronize()
        assert np.array_equal(x.get(), ref), (variant, other, rep)


def test_bps_sweeps_are_exact_while_another_stream_is_busy(ssm):
    import torch
    from ssm_b200 import bps
    n, l, m, r, p, nb = 512, 3, 3, 2, 2, 8192
    st0, c0 = _ctx_on_new_stream(ssm, torch)
    A = bps.BandedPlusSemiseparableMatrix.generate(n, l, m, r, p, nb, 20261021, ctx=c0)
    b = ssm.Vec(c0, n, nb).generate(20261021, array_id=5)
    x = ssm.Vec(c0, n, nb)
    F = bps.qr(A)
    F.solve(b, out=x)
    c0.sync()
    ref = x.get().copy()
    reff = [a.copy() for a in F.get()]
    st1 = torch.cuda.Stream()
    t1 = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
    t2 = torch.empty_like(t1)
    for rep in range(2):
        x.generate(1)
        c0.sync()
        for _ in range(3):
            with torch.cuda.stream(st1):
                t2.copy_(t1)
            F.refactor(A)
            F.solve(b, out=x)
            with torch.cuda.stream(st1):
                t1.copy_(t2)
        torch.cuda.synchronize()
        assert np.array_equal(x.get(), ref)
        assert all(np.array_equal(a, c) for a, c in zip(F.get(), reff))


@pytest.mark.parametrize("tiled", [False, True])
def test_chunked_solve_is_exact_while_another_stream_is_busy(ssm, tiled):
    """K10 (one large system): stage 1 reads its entering rows through a bulk-async ring, the tiled stage 3 streams mini-tiles
    through another; the solution must not change by a bit when a second stream keeps the GPU busy."""
    import torch
    from ssm_b200 import chunked
    n, l, u, r, seed = 1 << 21, 4, 4, 2, 20261019 + 3
    st0, c0 = _ctx_on_new_stream(ssm, torch)
    c0.set_option("abc_tiled_min_chunks", 1 if tiled else 1 << 30)
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, seed, dist=1, ctx=c0)
    with torch.cuda.stream(st0):
        b = torch.empty(n, dtype=torch.float64, device="cuda")
        x = torch.empty_like(b)
    A.generate_rhs(b, seed)
    for _ in range(3):                   # the stage 1 of tiled (8, 2) systems changes kernel at the second solve of the same operands
        A.solve(b, out=x, nchunks=4096)  # (lanes = columns, then one lane per chunk through mbarrier-coupled warps): the reference is the steady state
    c0.sync()
    assert A.residual(x, b) <= 1e-13
    ref = x.cpu().numpy().copy()
    st1 = torch.cuda.Stream()
    t1 = torch.empty(1 << 27, dtype=torch.float64, device="cuda")
    t2 = torch.empty_like(t1)
    torch.cuda.synchronize()
    for rep in range(2):
        with torch.cuda.stream(st0):
            x.fill_(float(rep + 1))      # scribble, so a stale result cannot pass
        c0.sync()
        for _ in range(3):
            with torch.cuda.stream(st1):
                t2.copy_(t1)
            A.solve(b, out=x, nchunks=4096, sync=False)
            with torch.cuda.stream(st1):
                t1.copy_(t2)
        torch.cuda.synchronize()
        assert np.array_equal(x.cpu().numpy(), ref), (tiled, rep)
