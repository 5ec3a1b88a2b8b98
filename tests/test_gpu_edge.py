"""GPU tier: edge sizes — systems smaller than the Householder window (n = 1 .. 7 against bandwidths up to 4), through every kernel
variant of the batched path and through the chunked path, against dense LAPACK (rel. error <= 1e-12)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "emul"))
import oracle  # noqa: E402
from sof_reference import dense  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("l,u,r", [(2, 2, 2), (1, 0, 1), (4, 4, 2), (3, 1, 3), (0, 1, 1)])
def test_batched_path_tiny_systems(ctx, l, u, r):
    import ssm_b200 as ssm
    try:
        for n in (1, 2, 3, 4, 5, 7):
            for variant in (0, 1, 2, 4, 9):
                ctx.set_option("ab_variant", variant)
                nb = 5
                bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=3)
                A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
                x = ssm.qr(A).ldiv(b.copy())
                x2 = A.solve(b.copy())
                for s in range(nb):
                    x0 = np.linalg.solve(dense(bands[s], U[s], V[s], l, u), b[s])
                    assert np.linalg.norm(x[s] - x0) <= 1e-12 * np.linalg.norm(x0), (n, variant, s)
                    assert np.linalg.norm(x2[s] - x0) <= 1e-12 * np.linalg.norm(x0), (n, variant, s)
    finally:
        ctx.set_option("ab_variant", 0)


@pytest.mark.parametrize("l,u,r", [(2, 2, 2), (1, 0, 1), (4, 4, 2), (1, 1, 0)])
def test_chunked_path_tiny_systems(ctx, l, u, r):
    from ssm_b200 import chunked
    for n in (1, 2, 3, 5, 9, 17, 18, 40):
        bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=4)
        A = chunked.LargeAlmostBandedMatrix(bands[0], (U[0], V[0]), (l, u), ctx=ctx)
        x0 = np.linalg.solve(dense(bands[0], U[0], V[0], l, u), b[0])
        default = ctx.get_option("abc_tiled_min_chunks")
        try:
            for min_chunks in (default, 1):                 # 1: the tiled stage 3 of large systems, forced onto the tiny one
                ctx.set_option("abc_tiled_min_chunks", min_chunks)
                x = A.solve(b[0])
                assert np.linalg.norm(x - x0) <= 1e-12 * np.linalg.norm(x0), (n, A.chunks, min_chunks)
        finally:
            ctx.set_option("abc_tiled_min_chunks", default)


def _one_qr_solve(ssm, ctx, bands, U, V, b, l, u):
    """what the Julia glue does for qr(A) \\ b on one matrix: operand batch, factor object and vector made and destroyed per call"""
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    F = ssm.qr(A)
    x = F.ldiv(b.copy())
    fac, Uf, tau = F.get()
    del F, A
    return x, fac, tau


def test_recycled_device_blocks_give_identical_results(ctx):
    """Per-device free list (ssm_pool_*, api.cu): the second and later calls take their device blocks off the list (hits grow),
    results stay bit-identical to the first call's, and stale contents of a recycled block (here: NaN) never reach a result."""
    import gc
    import ssm_b200 as ssm
    n, l, u, r = 1000, 2, 2, 2
    bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=11)
    ssm.pool_trim(0)
    x0, fac0, tau0 = _one_qr_solve(ssm, ctx, bands, U, V, b, l, u)
    gc.collect()
    s0 = ssm.pool_stats(0)
    assert s0["blocks"] > 0 and 0 < s0["bytes"] <= 512 << 20
    for _ in range(3):
        x, fac, tau = _one_qr_solve(ssm, ctx, bands, U, V, b, l, u)
        assert np.array_equal(x, x0) and np.array_equal(fac, fac0) and np.array_equal(tau, tau0)
    gc.collect()
    s1 = ssm.pool_stats(0)
    assert s1["hits"] >= s0["hits"] + 9, (s0, s1)       # operands (2) + factors (2) + vector per call, at least
    # poison every listed block size with NaN through vectors of the same footprint, then solve again
    sizes = sorted({(n + 48) * f for f in (1, r, l + 1, l + u + 1 + r)})
    for rep in range(2):
        vs = [ssm.Vec(ctx, cnt - 48, 32, data=np.full((32, cnt - 48), np.nan)) for cnt in sizes for _ in range(2)]
        del vs
        gc.collect()
    x, fac, tau = _one_qr_solve(ssm, ctx, bands, U, V, b, l, u)
    assert np.array_equal(x, x0) and np.array_equal(fac, fac0) and np.array_equal(tau, tau0)
    xd = np.linalg.solve(dense(bands[0], U[0], V[0], l, u), b[0])
    assert np.linalg.norm(x[0] - xd) <= 1e-12 * np.linalg.norm(xd)
    # a different context (another stream) may take the same blocks; large batches never enter the list
    ctx2 = ssm.Context(0)
    x2, _, _ = _one_qr_solve(ssm, ctx2, bands, U, V, b, l, u)
    assert np.array_equal(x2, x0)
    big = ssm.Vec(ctx2, 1024, 16384)                     # 140 MB > the 64 MiB block limit
    before = ssm.pool_stats(0)["bytes"]
    del big
    gc.collect()
    assert ssm.pool_stats(0)["bytes"] == before
    ctx2.close()
    assert ssm.pool_trim(0) > 0
    s2 = ssm.pool_stats(0)
    assert s2["bytes"] == 0 and s2["blocks"] == 0
    x, _, _ = _one_qr_solve(ssm, ctx, bands, U, V, b, l, u)
    assert np.array_equal(x, x0)


@pytest.mark.parametrize("l,u,r", [(2, 1, 2), (1, 0, 1), (4, 4, 2)])
def test_resize_keeps_bands_and_fill_and_zero_pads(ctx, l, u, r):
    """resize(A, m, m) (AlmostBandedMatrix.jl:338-348) on the device: same bands and fill, new band entries / U rows / V columns zero,
    the dense matrix is [A 0; 0 0]; shrinking is a DimensionMismatch."""
    import ssm_b200 as ssm
    n, m, nb = 50, 83, 37                                   # two tiles, the second partly filled
    bands, U, V, b = oracle.ab_generate(n, l, u, r, nb, seed=5)
    A = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    B = A.resize(m)
    assert B.shape == (m, m) and B.almostbandwidths() == (l, u) and B.almostbandedrank() == r
    b2, (U2, V2) = B.bandpart(), B.fillpart()
    assert np.array_equal(b2[:, :n], bands) and not b2[:, n:].any()
    assert np.array_equal(U2[:, :, :n], U) and not U2[:, :, n:].any()
    assert np.array_equal(V2[:, :n], V) and not V2[:, n:].any()
    for s in (0, nb - 1):
        D = dense(b2[s], U2[s], V2[s], l, u)
        assert np.array_equal(D[:n, :n], dense(bands[s], U[s], V[s], l, u)) and not D[n:].any() and not D[:, n:].any()
    b0, (U0, V0) = A.bandpart(), A.fillpart()               # the operand itself is untouched
    assert np.array_equal(b0, bands) and np.array_equal(U0, U) and np.array_equal(V0, V)
    assert np.array_equal(A.resize(n).bandpart(), bands)    # same size: a copy
    with pytest.raises(ssm.DimensionMismatch):
        A.resize(n - 1)


@pytest.mark.parametrize("l,u,r", [(2, 2, 2), (1, 0, 1), (3, 1, 3)])
@pytest.mark.parametrize("on_device", [False, True])
def test_grow_then_set_rows_equals_fresh_upload(ctx, l, u, r, on_device):
    """The resizedata! flow (AlmostBandedMatrix.jl:350-394) on the device: the leading n x n part is uploaded, the batch is grown to
    m x m (ssm_ab_resize) and only the rows from n-u on are sent (ssm_ab_set_rows).  Operands, factors and solutions are then
    bit-identical to those of a batch built from the full arrays; rows outside the range keep their contents."""
    import torch
    import ssm_b200 as ssm
    n, m, nb = 40, 97, 35
    bands, U, V, b = oracle.ab_generate(m, l, u, r, nb, seed=8)
    full = ssm.AlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    lead_b = np.ascontiguousarray(bands[:, :n]); lead_U = np.ascontiguousarray(U[:, :, :n]); lead_V = np.ascontiguousarray(V[:, :n])
    for j in range(n):                                     # band entries of the leading part that lie outside the n x n matrix
        for d in range(l + u + 1):
            if not 0 <= j + d - u < n:
                lead_b[:, j, d] = 0.0
    A = ssm.AlmostBandedMatrix(lead_b, (lead_U, lead_V), (l, u), ctx=ctx).resize(m)
    k0 = max(n - u, 0)                                     # rows n-u..n-1 gain entries in the new columns
    if on_device:
        dev = [torch.from_numpy(a).cuda() for a in (bands, U, V)]
        A.set_rows(k0, m, dev[0], (dev[1], dev[2]))
    else:
        A.set_rows(k0, m, bands, (U, V))
    b1, (U1, V1) = A.bandpart(), A.fillpart()
    b0, (U0, V0) = full.bandpart(), full.fillpart()
    assert np.array_equal(b1, b0) and np.array_equal(U1, U0) and np.array_equal(V1, V0)
    x0 = ssm.qr(full).ldiv(b.copy()); x1 = ssm.qr(A).ldiv(b.copy())
    assert np.array_equal(x0, x1)
    xd = np.linalg.solve(dense(bands[3], U[3], V[3], l, u), b[3])
    assert np.linalg.norm(x1[3] - xd) <= 1e-12 * np.linalg.norm(xd)
    # a range in the middle leaves the rest alone
    junk = np.full_like(bands, 7.0), (np.full_like(U, 7.0), np.full_like(V, 7.0))
    A.set_rows(10, 13, junk[0], junk[1])
    b2, (U2, V2) = A.bandpart(), A.fillpart()
    rows = np.r_[0:10, 13:m]
    assert np.array_equal(U2[:, :, rows], U0[:, :, rows]) and np.array_equal(V2[:, rows], V0[:, rows]) and (U2[:, :, 10:13] == 7.0).all()
    for i in rows:                                         # band row i = entries (i, j), stored at bands[j, u + i - j]
        for j in range(max(i - l, 0), min(i + u, m - 1) + 1):
            assert np.array_equal(b2[:, j, u + i - j], b0[:, j, u + i - j])
    with pytest.raises(ssm.DimensionMismatch):
        A.set_rows(5, m + 1, bands, (U, V))


def test_free_list_is_bounded(ctx):
    """the per-device free list keeps at most 256 blocks (oldest evicted first), however many small objects are destroyed at once"""
    import gc
    import ssm_b200 as ssm
    ssm.pool_trim(0)
    vs = [ssm.Vec(ctx, 8 + i, 1) for i in range(300)]
    del vs
    gc.collect()
    s = ssm.pool_stats(0)
    assert 0 < s["blocks"] <= 256 and s["bytes"] <= 512 << 20, s
    assert ssm.pool_trim(0) == s["bytes"]
    assert ssm.pool_stats(0)["blocks"] == 0
