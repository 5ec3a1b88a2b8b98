"""GPU tier: one large AlmostBandedMatrix system through the chunked n-axis path (K10, ssm_abc_*), against dense LAPACK,
the numpy restatement of the same algebra (stage by stage), the reference-restating oracle, and — at BASELINE.json
configs[3]'s full size n = 2^24 — the oracle's sequential sweep plus the structured residual.
Tolerance: relative solution error <= 1e-12 (north_star)."""
import math
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "emul"))
import oracle  # noqa: E402
from helpers import readme_expz  # noqa: E402
from sof_reference import chunk_reduce, dense, pad_system, regular_layout  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture()
def chunked():
    from ssm_b200 import chunked
    return chunked


def _rand(rng, n, l, u, r):
    bands = rng.standard_normal((n, l + u + 1)); U = rng.standard_normal((r, n)); V = rng.standard_normal((n, r)) / np.sqrt(n)
    bands[:, u] += 2 * (l + u + 1) * np.sign(bands[:, u])          # well conditioned: the bar is 1e-12 on the solution
    for j in range(n):
        for d in range(l + u + 1):
            k = j + d - u
            if k < 0 or k >= n:
                bands[j, d] = 0.0
    return bands, U, V, rng.standard_normal(n)


@pytest.mark.parametrize("n,l,u,r,P", [(64, 2, 2, 2, 4), (50, 1, 0, 1, 3), (97, 4, 4, 2, 4), (40, 2, 0, 2, 2), (33, 1, 1, 1, 1),
                                       (200, 3, 1, 2, 8), (500, 2, 2, 2, 16), (777, 4, 4, 2, 7), (130, 1, 1, 2, 5), (64, 2, 1, 1, 3),
                                       (300, 2, 2, 0, 6), (90, 1, 0, 0, 3), (128, 1, 1, 0, 4)])
@pytest.mark.parametrize("tiled", [False, True])
@pytest.mark.parametrize("gs", [8, 4, 1])
def test_chunked_small_vs_dense_and_stagewise(chunked, ctx, n, l, u, r, P, tiled, gs):
    """`tiled`: stage 3 with one lane per chunk over tile-interleaved records (the path large systems take) forced onto the
    small case; otherwise one warp per chunk over row-major records.  `gs`: lanes that share a chunk in stage 1 (the wide
    active blocks, 2(l+u) + 2r + 2 >= 18 columns, have a 4-lane instance; the others always run 8 lanes per chunk)."""
    gs_eff = 4 if gs == 4 and 2 * (l + u) + 2 * r + 2 >= 18 else 8
    if gs == 4 and gs_eff == 8:
        pytest.skip("shape has no 4-lane instance")
    if gs == 1 and not (tiled and r == 2 and l + u in (4, 6, 8)):
        pytest.skip("one lane per chunk: shapes (l+u, r) = (4 | 6 | 8, 2) over tiled records")
    rng = np.random.default_rng(n * 13 + P)
    bands, U, V, b = _rand(rng, n, l, u, r)
    A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    default, default_gs, default_lpc = ctx.get_option("abc_tiled_min_chunks"), ctx.get_option("abc_gs"), ctx.get_option("abc_lpc")
    ctx.set_option("abc_tiled_min_chunks", 1 if tiled else 1 << 30)
    # gs = 1: the one-lane-per-chunk stage 1 (what large systems of these shapes run from their second solve on); 4 | 8: lanes = columns
    ctx.set_option("abc_lpc", 2 if gs == 1 else 0)
    ctx.set_option("abc_gs", 8 if gs == 1 else gs)
    try:
        x = A.solve(b, nchunks=P)
    finally:
        ctx.set_option("abc_tiled_min_chunks", default)
        ctx.set_option("abc_gs", default_gs)
        ctx.set_option("abc_lpc", default_lpc)
    lp, w = l + u, l + u + r
    P, bnd = regular_layout(n, lp, P, same=32 if tiled else 32 // gs_eff)
    assert A.chunks == P
    pb, pU, pV, pbv = pad_system(bands, U, V, b, l, u, bnd[-1])
    # stage 1 output: level-0 reduced blocks [F | G | h] per chunk, against the numpy restatement (rows may differ by
    # nothing: same reflector convention and order)
    E = A._debug(0, P * w * (2 * w + 1)).reshape(P, w, 2 * w + 1)
    for c in range(P):
        _, left = chunk_reduce(pb, pU, pV, pbv, l, u, bnd[c], bnd[c + 1])
        G0, C0, T0 = lp + 1, lp + 1 + r, lp + 1 + r + lp
        ref = np.hstack([left[:, C0:C0 + lp], left[:, T0:T0 + r], left[:, 0:lp], left[:, G0:G0 + r], left[:, -1:]])
        assert np.allclose(E[c], ref, rtol=1e-11, atol=1e-11 * np.abs(ref).max()), (c, np.abs(E[c] - ref).max())
    x0 = np.linalg.solve(dense(bands, U, V, l, u), b)
    assert np.linalg.norm(x - x0) / np.linalg.norm(x0) <= 1e-12
    assert A.residual(x, b) <= 1e-13


def test_chunked_readme_known_answer(chunked, ctx):
    (l, u, r), bands, U, V, b = readme_expz(20)
    A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    for P in (1, 2, 3):
        x = A.solve(b, nchunks=P)
        assert A.chunks == P
        ref = np.array([1 / math.factorial(k) for k in range(16)])
        # normwise: the partitioned factorisation is backward stable, not componentwise accurate on the graded tail
        assert np.linalg.norm(x[:16] - ref) / np.linalg.norm(ref) <= 1e-14


def test_chunked_readme_million(chunked, ctx):
    """BASELINE.json configs[0] at its large size (README.md:96-97, n = 1,000,000): `A \\ b` of the README system through the
    chunked path, against the known answer 1/k!, the oracle's sequential sweep and the structured residual."""
    n = 1_000_000
    (l, u, r), bands, U, V, b = readme_expz(n)
    ox = oracle.ab_solve(bands[None], U[None], V[None], b[None], l, u)[0]
    A = chunked.LargeAlmostBandedMatrix(bands, (U, V), (l, u), ctx=ctx)
    x = A.solve(b)
    assert A.chunks > 1000
    ref = np.array([1 / math.factorial(k) for k in range(16)])
    assert np.linalg.norm(x[:16] - ref) / np.linalg.norm(ref) <= 1e-14
    assert np.linalg.norm(x - ox) / np.linalg.norm(ox) <= 1e-12
    assert A.residual(x, b) <= 1e-13


@pytest.mark.parametrize("n,l,u,r,dist", [(100_000, 4, 4, 2, 1), (100_000, 2, 2, 2, 0), (262_144, 1, 0, 1, 0), (65_537, 4, 4, 2, 0)])
def test_chunked_matches_oracle_sweep(chunked, ctx, n, l, u, r, dist):
    import torch
    seed = 20261019 + 3
    bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=seed, dist=dist)
    ox = oracle.ab_solve(bands, U, V, b, l, u)[0]
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, seed, dist=dist, ctx=ctx)
    rec = A._debug(3, 8 * (l + u + 1 + r)).reshape(8, l + u + 1 + r)          # generator = oracle's, bit for bit
    for i in range(8):
        for d in range(l + u + 1):
            j = i - l + d
            assert rec[i, d] == (bands[0, j, u + i - j] if 0 <= j < n else 0.0)
    bd = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), seed)
    assert np.array_equal(bd.cpu().numpy(), b[0])
    x = A.solve(bd)
    xs = x.cpu().numpy()
    assert np.linalg.norm(xs - ox) / np.linalg.norm(ox) <= 1e-12
    assert A.residual(x, bd) <= 1e-13
    # host-pointer entry and explicit operands give the same answer as the device-generated ones
    A2 = chunked.LargeAlmostBandedMatrix(bands[0], (U[0], V[0]), (l, u), ctx=ctx)
    x2 = A2.solve(b[0], nchunks=A.chunks)
    assert np.array_equal(x2, xs)
    # many short chunks: stage 3 runs one lane per chunk over tiled records (a partial last tile included)
    default = ctx.get_option("abc_tiled_min_chunks")
    ctx.set_option("abc_tiled_min_chunks", 2048)
    try:
        for P in (2500, 4096):
            x3 = A.solve(bd, nchunks=P).cpu().numpy()
            assert A.chunks >= 2048                            # (P itself shrinks when the chunks would get too short)
            assert np.linalg.norm(x3 - ox) / np.linalg.norm(ox) <= 1e-12, P
    finally:
        ctx.set_option("abc_tiled_min_chunks", default)


def test_config4_full_size(chunked, ctx):
    """BASELINE.json configs[3]: single large almost-banded spectral system, n = 16,777,216, l = u = 4, r = 2."""
    import torch
    n, l, u, r, seed = 1 << 24, 4, 4, 2, 20261019 + 3
    A = chunked.LargeAlmostBandedMatrix.generate(n, l, u, r, seed, dist=1, ctx=ctx)
    b = A.generate_rhs(torch.empty(n, dtype=torch.float64, device="cuda"), seed)
    x = A.solve(b)
    assert A.residual(x, b) <= 1e-13
    bands, U, V, bb = oracle.ab_generate(n, l, u, r, 1, seed=seed, dist=1)
    ox = oracle.ab_solve(bands, U, V, bb, l, u)[0]
    xs = x.cpu().numpy()
    assert np.linalg.norm(xs - ox) / np.linalg.norm(ox) <= 1e-12
    # a different chunking gives the same solution to rounding
    x2 = A.solve(b, nchunks=1000).cpu().numpy()
    assert np.linalg.norm(x2 - ox) / np.linalg.norm(ox) <= 1e-12
    # later solves of the same operands at the default chunking run stage 1 with one lane per chunk (the form bench.py times)
    for _ in range(2):
        x3 = A.solve(b)
    assert A.residual(x3, b) <= 1e-13
    assert np.linalg.norm(x3.cpu().numpy() - ox) / np.linalg.norm(ox) <= 1e-12


def test_lpc_interleaved_copy_follows_the_operands(chunked, ctx):
    """The one-lane-per-chunk stage 1 sweeps a chunk-interleaved copy of the operands that the library builds on the first solve and
    keeps: new operands in the same (recycled) object, another chunk count, and the lanes = columns kernels must all give the
    oracle's solution."""
    import gc
    n, l, u, r = 300_000, 4, 4, 2
    default, default_lpc = ctx.get_option("abc_tiled_min_chunks"), ctx.get_option("abc_lpc")
    ctx.set_option("abc_tiled_min_chunks", 1)
    try:
        for seed in (11, 12):                                   # the second object recycles the first one's storage (context cache)
            bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=seed, dist=1)
            ox = oracle.ab_solve(bands, U, V, b, l, u)[0]
            A = chunked.LargeAlmostBandedMatrix(bands[0], (U[0], V[0]), (l, u), ctx=ctx)
            # abc_lpc: 1 = lanes = columns on the first solve of new operands, one lane per chunk from the second on; 2 = always; 0 = never
            for lpc, nchunks in ((1, 2048), (1, 2048), (1, 2048), (2, 640), (0, 2048), (2, 2048)):
                ctx.set_option("abc_lpc", lpc)
                x = A.solve(b[0], nchunks=nchunks)
                assert np.linalg.norm(x - ox) <= 1e-12 * np.linalg.norm(ox), (seed, lpc, nchunks)
                assert A.residual(x, b[0]) <= 1e-13
            del A
            gc.collect()
    finally:
        ctx.set_option("abc_tiled_min_chunks", default)
        ctx.set_option("abc_lpc", default_lpc)


def test_lpc_drawn_systems_match_oracle(ctx):
    """The one-lane-per-chunk stage 1 alone (abc_lpc = 2, tiled records forced): splits of l + u = 8, 6 and 4 at r = 2, sizes from a few rows per
    chunk to tens of thousands of rows, chunk counts that leave partial tiles of 32 and both kinds of chunk length, both generators —
    the solution against the oracle's sequential sweep and the structured residual; and against the lanes = columns kernels."""
    from ssm_b200 import chunked
    rng = np.random.default_rng(4242)
    default, default_lpc = ctx.get_option("abc_tiled_min_chunks"), ctx.get_option("abc_lpc")
    ctx.set_option("abc_tiled_min_chunks", 1)
    try:
        for it in range(24):
            lp = int(rng.choice([8, 8, 6, 4])); l = int(rng.integers(0, lp + 1)); u = lp - l; r = 2
            n = int(rng.choice([40, 333, 2000, 4097, 33000, 70001]))
            dist = int(rng.integers(0, 2))
            bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=7000 + it, dist=dist)
            ox = oracle.ab_solve(bands[0], U[0], V[0], b[0], l, u)
            A = chunked.LargeAlmostBandedMatrix(bands[0], (U[0], V[0]), (l, u), ctx=ctx)
            for nchunks in sorted({1, 33, int(rng.integers(1, max(2, n // 27 + 1))), int(rng.integers(1, max(2, n // 200 + 1)))}):
                ctx.set_option("abc_lpc", 2)
                x = A.solve(b[0], nchunks=nchunks)
                tag = (it, n, l, u, dist, nchunks, A.chunks)
                assert np.linalg.norm(x - ox) <= 1e-12 * np.linalg.norm(ox), tag
                assert A.residual(x, b[0]) <= 1e-13, tag
                ctx.set_option("abc_lpc", 0)
                x0 = A.solve(b[0], nchunks=nchunks)
                assert np.linalg.norm(x - x0) <= 1e-13 * np.linalg.norm(ox), tag
    finally:
        ctx.set_option("abc_tiled_min_chunks", default)
        ctx.set_option("abc_lpc", default_lpc)


def test_random_shapes_and_chunk_counts_match_oracle(ctx):
    """Sweep of 40 drawn single systems over every compiled (l+u, r) of the chunked path, sizes from 1 row to a few thousand and chunk
    counts from 1 to n/8, both stage-3 variants: the solution against the oracle's sequential sweep and the structured residual."""
    from ssm_b200 import chunked
    shapes = [(1, 0), (2, 0), (4, 0), (1, 1), (1, 2), (2, 1), (2, 2), (3, 1), (3, 2), (4, 2), (6, 2), (8, 2)]
    rng = np.random.default_rng(2028)
    default, default_gs, default_lpc = ctx.get_option("abc_tiled_min_chunks"), ctx.get_option("abc_gs"), ctx.get_option("abc_lpc")
    try:
        for it in range(40):
            ctx.set_option("abc_gs", 4 if it % 2 == 0 else 8)      # lanes per chunk in stage 1 of the wide shapes
            ctx.set_option("abc_lpc", 2 if it % 4 < 2 else 0)      # one lane per chunk for (8, 2) over tiled records
            lp, r = shapes[int(rng.integers(0, len(shapes)))]
            l = int(rng.integers(0, lp + 1)); u = lp - l
            n = int(rng.choice([1, 2, lp + 1, 2 * lp + 3, 100, 1000, 4097, 20000]))
            dist = int(rng.integers(0, 2)) if r > 0 and n > r else 0
            bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=3000 + it, dist=dist)
            A = chunked.LargeAlmostBandedMatrix(bands[0], (U[0], V[0]), (l, u), ctx=ctx)
            ox = oracle.ab_solve(bands[0], U[0], V[0], b[0], l, u)
            for min_chunks in (default, 1):
                ctx.set_option("abc_tiled_min_chunks", min_chunks)
                for nchunks in sorted({0, 1, int(rng.integers(1, max(2, n // 8 + 1)))}):
                    x = A.solve(b[0], nchunks=nchunks)
                    tag = (it, n, l, u, r, dist, min_chunks, nchunks)
                    assert np.linalg.norm(x - ox) <= 1e-12 * np.linalg.norm(ox), tag
                    assert A.residual(x, b[0]) <= 1e-13, tag
    finally:
        ctx.set_option("abc_tiled_min_chunks", default)
        ctx.set_option("abc_gs", default_gs)
        ctx.set_option("abc_lpc", default_lpc)
