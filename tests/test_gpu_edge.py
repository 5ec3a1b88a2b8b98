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
