"""CPU tier: the algebra of the chunked n-axis solve (K10) — tests/emul/sof_reference.py, the numpy restatement the CUDA
kernels in csrc/abc_kernels.cu follow — against dense LAPACK and against the reference-restating oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "emul"))
import oracle  # noqa: E402
from helpers import readme_expz  # noqa: E402
from sof_reference import dense, solve_chunked  # noqa: E402


@pytest.mark.parametrize("n,l,u,r,P", [(64, 2, 2, 2, 4), (50, 1, 0, 1, 3), (97, 4, 4, 2, 4), (40, 2, 0, 2, 2), (33, 1, 1, 1, 1),
                                       (120, 3, 1, 2, 8), (64, 0, 2, 1, 4), (64, 2, 2, 0, 4)])
def test_chunked_algebra_matches_dense(n, l, u, r, P):
    rng = np.random.default_rng(n + 7 * P)
    bands = rng.standard_normal((n, l + u + 1)); U = rng.standard_normal((r, n)); V = rng.standard_normal((n, r))
    bands[:, u] += 4 * np.sign(bands[:, u])
    b = rng.standard_normal(n)
    A = dense(bands, U, V, l, u)
    x0 = np.linalg.solve(A, b)
    for regular in (False, True):
        x = solve_chunked(bands, U, V, b, l, u, P, regular=regular)
        assert np.linalg.norm(x - x0) / np.linalg.norm(x0) <= 1e-12 * max(1.0, np.linalg.cond(A) / 100)


def test_chunked_algebra_readme_known_answer():
    (l, u, r), bands, U, V, b = readme_expz(24)
    import math
    for P in (1, 2, 4):
        x = solve_chunked(bands, U, V, b, l, u, P)
        ref = np.array([1 / math.factorial(k) for k in range(16)])
        # normwise: the partitioned factorisation is backward stable, not componentwise accurate on the graded tail
        assert np.linalg.norm(x[:16] - ref) / np.linalg.norm(ref) <= 1e-14


def test_chunked_algebra_matches_oracle_on_generated_inputs():
    n, l, u, r = 300, 4, 4, 2
    for dist in (0, 1):
        bands, U, V, b = oracle.ab_generate(n, l, u, r, 1, seed=20261022, dist=dist)
        ox = oracle.ab_solve(bands, U, V, b, l, u)[0]
        x = solve_chunked(bands[0], U[0], V[0], b[0], l, u, 5)
        assert np.linalg.norm(x - ox) / np.linalg.norm(ox) <= 1e-13


def test_regular_layout_invariants():
    """The chunk layout the CUDA path uses (abc_layout in csrc/abc_kernels.cu, mirrored by regular_layout): every chunk eliminates a
    whole number of unroll periods, groups of `same` consecutive chunks share a length (4 = one stage-1 warp, 32 = one stage-3
    tile of the many-chunk path), the chunks tile the padded system, and the padding stays inside what the library allocates
    behind the operands (ABC_PAD = 512 rows minus the rows the sweeps read ahead)."""
    from sof_reference import regular_layout
    rng = np.random.default_rng(5)
    cases = [(n, lp, P) for n in (1, 7, 33, 500, 4097, 100_000, 1 << 20, (1 << 24) + 3) for lp in (1, 2, 4, 8) for P in (1, 3, 64, 1000, 32768)]
    cases += [(int(rng.integers(1, 3_000_000)), int(rng.choice([1, 2, 3, 4, 6, 8])), int(rng.integers(1, 40_000))) for _ in range(300)]
    for n, lp, P0 in cases:
        for same in (4, 32):
            P, bnd = regular_layout(n, lp, P0, same=same)
            pw = lp + 1
            lens = np.diff(bnd)
            assert 1 <= P <= max(P0, 1) and len(bnd) == P + 1 and bnd[0] == 0
            assert np.all(lens % pw == lp) and np.all(lens >= 2 * pw - 1), (n, lp, P0, same)
            assert bnd[-1] >= n, (n, lp, P0, same)                                   # the chunks cover the system ...
            assert bnd[-1] - n <= 512 - (2 * pw + lp + 8), (n, lp, P0, same)         # ... and the padding fits (u <= lp)
            for g in range(0, P, same):                                              # groups share a length
                assert len(set(lens[g:g + same])) == 1, (n, lp, P0, same, g)
            assert np.all(lens[:-1] >= lens[1:])                                     # long chunks first
