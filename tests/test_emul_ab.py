"""CPU tier: the device arithmetic (csrc/ab_core.cuh) compiled for the host by tests/emul/ and run against the
oracle.  This checks the register-window bookkeeping the CUDA kernels instantiate; it is a debugging harness,
not a product path (the product has no CPU fallback)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from helpers import randn_system, relerr

EMUL = Path(__file__).parent / "emul"
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    so = EMUL / "libab_emul.so"
    srcs = [EMUL / "ab_emul.cpp", EMUL.parent.parent / "semiseparablematrices.jl_b200" / "csrc" / "ab_core.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-o", str(so), str(srcs[0])], check=True)
    return C.CDLL(str(so))


SHAPES = [(20, 1, 0, 1), (80, 1, 1, 1), (64, 2, 2, 2), (257, 2, 2, 2), (50, 4, 4, 2), (33, 2, 0, 2), (7, 1, 0, 2),
          (40, 3, 3, 2), (5, 4, 4, 2), (3, 2, 2, 2), (2, 2, 2, 2), (1, 2, 2, 2), (30, 0, 2, 1), (41, 3, 1, 3)]


@pytest.mark.parametrize("n,l,u,r", SHAPES)
@pytest.mark.parametrize("dist", ["dd", "randn"])
def test_device_core_matches_oracle(emul, n, l, u, r, dist):
    if dist == "dd":
        bands, U, V, b = (a[0] for a in oracle.ab_generate(n, l, u, r, 1, seed=7 + n))
    else:
        bands, U, V, b = randn_system(np.random.default_rng(n), n, l, u, r)
    F = np.zeros((n, 2 * l + u + 1)); Uo = np.zeros((r, n)); tau = np.zeros(n); c = np.zeros(n); x = np.zeros(n)
    P = lambda a: a.ctypes.data_as(dp)
    rc = emul.emul_ab(n, l, u, r, P(bands), P(U), P(V), P(b), P(F), P(Uo), P(tau), P(c), P(x))
    assert rc == 0
    F0, U0, t0 = oracle.ab_qr(bands, U, V, l, u)
    tol = 1e-14 if dist == "dd" else 1e-11
    assert np.abs(F - F0).max() <= tol * np.abs(F0).max()
    assert np.abs(Uo - U0).max() <= tol * max(1.0, np.abs(U0).max())
    assert np.abs(tau - t0).max() <= tol
    assert np.abs(c - oracle.ab_lmul_qt(F0, t0, b, l, l + u)).max() <= tol * max(1.0, np.abs(b).max())
    assert relerr(x, oracle.ab_solve(bands, U, V, b, l, u)) <= (1e-14 if dist == "dd" else 1e-9)
