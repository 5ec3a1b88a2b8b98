"""CPU tier: the streaming BPS device arithmetic (csrc/bps_core.cuh: running sums instead of the reference's O(n)
lookup tables, zero-padding of small shapes into the compiled kernel shapes) compiled for the host and checked
against the oracle.  Debugging harness only — the product has no CPU path."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from helpers import bps_randn

EMUL = Path(__file__).parent / "emul"
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    so = EMUL / "libbps_emul.so"
    srcs = [EMUL / "bps_emul.cpp", EMUL.parent.parent / "semiseparablematrices.jl_b200" / "csrc" / "bps_core.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas",
                        "-o", str(so), str(srcs[0])], check=True)
    return C.CDLL(str(so))


SHAPES = [(20, 4, 5, 2, 3), (64, 3, 3, 2, 2), (30, 3, 1, 1, 1), (30, 1, 3, 2, 2), (25, 2, 2, 3, 1), (30, 0, 3, 2, 2),
          (30, 0, 0, 1, 1), (30, 3, 0, 1, 1), (6, 4, 5, 2, 3), (3, 3, 3, 2, 2), (2, 1, 1, 1, 1), (1, 1, 1, 1, 1), (40, 2, 1, 1, 2)]


@pytest.mark.parametrize("n,l,m,r,p", SHAPES + [(1024, 3, 3, 2, 2)])
@pytest.mark.parametrize("dist", ["dd", "randn"])
def test_streaming_core_matches_oracle(emul, n, l, m, r, p, dist):
    if dist == "dd":
        B, U, V, W, S, b = (a[0] for a in oracle.bps_generate(n, l, m, r, p, 1, seed=5 + n))
    else:
        B, U, V, W, S, b = bps_randn(np.random.default_rng(n + l), n, l, m, r, p, dominant=4.0)
    Bf = np.zeros((n, 2 * l + m + 1)); Vf = np.zeros((r, n)); Wf = np.zeros((p + r, n)); Sf = np.zeros((p + r, n))
    tau = np.zeros(n); c = np.zeros(n); x = np.zeros(n)
    P = lambda a: a.ctypes.data_as(dp)
    rc = emul.emul_bps(n, l, m, r, p, P(B), P(U), P(V), P(W), P(S), P(b), P(Bf), P(Vf), P(Wf), P(Sf), P(tau), P(c), P(x))
    assert rc == 0
    oB, oV, oW, oS, ot = oracle.bps_qr(B, U, V, W, S, l, m)
    oc = oracle.bps_lmul_qt(oB, U, oV, ot, b, l, l + m)
    ox = oracle.bps_upper_ldiv(oB, oW, oS, oc, l, l + m)
    tol = 2e-14 if dist == "dd" else 1e-11
    e = lambda a, r_: np.abs(a - r_).max() / max(1.0, np.abs(r_).max())
    assert e(Bf, oB) <= tol and e(Vf, oV) <= tol and e(Wf, oW) <= tol and e(Sf, oS) <= tol
    assert np.abs(tau - ot).max() <= tol
    assert e(c, oc) <= tol
    assert np.linalg.norm(x - ox) / np.linalg.norm(ox) <= (1e-13 if dist == "dd" else 1e-9)
