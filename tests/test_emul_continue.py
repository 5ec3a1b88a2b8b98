"""CPU tier: the runtime-shape device routine (csrc/ab_generic.cuh) compiled for the host by tests/emul/, checking the second entry of the
sweep that ssm_abqr_continue uses: a partial factorisation continued from its own factor records equals the one-shot partial
factorisation bit for bit, and both match the oracle's _almostbanded_qr!(A, tau, ncols) (AlmostBandedMatrix.jl:139-172).  A debugging
harness for the device code, not a product path (the product has no CPU fallback)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from helpers import randn_system

EMUL = Path(__file__).parent / "emul"
dp = C.POINTER(C.c_double)
PAD = 48


@pytest.fixture(scope="module")
def emul():
    so = EMUL / "libgeneric_emul.so"
    csrc = EMUL.parent.parent / "semiseparablematrices.jl_b200" / "csrc"
    srcs = [EMUL / "generic_emul.cpp", csrc / "ab_generic.cuh", csrc / "ssm_internal.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-attributes",
                        "-I/usr/local/cuda/include", "-o", str(so), str(srcs[0])], check=True)
    return C.CDLL(str(so))


def pack(bands, U, V, l, u, r):
    """one system into lane 0 of one tile of the device layout (mirrors ab_pack_kernel, csrc/util_kernels.cu)"""
    n = bands.shape[0]
    nfa = l + u + 1 + r
    AU = np.zeros((n + PAD, nfa, 32)); Vd = np.zeros((n + PAD, max(r, 1), 32))
    for i in range(n):
        for d in range(l + u + 1):
            j = i - l + d
            if 0 <= j < n:
                AU[i, d, 0] = bands[j, l + u - d]
        AU[i, l + u + 1:, 0] = U[:, i]
        Vd[i, :r, 0] = V[i]
    return AU, Vd


def run(emul, n, l, u, r, AU, Vd, RU, QT, ncols, kstart):
    emul.generic_qr(n, l, u, r, C.c_long(n + PAD), AU.ctypes.data_as(dp), Vd.ctypes.data_as(dp), RU.ctypes.data_as(dp), QT.ctypes.data_as(dp), ncols, kstart)


def unpack(RU, QT, n, l, u, r):
    """factor records -> the reference's factor storage (mirrors abqr_unpack_kernel)"""
    ub, nd = l + u, 2 * l + u + 1
    F = np.zeros((n, nd)); Uo = np.zeros((r, n)); tau = np.zeros(n)
    for k in range(n):
        for t in range(ub + 1):
            if k + t < n:
                F[k + t, ub - t] = RU[k, t, 0]
        for i in range(1, l + 1):
            if k + i < n:
                F[k, ub + i] = QT[k, i - 1, 0]
        Uo[:, k] = RU[k, ub + 1:, 0]
        tau[k] = QT[k, l, 0]
    return F, Uo, tau


@pytest.mark.parametrize("n,l,u,r", [(40, 2, 2, 2), (33, 1, 0, 1), (37, 4, 4, 2), (29, 3, 1, 3), (20, 0, 2, 1), (6, 4, 4, 2), (3, 2, 2, 2), (50, 8, 8, 4),
                                      (25, 2, 2, 0)])
def test_continuation_equals_one_shot_partial_qr(emul, n, l, u, r):
    rng = np.random.default_rng(7 * n + l)
    bands, U, V, _ = randn_system(rng, n, l, u, r)
    AU, Vd = pack(bands, U, V, l, u, r)
    nfr, nfq = l + u + 1 + r, l + 1
    pairs = {(1, 2), (min(l + u + 1, n - 1), max(n // 2 + 1, min(l + u + 1, n - 1))), (n // 2, n - 1), (min(3, n), n), (1, n - 1), (n - 1, n)}
    for n0, n1 in sorted(pairs):
        if not 1 <= n0 <= n1 <= n:
            continue
        RU1 = np.zeros((n + PAD, nfr, 32)); QT1 = np.zeros((n + PAD, nfq, 32))
        run(emul, n, l, u, r, AU, Vd, RU1, QT1, n1, 0)                 # one shot: n1 reflectors
        RU2 = np.zeros_like(RU1); QT2 = np.zeros_like(QT1)
        run(emul, n, l, u, r, AU, Vd, RU2, QT2, n0, 0)                 # n0 reflectors ...
        run(emul, n, l, u, r, AU, Vd, RU2, QT2, n1, n0)                # ... continued in place from the factor records
        assert np.array_equal(RU1, RU2) and np.array_equal(QT1, QT2), (n0, n1)
        mid = n0 + (n1 - n0) // 2                                      # and in two steps
        if n0 < mid < n1:
            RU3 = np.zeros_like(RU1); QT3 = np.zeros_like(QT1)
            run(emul, n, l, u, r, AU, Vd, RU3, QT3, n0, 0)
            run(emul, n, l, u, r, AU, Vd, RU3, QT3, mid, n0)
            run(emul, n, l, u, r, AU, Vd, RU3, QT3, n1, mid)
            assert np.array_equal(RU1, RU3) and np.array_equal(QT1, QT3), (n0, mid, n1)
        F, Uo, tau = unpack(RU2, QT2, n, l, u, r)
        oF, oU, otau = oracle.ab_qr(bands, U, V, l, u, ncols=n1)
        assert np.abs(F - oF).max() <= 1e-12 * max(1.0, np.abs(oF).max()), (n0, n1)
        assert r == 0 or np.abs(Uo - oU).max() <= 1e-12 * max(1.0, np.abs(oU).max())
        assert np.abs(tau - otau).max() <= 1e-13 and not tau[n1:].any()
