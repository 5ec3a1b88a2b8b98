"""CPU tier: the growth fix-up behind ssm_abqr_resize (csrc/ab_generic.cuh: generic_ab_grow_fix), compiled for the host by tests/emul/.

Adaptive QR factors the first k0 columns of a cached operand, the cache grows (AlmostBandedMatrix.jl:357-394: new rows, new columns of V,
new band entries), and the factorisation goes on (`_almostbanded_qr!(A, tau, ncols)` :139-172 on the trailing view).  Here: factor k0
columns of the leading n0 x n0 block, move the factor records into storage for n1 > n0, re-derive the entries that reach into the new
columns by replaying the stored reflectors, continue to k1 columns — and compare with the one-shot partial factorisation of the n1 x n1
matrix (1e-13) and with the oracle.  A debugging harness for the device code, not a product path."""
import ctypes as C

import numpy as np
import pytest

import oracle
from helpers import randn_system
from test_emul_continue import PAD, dp, emul, pack, run, unpack  # noqa: F401  (emul is a fixture)


def leading_block(bands, U, V, n0, l, u):
    """operands of A[:n0, :n0] in the reference layout"""
    b0 = bands[:n0].copy()
    for j in range(n0):
        for d in range(l + u + 1):
            k = j + d - u
            if k >= n0:
                b0[j, d] = 0.0
    return b0, U[:, :n0].copy(), V[:n0].copy()


def grow(emul, n0, n1, l, u, r, RU0, QT0, AU1, V1, k0):
    nfr, nfq = l + u + 1 + r, l + 1
    RU = np.zeros((n1 + PAD, nfr, 32)); QT = np.zeros((n1 + PAD, nfq, 32))
    RU[:n0] = RU0[:n0]; QT[:n0] = QT0[:n0]                      # ssm_abqr_resize: the records of the old rows move, everything new is zero
    emul.generic_grow_fix(n1, l, u, r, C.c_long(n1 + PAD), AU1.ctypes.data_as(dp), V1.ctypes.data_as(dp), RU.ctypes.data_as(dp), QT.ctypes.data_as(dp), n0, k0)
    run(emul, n1, l, u, r, AU1, V1, RU, QT, k0, k0)             # pass the rows from k0 on through (no new reflectors): F == qr_cols(A1, k0)
    return RU, QT


@pytest.mark.parametrize("l,u,r", [(2, 2, 2), (1, 0, 1), (4, 4, 2), (3, 1, 3), (2, 0, 2), (1, 3, 1), (8, 8, 4), (2, 2, 0), (0, 2, 1)])
def test_growth_then_continuation_equals_one_shot(emul, l, u, r):
    rng = np.random.default_rng(100 * l + 10 * u + r)
    n1 = 6 * (l + u) + 25
    bands, U, V, _ = randn_system(rng, n1, l, u, r)
    bands[:, u] += 4.0                                          # keep the partial factors well scaled
    AU1, V1 = pack(bands, U, V, l, u, r)
    nfr, nfq = l + u + 1 + r, l + 1
    for n0 in sorted({2 * l + u + 3, 3 * (l + u) + 7, n1 - 1, n1 - l - u - 2}):
        if not l + 1 <= n0 < n1:
            continue
        b0, U0, V0 = leading_block(bands, U, V, n0, l, u)
        AU0, Vd0 = pack(b0, U0, V0, l, u, r)
        for k0 in sorted({0, 1, max(0, n0 - 2 * l - u - 1), max(0, n0 - 2 * l - u), max(0, n0 - l - u), n0 - l - 1, n0 - l}):
            if not 0 <= k0 <= n0 - l:
                continue
            RU0 = np.zeros((n0 + PAD, nfr, 32)); QT0 = np.zeros((n0 + PAD, nfq, 32))
            run(emul, n0, l, u, r, AU0, Vd0, RU0, QT0, k0, 0)
            RU, QT = grow(emul, n0, n1, l, u, r, RU0, QT0, AU1, V1, k0)
            RUa = np.zeros_like(RU); QTa = np.zeros_like(QT)
            run(emul, n1, l, u, r, AU1, V1, RUa, QTa, k0, 0)     # one shot on the grown operand, the same k0 reflectors
            scale = max(1.0, np.abs(RUa).max())
            assert np.abs(RU - RUa).max() <= 1e-13 * scale and np.abs(QT - QTa).max() <= 1e-13 * scale, (n0, k0)
            for k1 in sorted({min(k0 + 3, n1 - 1), n1 - 1}):
                if k1 < k0:
                    continue
                RU2, QT2 = RU.copy(), QT.copy()
                run(emul, n1, l, u, r, AU1, V1, RU2, QT2, k1, k0)
                F, Uo, tau = unpack(RU2, QT2, n1, l, u, r)
                oF, oU, otau = oracle.ab_qr(bands, U, V, l, u, ncols=k1)
                assert np.abs(F - oF).max() <= 1e-12 * max(1.0, np.abs(oF).max()), (n0, k0, k1)
                assert r == 0 or np.abs(Uo - oU).max() <= 1e-12 * max(1.0, np.abs(oU).max())
                assert np.abs(tau - otau).max() <= 1e-13 and not tau[k1:].any()


def test_growth_in_three_steps(emul):
    """n0 -> n1 -> n2 -> n3 with a few more columns factored at each size == one-shot qr at n3 (and the oracle)."""
    l, u, r = 2, 2, 2
    rng = np.random.default_rng(5)
    sizes, cols = [20, 31, 47, 64], [12, 25, 40, 63]
    n3 = sizes[-1]
    bands, U, V, _ = randn_system(rng, n3, l, u, r)
    bands[:, u] += 4.0
    nfr, nfq = l + u + 1 + r, l + 1
    RU = QT = None
    kprev = 0
    for n, k in zip(sizes, cols):
        b0, U0, V0 = leading_block(bands, U, V, n, l, u)
        AU, Vd = pack(b0, U0, V0, l, u, r)
        if RU is None:
            RU = np.zeros((n + PAD, nfr, 32)); QT = np.zeros((n + PAD, nfq, 32))
            run(emul, n, l, u, r, AU, Vd, RU, QT, k, 0)
        else:
            RU, QT = grow(emul, nprev, n, l, u, r, RU, QT, AU, Vd, kprev)
            run(emul, n, l, u, r, AU, Vd, RU, QT, k, kprev)
        nprev, kprev = n, k
    F, Uo, tau = unpack(RU, QT, n3, l, u, r)
    oF, oU, otau = oracle.ab_qr(bands, U, V, l, u, ncols=cols[-1])
    assert np.abs(F - oF).max() <= 1e-13 * max(1.0, np.abs(oF).max())
    assert np.abs(Uo - oU).max() <= 1e-13 * max(1.0, np.abs(oU).max()) and np.abs(tau - otau).max() <= 1e-13
