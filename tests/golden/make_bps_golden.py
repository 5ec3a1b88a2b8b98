#!/usr/bin/env python
"""Generates tests/golden/bps_qr_mp_n20.json and tests/golden/bps_qr_ld_n256.npz: independent high-precision judges of the
BandedPlusSemiseparable QR (semiseparableqr.jl:104-152) for the oracle AND the GPU factors.

* n = 20, (l, m, r, p) = (4, 5, 2, 3) — the shape of the reference's own test (test/test_qr.jl:6-12): dense Householder QR in 50-digit mpmath
  arithmetic (LAPACK geqr2 / LinearAlgebra.reflector! convention), rounded to double.
* n = 256, (3, 3, 2, 2), randn data: the same algorithm in numpy longdouble (64-bit mantissa) — three digits beyond double.
Inputs are stored with the results, so the tests depend on no random generator.  Run from the repo root:  python tests/golden/make_bps_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import oracle  # noqa: E402
from helpers import bps_randn  # noqa: E402

OUT = Path(__file__).resolve().parent


def householder_qr(A, zero, one, sqrt, absf):
    """unblocked Householder QR in the arithmetic of A's entries: returns (factors, tau) in LAPACK's storage convention, tau[n-1] = 0"""
    n = len(A)
    A = [row[:] for row in A]
    tau = [zero] * n
    for k in range(n - 1):
        ssq = zero
        for i in range(k, n):
            ssq += A[i][k] * A[i][k]
        if ssq == zero:
            continue
        nrm = sqrt(ssq)
        x0 = A[k][k]
        nu = nrm if x0 >= zero else -nrm          # copysign(|x|, x0)
        xi = x0 + nu
        v = [one] + [A[i][k] / xi for i in range(k + 1, n)]
        t = xi / nu
        tau[k] = t
        for j in range(k + 1, n):
            s = zero
            for i in range(k, n):
                s += v[i - k] * A[i][j]
            s *= t
            for i in range(k, n):
                A[i][j] -= v[i - k] * s
        A[k][k] = -nu
        for i in range(k + 1, n):
            A[i][k] = v[i - k]
    return A, tau


def mp_case():
    import mpmath as mp
    mp.mp.dps = 50
    n, l, m, r, p = 20, 4, 5, 2, 3
    rng = np.random.default_rng(20261021)
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p)
    A = oracle.bps_dense(B, U, V, W, S, l, m)
    # the dense matrix is formed from the double inputs in mp arithmetic as well (products of generators are not exact in double)
    Am = [[mp.mpf(0)] * n for _ in range(n)]
    for k in range(n):
        for j in range(n):
            val = mp.mpf(0)
            if -l <= j - k <= m:
                val += mp.mpf(float(B[j, m + k - j]))
            if k > j:
                for q in range(r):
                    val += mp.mpf(float(U[q, k])) * mp.mpf(float(V[q, j]))
            if k < j:
                for q in range(p):
                    val += mp.mpf(float(W[q, k])) * mp.mpf(float(S[q, j]))
            Am[k][j] = val
    assert max(abs(float(Am[k][j]) - A[k, j]) for k in range(n) for j in range(n)) < 1e-13
    F, tau = householder_qr(Am, mp.mpf(0), mp.mpf(1), mp.sqrt, abs)
    json.dump({"n": n, "l": l, "m": m, "r": r, "p": p, "B": B.tolist(), "U": U.tolist(), "V": V.tolist(), "W": W.tolist(), "S": S.tolist(),
               "b": b.tolist(), "factors": [[float(x) for x in row] for row in F], "tau": [float(t) for t in tau],
               "how": "dense Householder QR, mpmath 50 digits, rounded to double (tests/golden/make_bps_golden.py)"},
              open(OUT / "bps_qr_mp_n20.json", "w"))


def ld_case():
    n, l, m, r, p = 256, 3, 3, 2, 2
    rng = np.random.default_rng(20261022)
    B, U, V, W, S, b = bps_randn(rng, n, l, m, r, p)
    ld = np.longdouble
    A = np.zeros((n, n), dtype=ld)
    for k in range(n):
        for j in range(max(0, k - l), min(n, k + m + 1)):
            A[k, j] += ld(B[j, m + k - j])
    A += np.tril(U.T.astype(ld) @ V.astype(ld), -1) + np.triu(W.T.astype(ld) @ S.astype(ld), 1)
    tau = np.zeros(n, dtype=ld)
    for k in range(n - 1):
        x = A[k:, k].copy()
        nrm = np.sqrt(np.sum(x * x))
        nu = nrm if x[0] >= 0 else -nrm
        xi = x[0] + nu
        v = x / xi; v[0] = 1
        tau[k] = xi / nu
        A[k:, k + 1:] -= np.outer(v, tau[k] * (v @ A[k:, k + 1:]))
        A[k, k] = -nu
        A[k + 1:, k] = v[1:]
    np.savez_compressed(OUT / "bps_qr_ld_n256.npz", shape=np.array([n, l, m, r, p]), B=B, U=U, V=V, W=W, S=S, b=b,
                        factors=A.astype(np.float64), tau=tau.astype(np.float64))


if __name__ == "__main__":
    mp_case()
    ld_case()
    print("written", sorted(p.name for p in OUT.glob("bps_qr_*")))
