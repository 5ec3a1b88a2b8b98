"""CPU oracle for the symmetric BandedPlusSemiseparable R*Q step (the QR-algorithm building block) — TEST INFRASTRUCTURE ONLY.

numpy restatement, term by term, of the reference's `rq_mul` (semiseparableqr.jl:170-311 driver and state, :642-845 one step
and lookup tables): given the QR factors of a SYMMETRIC BandedPlusSemiseparableMatrix (what `qr(A)` returns,
semiseparableqr.jl:128-131) it produces R*Q = Q'AQ as a symmetric BandedPlusSemiseparableMatrix
`BPS(B_sym, (U_new, V_new), (V_new, U_new))` (:282-287).  Pure-Python loops: small cases only.

Pinned the way the reference's own test pins it (test/test_rq.jl:26-34): dense `R0 * Q` from LAPACK-convention Householder QR.

Layout (one system, reference memory): Bf (n, 2l+m+1) band storage with bandwidths (l, mw = l+m), A[i,j] at Bf[j, mw+i-j];
U0 (r, n); Vf (r, n); Wf, Sf (p+r, n); tau (n).  Indices below are 0-based; q is the column being processed.
"""
from __future__ import annotations

import numpy as np


def _bget(Bf, l, mw, i, k):
    n = Bf.shape[0]
    if i < 0 or k < 0 or i >= n or k >= n or k - i > mw or i - k > l:
        return 0.0
    return Bf[k, mw + i - k]


def rq_mul(Bf, U0, Vf, Wf, Sf, tau, l, mw):
    """Returns (Bsym (n, 2l+1) band storage with bandwidths (l, l), U_new (r, n), V_new (r, n))."""
    n = Bf.shape[0]
    r = U0.shape[0]
    pw = Wf.shape[0]
    W = Wf.T.copy(); S = Sf.T.copy(); V = Vf.T.copy()                 # n x pw, n x pw, n x r
    Su = S[:, :r]                                                     # = U0 (constructor check :199-201)
    assert pw > r and np.array_equal(Su, U0.T), "S[:, 1:r] must equal U (semiseparableqr.jl:195-201)"
    assert mw >= l, "upper bandwidth of B must be no less than the lower one (:203-205)"
    B = lambda i, k: _bget(Bf, l, mw, i, k)                           # the ORIGINAL factor entries (R0 and reflector tails)

    # fast_RU (:818-835): U_new = R0 U0, R0 = triu(B) + triu(W S', 1);  r1U2 table (:800-816) = R0[q, q+1:] U0[q+1:, :]
    StU = S.T @ Su                                                    # pw x r, suffix over rows > i after the subtraction below
    Unew = np.zeros((n, r)); r1U2 = np.zeros((n, r))
    for i in range(n):
        StU = StU - np.outer(S[i], Su[i])
        off = sum(B(i, k) * Su[k] for k in range(i + 1, min(i + mw, n - 1) + 1))
        r1U2[i] = W[i] @ StU + off
        Unew[i] = r1U2[i] + B(i, i) * Su[i]
    # UtU table (:786-798): suffix Gram of U0 including row i
    UtU = np.zeros((n + 1, r, r))
    for i in range(n - 1, -1, -1):
        UtU[i] = UtU[i + 1] + np.outer(Su[i], Su[i])

    lphi = min(l, n)
    Om = np.zeros((r, r)); Phi = np.zeros((lphi, r)); Psi = np.zeros((r, lphi)); Lam = np.zeros((lphi, lphi))
    Bdiag = np.zeros(n); Blow = np.zeros((n, l))                      # new diagonal, new sub-diagonal column tails
    Vnew = np.zeros((n, r))
    for q in range(n - 1):                                            # onestep_rq! (:296-311)
        lb = min(l, n - 1 - q)
        b = np.array([B(q + 1 + t, q) for t in range(lb)])
        kb = V[q].copy()
        tq = tau[q]
        # compute_gamma (:642-653): gamma = R0[q:q+lb, q+1:q+lb] b
        gam = np.zeros(lb + 1)
        for i in range(1, lb + 1):
            for t in range(i + 1):
                gam[t] += b[i - 1] * B(q + t, q + i)
            for t in range(i):
                gam[t] += b[i - 1] * (W[q + t] @ S[q + i])
        # compute_variables_delta (:655-685)
        d2 = Su[q] + UtU[q + 1] @ kb + (Su[q + 1:q + 1 + lb].T @ b if lb else 0.0)
        d1 = Om @ d2
        nv = min(l, n - q)
        yh = (Su[q + 1:q + nv] @ kb + b[:nv - 1]) if nv > 1 else np.zeros(0)
        d3 = Psi[:, 1:nv] @ yh + (Psi[:, 0] if l > 0 else 0.0)
        d4 = Lam[:nv, 1:nv] @ yh + (Lam[:nv, 0] if l > 0 else 0.0)
        u1 = Su[q]
        # update_lower_triangular_part_rq! (:687-717)
        eta = -tq * gam[1:]
        k1 = min(len(eta), lphi - 1)
        eta[:k1] += Phi[1:1 + k1] @ u1 - tq * (Phi[1:1 + k1] @ d2)
        eta[:len(d4) - 1] += -tq * d4[1:]
        if l > 0:
            eta[:k1] += Lam[1:1 + k1, 0]
        mu = -tq * kb + Om @ u1 - tq * d1 - tq * d3 + (Psi[:, 0] if l > 0 else 0.0)
        Blow[q, :lb] = eta
        Vnew[q] = mu
        # update_diagonal_rq! (:719-737)
        d0 = B(q, q); rho = r1U2[q]
        dg = (d0 - tq * d0 - tq * (rho @ kb) + d0 * (u1 @ Om @ u1 - tq * (u1 @ d1)) + rho @ (Om @ u1 - tq * d1)
              - tq * d0 * (u1 @ d3) - tq * (rho @ d3))
        if l > 0:
            dg += (-tq * gam[0] + Phi[0] @ u1 + d0 * (u1 @ Psi[:, 0]) - tq * (Phi[0] @ d2) + rho @ Psi[:, 0] - tq * d4[0] + Lam[0, 0])
        Bdiag[q] = dg
        # update_next_submatrix_rq! (:739-784)
        kk = kb + d1 + d3
        Om = Om - tq * np.outer(kk, kb)
        Psi_old = Psi.copy()
        Psi = np.zeros_like(Psi)
        Psi[:, :lb] = -tq * np.outer(kk, b)
        Psi[:, :lphi - 1] += Psi_old[:, 1:]
        Psi[:, lb:] = 0.0
        Lam_old = Lam.copy(); Phi_old = Phi.copy()
        Lam = np.zeros_like(Lam)
        Lam[:lb, :lb] = -tq * np.outer(gam[1:], b)
        Lam[:lphi - 1, :lb] += -tq * np.outer(Phi_old[1:] @ d2, b)
        Lam[:lphi - 1, :lphi - 1] += Lam_old[1:, 1:]
        Lam[:len(d4) - 1, :lb] += -tq * np.outer(d4[1:], b)
        Lam[lb:, :] = 0.0
        Lam[:, lb:] = 0.0
        tmp = Phi_old[1:] - tq * np.outer(Phi_old[1:] @ d2, kb)
        Phi = Phi_old.copy()                                          # rows the reference does not overwrite keep their values (:775-783)
        Phi[:lb] = -tq * np.outer(gam[1:], kb)
        Phi[:lphi - 1] += tmp
        Phi[:len(d4) - 1] += -tq * np.outer(d4[1:], kb)
        Phi[lb:, 0] = 0.0
    # last diagonal entry (:278, getindex :214-269 with k = i = n, j = n-1): RU row of the last row is B[n,n] U0[n,:]
    q = n - 1
    ru = B(q, q) * Su[q]
    Bdiag[q] = B(q, q) + ru @ Om @ Su[q] + (Phi[0] @ Su[q] + ru @ Psi[:, 0] + Lam[0, 0] if l > 0 else 0.0)
    # symmetrize_from_lower! (:837-845) -> band storage with bandwidths (l, l)
    Bs = np.zeros((n, 2 * l + 1))
    for j in range(n):
        Bs[j, l] = Bdiag[j]
        for t in range(1, l + 1):
            if j + t < n:
                Bs[j, l + t] = Blow[j, t - 1]                         # A[j+t, j]
                Bs[j + t, l - t] = Blow[j, t - 1]                     # A[j, j+t]
    return Bs, Unew.T.copy(), Vnew.T.copy()


def rq_dense(Bs, Unew, Vnew, l):
    """Matrix(BandedPlusSemiseparableMatrix(B, (U, V), (V, U))) for the result above."""
    n = Bs.shape[0]
    A = np.zeros((n, n))
    for j in range(n):
        for d in range(2 * l + 1):
            i = j + d - l
            if 0 <= i < n:
                A[i, j] = Bs[j, d]
    low = np.tril(Unew.T @ Vnew, -1)
    return A + low + low.T
