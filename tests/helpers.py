"""Shared test helpers: re-instantiations of the reference's test matrices in the reference memory layout."""
import math

import numpy as np

import oracle


def zero_corners(bands, l, u):
    n = bands.shape[-2]
    for j in range(n):
        for d in range(l + u + 1):
            k = j + d - u
            if k < 0 or k >= n:
                bands[..., j, d] = 0.0
    return bands


def randn_system(rng, n, l, u, r, nb=None):
    shp = (n,) if nb is None else (nb, n)
    lead = () if nb is None else (nb,)
    bands = zero_corners(rng.standard_normal(lead + (n, l + u + 1)), l, u)
    U = rng.standard_normal(lead + (r, n)); V = rng.standard_normal(lead + (n, r)); b = rng.standard_normal(shp)
    return bands, U, V, b


def readme_expz(n=20):
    """README.md:50 — A = AlmostBandedMatrix(BandedMatrix(0 => [1; 1:n-1], -1 => Fill(-1.0,n-1)),
    ApplyMatrix(*, [1; zeros(n-1)], Fill(1.0,1,n))), b = [e; zeros(n-1)]"""
    l, u, r = 1, 0, 1
    bands = np.zeros((n, 2))
    bands[0, 0] = 1.0
    bands[1:, 0] = np.arange(1, n)
    bands[: n - 1, 1] = -1.0
    U = np.zeros((1, n)); U[0, 0] = 1.0
    V = np.ones((n, 1))
    b = np.zeros(n); b[0] = math.e
    return (l, u, r), bands, U, V, b


def kat_n80(n=80):
    """test_almostbanded.jl:135-137 — BandedMatrix(fill(2.0,n,n),(1,1)) with diag += 1:n; fill fill(3.0,n) * ones(1,n)"""
    l, u, r = 1, 1, 1
    A = np.full((n, n), 2.0) + np.diag(np.arange(1, n + 1, dtype=float))
    bands = oracle.banded_to_storage(A, l, u)
    U = np.full((1, n), 3.0); V = np.ones((n, 1))
    return (l, u, r), bands, U, V


def vcat_system(rng, n=80):
    """test_almostbanded.jl:160 — Vcat(randn(2,n), BandedMatrix((0=>-Ones(n-1), 1=>1:(n-1), 2=>Ones(n-2)), (n-2,n)))
    copied the way _copyto! does (AlmostBandedMatrix.jl:316-337): almostbandwidths (2,0), rank 2, U=[I_2;0], V=a."""
    a = rng.standard_normal((2, n))
    Bd = np.zeros((n - 2, n))
    for k in range(n - 2):
        Bd[k, k] = -1.0
        if k + 1 < n: Bd[k, k + 1] = k + 1.0
        if k + 2 < n: Bd[k, k + 2] = 1.0
    A = np.vstack([a, Bd])
    l, u, r = 2, 0, 2
    bands = oracle.banded_to_storage(A, l, u)
    U = np.zeros((r, n)); U[0, 0] = 1.0; U[1, 1] = 1.0
    V = np.ascontiguousarray(a.T)
    return (l, u, r), bands, U, V, A


def degenerate_system(rng):
    """test_almostbanded.jl:180 — Vcat(randn(2,7), BandedMatrix(2 => 1:5)[1:5,:]): almostbandwidths (1,0), rank 2."""
    n = 7
    a = rng.standard_normal((2, n))
    Bd = np.zeros((5, n))
    for k in range(5):
        Bd[k, k + 2] = k + 1.0
    A = np.vstack([a, Bd])
    l, u, r = 1, 0, 2
    bands = oracle.banded_to_storage(A, l, u)
    U = np.zeros((r, n)); U[0, 0] = 1.0; U[1, 1] = 1.0
    V = np.ascontiguousarray(a.T)
    return (l, u, r), bands, U, V, A


def R_from_factors(F, Uf, V, l, ub):
    """Dense upper-triangular R from AlmostBanded QR factors (band part + fill beyond the band)."""
    n = F.shape[0]
    R = np.triu(oracle.storage_to_dense(F, l, ub))
    for k in range(n):
        for j in range(k + ub + 1, n):
            R[k, j] = Uf[:, k] @ V[j, :]
    return R


def relerr(x, ref):
    return float(np.linalg.norm(np.asarray(x) - np.asarray(ref)) / np.linalg.norm(ref))


def bps_zero_corners(B, l, m):
    n = B.shape[-2]
    for j in range(n):
        for d in range(l + m + 1):
            k = j + d - m
            if k < 0 or k >= n:
                B[..., j, d] = 0.0
    return B


def bps_randn(rng, n, l, m, r, p, nb=None, dominant=0.0):
    """brandn(n,n,l,m) band + randn generators (test_qr.jl:6-12); `dominant` adds to |diag| to tame conditioning."""
    lead = () if nb is None else (nb,)
    B = bps_zero_corners(rng.standard_normal(lead + (n, l + m + 1)), l, m)
    if dominant:
        B[..., m] += dominant * np.sign(B[..., m])
    U = rng.standard_normal(lead + (r, n)); V = rng.standard_normal(lead + (r, n))
    W = rng.standard_normal(lead + (p, n)); S = rng.standard_normal(lead + (p, n))
    b = rng.standard_normal(lead + (n,))
    return B, U, V, W, S, b


def explicit_qt(qr_, tau, b):
    """test_qr.jl:40-46: apply (I - tau_i y y') for i = 1..n-1 with y = e_i + F[i+1:n, i]."""
    n = len(b)
    out = np.array(b, dtype=float)
    for i in range(n - 1):
        y = np.zeros(n); y[i] = 1.0; y[i + 1:] = qr_[i + 1:, i]
        out = out - tau[i] * y * (y @ out)
    return out
