"""CPU oracle for the structured-QR hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``ssm_b200``) never does.

Two layers:

* ``libssm_oracle.so`` (``ab_oracle.c``, ``bps_oracle.c``): plain-C O(n) restatements that follow the
  reference's loop structure on the reference's memory layout (file:line citations in the C sources).
* dense helpers below (numpy / LAPACK ``dgeqrf``): what the reference's own tests compare against
  (``test/test_almostbanded.jl:145,153``, ``test/test_qr.jl:16-18``).

Parity pinning: Julia is not installed in this image, so the reference cannot be executed.  The C
restatement is pinned against (i) the README known-answer vector (``README.md:73-94``) and (ii) dense
LAPACK on re-instantiations of the reference's tests — exactly the checks the reference's suite makes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

c_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> Path:
    so = _HERE / "libssm_oracle.so"
    srcs = [_HERE / "ab_oracle.c", _HERE / "bps_oracle.c", _HERE.parent / "include" / "ssm_rng.h"]
    stale = (not so.exists()) or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs if s.exists())
    if force or stale:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = _HERE / "libssm_oracle.so"
        if not so.exists() or os.environ.get("SSM_ORACLE_REBUILD"):
            build()
        else:
            try:
                build()
            except Exception:
                pass  # prebuilt .so on a box without a toolchain
        _LIB = C.CDLL(str(so))
        _LIB.ssm_oracle_ab_relres.restype = C.c_double
        if hasattr(_LIB, "ssm_oracle_bps_relres"):
            _LIB.ssm_oracle_bps_relres.restype = C.c_double
    return _LIB


def threads(want: int = 0) -> int:
    """OpenMP team size of the batched drivers; ``want`` > 0 sets it first (overrides an inherited OMP_NUM_THREADS)."""
    return int(lib().ssm_oracle_threads(int(want)))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


def _p(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(c_dp)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------------------
# AlmostBanded — reference layout: bands (nb, n, l+u+1) [per system: column-major (l+u+1) x n],
#                U (nb, r, n) [column-major n x r], V (nb, n, r) [column-major r x n], b (nb, n)
# i.e. numpy C-order arrays whose trailing axes are the Julia arrays' memory order.
# --------------------------------------------------------------------------------------------------
def ab_generate(n, l, u, r, nb, seed, dist=0, s0=0):
    bands = np.empty((nb, n, l + u + 1)); U = np.empty((nb, r, n)); V = np.empty((nb, n, r)); b = np.empty((nb, n))
    lib().ssm_oracle_ab_generate(n, l, u, r, C.c_int64(s0), C.c_int64(nb), C.c_uint64(seed), dist,
                                 _p(bands), _p(U), _p(V), _p(b))
    return bands, U, V, b


def ab_dense(bands, U, V, l, u):
    """Dense matrix of ONE system (n x n, numpy row-major view of the column-major result)."""
    n = bands.shape[0]; r = U.shape[0]
    out = np.empty((n, n))
    lib().ssm_oracle_ab_dense(n, l, u, r, _p(_f64(bands)), _p(_f64(U)), _p(_f64(V)), _p(out))
    return out.T.copy()


def ab_qr(bands, U, V, l, u, ncols=0):
    """qr(A) of one system -> (factors (n, 2l+u+1), Uout (r, n), tau (n,))  [AlmostBandedMatrix.jl:129-179]"""
    n = bands.shape[0]; r = U.shape[0]
    F = np.zeros((n, 2 * l + u + 1)); Uo = np.zeros((r, n)); tau = np.zeros(n)
    rc = lib().ssm_oracle_ab_qr(n, l, u, r, _p(_f64(bands)), _p(_f64(U)), _p(_f64(V)), _p(F), _p(Uo), _p(tau), ncols)
    assert rc == 0
    return F, Uo, tau


def ab_ldiv(F, Uf, V, tau, b, l, ub):
    """ldiv!(F, b) of one system [AlmostBandedMatrix.jl:187-192]"""
    n = F.shape[0]; r = Uf.shape[0]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_ab_ldiv(n, l, ub, r, _p(_f64(F)), _p(_f64(Uf)), _p(_f64(V)), _p(_f64(tau)), _p(x))
    return x


def ab_lmul_qt(F, tau, b, l, ub, adjoint=True):
    n = F.shape[0]
    x = np.array(b, dtype=np.float64, copy=True)
    fn = lib().ssm_oracle_ab_lmul_qt if adjoint else lib().ssm_oracle_ab_lmul_q
    fn(n, l, ub, _p(_f64(F)), _p(_f64(tau)), _p(x))
    return x


def ab_tri_ldiv(bands, U, V, b, l, u, upper=True, unit=False):
    n = bands.shape[0]; r = U.shape[0]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_ab_tri_ldiv(n, l, u, r, _p(_f64(bands)), _p(_f64(U)), _p(_f64(V)), _p(x), int(upper), int(unit))
    return x


def ab_solve(bands, U, V, b, l, u):
    """A \\ b for one system (bands (n,l+u+1)) or a batch (bands (nb,n,l+u+1))."""
    bands = _f64(bands); U = _f64(U); V = _f64(V); b = _f64(b)
    if bands.ndim == 2:
        bands, U, V, b = bands[None], U[None], V[None], b[None]
        return ab_solve(bands, U, V, b, l, u)[0]
    nb, n, _ = bands.shape; r = U.shape[1]
    x = np.empty((nb, n))
    rc = lib().ssm_oracle_ab_solve_batch(n, l, u, r, C.c_int64(nb), _p(bands), _p(U), _p(V), _p(b), _p(x))
    assert rc == 0
    return x


def ab_qr_batch(bands, U, V, l, u):
    bands = _f64(bands); U = _f64(U); V = _f64(V)
    nb, n, _ = bands.shape; r = U.shape[1]
    F = np.zeros((nb, n, 2 * l + u + 1)); Uo = np.zeros((nb, r, n)); tau = np.zeros((nb, n))
    lib().ssm_oracle_ab_qr_batch(n, l, u, r, C.c_int64(nb), _p(bands), _p(U), _p(V), _p(F), _p(Uo), _p(tau))
    return F, Uo, tau


def ab_ldiv_batch(F, Uf, V, tau, b, l, ub):
    nb, n, _ = F.shape; r = Uf.shape[1]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_ab_ldiv_batch(n, l, ub, r, C.c_int64(nb), _p(_f64(F)), _p(_f64(Uf)), _p(_f64(V)), _p(_f64(tau)), _p(x))
    return x


def ab_relres(bands, U, V, x, b, l, u):
    n = bands.shape[0]; r = U.shape[0]
    return float(lib().ssm_oracle_ab_relres(n, l, u, r, _p(_f64(bands)), _p(_f64(U)), _p(_f64(V)), _p(_f64(x)), _p(_f64(b))))


# ---- dense helpers (what the reference's tests compare against) ---------------------------------
def banded_to_storage(Adense, l, u):
    """(n, l+u+1) array = column-major band storage of the band part of a dense matrix."""
    n = Adense.shape[0]
    out = np.zeros((n, l + u + 1))
    for j in range(n):
        for k in range(max(0, j - u), min(n, j + l + 1)):
            out[j, u + k - j] = Adense[k, j]
    return out


def storage_to_dense(data, l, u):
    n = data.shape[0]
    A = np.zeros((n, n))
    for j in range(n):
        for k in range(max(0, j - u), min(n, j + l + 1)):
            A[k, j] = data[j, u + k - j]
    return A


def dense_qr_unblocked(A):
    """LinearAlgebra.qrfactUnblocked!(Matrix(A)) == LAPACK geqr2 convention (test_qr.jl:16): factors, tau."""
    from scipy.linalg import lapack
    qr_, tau, _, info = lapack.dgeqrf(np.asfortranarray(A))
    assert info == 0
    n = A.shape[0]
    t = np.zeros(n); t[: len(tau)] = tau
    t[n - 1] = 0.0  # geqrf's last tau for a square real matrix is 0
    return qr_, t


# --------------------------------------------------------------------------------------------------
# BandedPlusSemiseparable — reference layout, stacked: B (nb, n, l+m+1); U,V (nb, r, n); W,S (nb, p, n); b (nb, n)
# --------------------------------------------------------------------------------------------------
def bps_generate(n, l, m, r, p, nb, seed, s0=0):
    B = np.empty((nb, n, l + m + 1)); U = np.empty((nb, r, n)); V = np.empty((nb, r, n))
    W = np.empty((nb, p, n)); S = np.empty((nb, p, n)); b = np.empty((nb, n))
    lib().ssm_oracle_bps_generate(n, l, m, r, p, C.c_int64(s0), C.c_int64(nb), C.c_uint64(seed), _p(B), _p(U), _p(V), _p(W), _p(S), _p(b))
    return B, U, V, W, S, b


def bps_dense(B, U, V, W, S, l, m):
    """Dense n x n matrix of ONE system: B + tril(UV',-1) + triu(WS',1) (BandedPlusSemiseparableMatrix.jl:2)."""
    n = B.shape[0]; r = U.shape[0]; p = W.shape[0]
    out = np.empty((n, n))
    lib().ssm_oracle_bps_dense(n, l, m, r, p, _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)), _p(out))
    return out.T.copy()


def bps_qr(B, U, V, W, S, l, m):
    """qr(A::BPS) of one system -> (Bf (n,2l+m+1), Vf (r,n), Wf (p+r,n), Sf (p+r,n), tau (n,))  [semiseparableqr.jl:128-131]"""
    n = B.shape[0]; r = U.shape[0]; p = W.shape[0]
    Bf = np.zeros((n, 2 * l + m + 1)); Vf = np.zeros((r, n)); Wf = np.zeros((p + r, n)); Sf = np.zeros((p + r, n)); tau = np.zeros(n)
    rc = lib().ssm_oracle_bps_qr(n, l, m, r, p, _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)),
                                 _p(Bf), _p(Vf), _p(Wf), _p(Sf), _p(tau))
    assert rc == 0
    return Bf, Vf, Wf, Sf, tau


def bps_qr_steps(B, U, V, W, S, l, m, nsteps):
    """The sweep of ``bps_qr`` stopped after ``nsteps`` columns -> (Bf, Vf, Wf, Sf, tau, state) with ``state`` = dict of the perturbed-factor
    matrices Q (r,p), K (r,r), E (r,lm), X (lx,p), Y (lx,r), Z (lx,lm)  [semiseparableqr.jl:14-48; invariant :1-12, test_getindex.jl:32-38]."""
    n = B.shape[0]; r = U.shape[0]; p = W.shape[0]
    lm, lx = min(l + m, n), min(l, n)
    Bf = np.zeros((n, 2 * l + m + 1)); Vf = np.zeros((r, n)); Wf = np.zeros((p + r, n)); Sf = np.zeros((p + r, n)); tau = np.zeros(n)
    sizes = [(r, p), (r, r), (r, lm), (lx, p), (lx, r), (lx, lm)]
    st = np.zeros(sum(a * b for a, b in sizes) + 1)
    rc = lib().ssm_oracle_bps_qr_steps(n, l, m, r, p, _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)), int(nsteps),
                                       _p(Bf), _p(Vf), _p(Wf), _p(Sf), _p(tau), _p(st))
    assert rc == 0
    out, o = {}, 0
    for name, (a, b) in zip("QKEXYZ", sizes):
        out[name] = st[o:o + a * b].reshape((b, a)).T.copy()      # column-major (a x b)
        o += a * b
    return Bf, Vf, Wf, Sf, tau, out


def bps_qr_batch(B, U, V, W, S, l, m):
    nb, n, _ = B.shape; r = U.shape[1]; p = W.shape[1]
    Bf = np.zeros((nb, n, 2 * l + m + 1)); Vf = np.zeros((nb, r, n)); Wf = np.zeros((nb, p + r, n)); Sf = np.zeros((nb, p + r, n)); tau = np.zeros((nb, n))
    rc = lib().ssm_oracle_bps_qr_batch(n, l, m, r, p, C.c_int64(nb), _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)),
                                       _p(Bf), _p(Vf), _p(Wf), _p(Sf), _p(tau))
    assert rc == 0
    return Bf, Vf, Wf, Sf, tau


def bps_lmul_qt(Bf, U, Vf, tau, b, l, mw):
    """lmul!(F.Q', b) [semiseparableqr.jl:863-905]; mw = upper bandwidth of the factor's B (= l + m)."""
    n = Bf.shape[0]; r = U.shape[0]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_bps_lmul_qt(n, l, mw, r, _p(_f64(Bf)), _p(_f64(U)), _p(_f64(Vf)), _p(_f64(tau)), _p(x))
    return x


def bps_upper_ldiv(Bf, Wf, Sf, b, l, mw):
    """ldiv!(UpperTriangular(F), b) [BandedPlusSemiseparableMatrix.jl:53-70]."""
    n = Bf.shape[0]; pw = Wf.shape[0]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_bps_upper_ldiv(n, l, mw, pw, _p(_f64(Bf)), _p(_f64(Wf)), _p(_f64(Sf)), _p(x))
    return x


def bps_solve_batch(Bf, U, Vf, Wf, Sf, tau, b, l, mw):
    nb, n, _ = Bf.shape; r = U.shape[1]; pw = Wf.shape[1]
    x = np.array(b, dtype=np.float64, copy=True)
    lib().ssm_oracle_bps_solve_batch(n, l, mw, r, pw, C.c_int64(nb), _p(_f64(Bf)), _p(_f64(U)), _p(_f64(Vf)), _p(_f64(Wf)), _p(_f64(Sf)), _p(_f64(tau)), _p(x))
    return x


def bps_mul(B, U, V, W, S, x, l, m):
    n = B.shape[0]; r = U.shape[0]; p = W.shape[0]
    y = np.empty(n)
    lib().ssm_oracle_bps_mul(n, l, m, r, p, _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)), _p(_f64(x)), _p(y))
    return y


def bps_relres(B, U, V, W, S, x, b, l, m):
    n = B.shape[0]; r = U.shape[0]; p = W.shape[0]
    return float(lib().ssm_oracle_bps_relres(n, l, m, r, p, _p(_f64(B)), _p(_f64(U)), _p(_f64(V)), _p(_f64(W)), _p(_f64(S)), _p(_f64(x)), _p(_f64(b))))


def bps_factors_dense(Bf, U, Vf, Wf, Sf, l, mw):
    """Matrix(F.factors): R in the upper triangle, reflector tails below (what test_qr.jl:17,29 compares)."""
    return bps_dense(Bf, U, Vf, Wf, Sf, l, mw)
