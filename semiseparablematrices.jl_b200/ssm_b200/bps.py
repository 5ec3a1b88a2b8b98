"""Host-side mirror of the reference's BandedPlusSemiseparableMatrix interface, batched, on libssmb200.so.

Names follow /root/reference/src: ``BandedPlusSemiseparableMatrix(B, (U,V), (W,S))``
(BandedPlusSemiseparableMatrix.jl:1-16), ``qr`` (semiseparableqr.jl:128-131), ``F.Q`` / ``F.R`` (:852-861),
``lmul!(Q', b)`` (:863-905), ``ldiv!(UpperTriangular(F), b)`` (BandedPlusSemiseparableMatrix.jl:53-70),
``A * b`` (:21-41).

Array convention = the Julia arrays' memory, stacked:  B (nb, n, l+m+1) [column-major (l+m+1) x n band storage];
U, V (nb, r, n) [column-major n x r];  W, S (nb, p, n);  b (nb, n).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DimensionMismatch, check, c_vp
from .almostbanded import Context, Vec, _TorchOrder, _as_ptr, _is_torch, _to_vec, default_context


class BandedPlusSemiseparableMatrix:
    """``B + tril(U V', -1) + triu(W S', 1)``, a batch of nb systems."""

    def __init__(self, B, UV, WS, lm, ctx: Context | None = None):
        l, m = (int(x) for x in lm)
        U, V = UV
        W, S = WS
        single = (np.ndim(B) == 2) if not _is_torch(B) else (B.dim() == 2)
        if single:
            B, U, V, W, S = B[None], U[None], V[None], W[None], S[None]
        nb, n, w = tuple(B.shape)
        r, p = int(U.shape[1]), int(W.shape[1])
        ok = (w == l + m + 1 and tuple(U.shape) == (nb, r, n) and tuple(V.shape) == (nb, r, n)
              and tuple(W.shape) == (nb, p, n) and tuple(S.shape) == (nb, p, n))
        if not ok:
            raise DimensionMismatch("Dimensions are not compatible.")     # BandedPlusSemiseparableMatrix.jl:14
        self.ctx = ctx or default_context()
        self.n, self.l, self.m, self.r, self.p, self.nb, self.single = n, l, m, r, p, nb, single
        self._h = c_vp()
        check(_lib.lib().ssm_bps_create(self.ctx._h, n, l, m, r, p, nb, C.byref(self._h)), "ssm_bps_create")
        ptrs = [_as_ptr(a, None, nm) for a, nm in ((B, "B"), (U, "U"), (V, "V"), (W, "W"), (S, "S"))]
        with _TorchOrder(self.ctx, *[k for _, k in ptrs]):
            check(_lib.lib().ssm_bps_set(self._h, c_vp(ptrs[0][0]), 0, c_vp(ptrs[1][0]), c_vp(ptrs[2][0]), 0,
                                         c_vp(ptrs[3][0]), c_vp(ptrs[4][0]), 0), "ssm_bps_set")

    @classmethod
    def generate(cls, n, l, m, r, p, nb, seed, ctx: Context | None = None):
        """Device-side synthetic 'dd' operands; bit-identical to ``oracle.bps_generate``."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.n, self.l, self.m, self.r, self.p, self.nb, self.single = n, l, m, r, p, nb, False
        self._h = c_vp()
        check(_lib.lib().ssm_bps_create(self.ctx._h, n, l, m, r, p, nb, C.byref(self._h)), "ssm_bps_create")
        check(_lib.lib().ssm_bps_generate(self._h, seed), "ssm_bps_generate")
        return self

    @classmethod
    def _wrap(cls, ctx, handle, n, l, m, r, p, nb, single):
        self = cls.__new__(cls)
        self.ctx, self._h = ctx, handle
        self.n, self.l, self.m, self.r, self.p, self.nb, self.single = n, l, m, r, p, nb, single
        return self

    def get(self):
        """(B (nb,n,l+m+1), U (nb,r,n), V (nb,r,n), W (nb,p,n), S (nb,p,n)) — the operands in the reference layout"""
        n, l, m, r, p, nb = self.n, self.l, self.m, self.r, self.p, self.nb
        B = np.zeros((nb, n, l + m + 1)); U = np.empty((nb, r, n)); V = np.empty((nb, r, n)); W = np.empty((nb, p, n)); S = np.empty((nb, p, n))
        P = lambda a: c_vp(a.ctypes.data)
        check(_lib.lib().ssm_bps_get(self._h, P(B), 0, P(U), P(V), 0, P(W), P(S), 0), "ssm_bps_get")
        return B, U, V, W, S

    @property
    def shape(self):
        return (self.n, self.n)

    def mul(self, x):
        """``A * x`` (BandedPlusSemiseparableMatrix.jl:21-41)."""
        xv, inplace = _to_vec(self.ctx, self.n, self.nb, x)
        y = Vec(self.ctx, self.n, self.nb)
        check(_lib.lib().ssm_bps_mul(self._h, xv._h, y._h), "ssm_bps_mul")
        if inplace:
            return y
        out = y.get()
        return out[0] if self.single else out

    __matmul__ = mul

    def residual(self, x, b):
        xv, _ = _to_vec(self.ctx, self.n, self.nb, x)
        bv, _ = _to_vec(self.ctx, self.n, self.nb, b)
        out = np.empty(self.nb)
        check(_lib.lib().ssm_bps_residual(self._h, xv._h, bv._h, c_vp(out.ctypes.data)), "ssm_bps_residual")
        return out

    def upper_ldiv(self, b):
        """``ldiv!(UpperTriangular(A), b)`` on the unfactored matrix (test_bandedplussemiseparable.jl:17)."""
        v, inplace = _to_vec(self.ctx, self.n, self.nb, b)
        check(_lib.lib().ssm_bps_upper_ldiv(self._h, v._h), "ssm_bps_upper_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if self.single else x

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_bps_destroy(self._h)
        except Exception:
            pass


class _BPSQ:
    def __init__(self, F, adj=False):
        self.F, self.adj = F, adj

    def adjoint(self):
        return _BPSQ(self.F, not self.adj)

    T = property(adjoint)

    def lmul(self, b):
        if not self.adj:
            raise NotImplementedError("the reference defines lmul! only for Q' (semiseparableqr.jl:863)")
        F = self.F
        v, inplace = _to_vec(F.ctx, F.n, F.nb, b)
        check(_lib.lib().ssm_bpsqr_lmul_qt(F._h, v._h), "ssm_bpsqr_lmul_qt")
        if inplace:
            return v
        x = v.get()
        return x[0] if F.single else x


class _BPSR:
    def __init__(self, F):
        self.F = F

    def solve(self, b):
        F = self.F
        v, inplace = _to_vec(F.ctx, F.n, F.nb, b)
        check(_lib.lib().ssm_bpsqr_upper_ldiv(F._h, v._h), "ssm_bpsqr_upper_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if F.single else x


class BPSQR:
    """``QR(BandedPlusSemiseparableMatrix(F.factors), tau)`` (semiseparableqr.jl:130)."""

    def __init__(self, A: BandedPlusSemiseparableMatrix):
        self.ctx, self.n, self.l, self.m, self.r, self.p, self.nb, self.single = A.ctx, A.n, A.l, A.m, A.r, A.p, A.nb, A.single
        self._A = A
        self._h = c_vp()
        check(_lib.lib().ssm_bps_qr(A._h, C.byref(self._h)), "ssm_bps_qr")

    def refactor(self, A):
        check(_lib.lib().ssm_bps_qr(A._h, C.byref(self._h)), "ssm_bps_qr")
        self._A = A
        return self

    def get(self):
        """(B (nb,n,2l+m+1) bandwidths (l,l+m), U (nb,r,n), V (nb,r,n), W (nb,p+r,n), S (nb,p+r,n), tau (nb,n))"""
        n, l, m, r, p, nb = self.n, self.l, self.m, self.r, self.p, self.nb
        B = np.zeros((nb, n, 2 * l + m + 1)); U = np.empty((nb, r, n)); V = np.empty((nb, r, n))
        W = np.empty((nb, p + r, n)); S = np.empty((nb, p + r, n)); tau = np.empty((nb, n))
        P = lambda a: c_vp(a.ctypes.data)
        check(_lib.lib().ssm_bpsqr_get(self._h, P(B), 0, P(U), P(V), 0, P(W), P(S), 0, P(tau), 0), "ssm_bpsqr_get")
        return B, U, V, W, S, tau

    def rq_mul(self) -> BandedPlusSemiseparableMatrix:
        """``rq_mul(F.factors, F.tau)`` (semiseparableqr.jl:282-287): R*Q = Q'AQ of a SYMMETRIC BandedPlusSemiseparableMatrix, again as
        a symmetric BPS batch ``BPS(B (l,l), (U_new, V_new), (V_new, U_new))`` — one step of the batched QR iteration."""
        h = c_vp()
        check(_lib.lib().ssm_bpsqr_rq_mul(self._h, C.byref(h)), "ssm_bpsqr_rq_mul")
        return BandedPlusSemiseparableMatrix._wrap(self.ctx, h, self.n, self.l, self.l, self.r, self.r, self.nb, self.single)

    @property
    def factors(self):
        return self.get()[:5]

    @property
    def tau(self):
        return self.get()[5]

    @property
    def Q(self):
        return _BPSQ(self)

    @property
    def R(self):
        return _BPSR(self)

    def ldiv(self, b):
        """``ldiv!(F.R, lmul!(F.Q', b))`` — the two halves test_qr.jl:47-53 exercises, fused."""
        v, inplace = _to_vec(self.ctx, self.n, self.nb, b)
        check(_lib.lib().ssm_bpsqr_ldiv(self._h, v._h), "ssm_bpsqr_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if self.single else x

    def solve(self, b, out: Vec | None = None):
        if out is None:
            return self.ldiv(b.copy() if isinstance(b, Vec) else b)
        v, _ = _to_vec(self.ctx, self.n, self.nb, b)
        check(_lib.lib().ssm_bpsqr_solve(self._h, v._h, out._h), "ssm_bpsqr_solve")
        return out

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_bpsqr_destroy(self._h)
        except Exception:
            pass


def qr(A: BandedPlusSemiseparableMatrix) -> BPSQR:
    return BPSQR(A)
