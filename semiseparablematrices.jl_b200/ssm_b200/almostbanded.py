"""Host-side mirror of the reference's AlmostBandedMatrix operator interface, batched, on libssmb200.so.

Names follow /root/reference/src/AlmostBandedMatrix.jl: ``AlmostBandedMatrix(bands, fill)`` (:7-24),
``almostbandwidths`` / ``almostbandedrank`` (:65-68), ``bandpart`` / ``fillpart`` (:62-63), ``qr`` (:125-133),
``ldiv`` (= ``ldiv!``, :187-192), ``solve`` (= ``A \\ b``), ``F.Q`` / ``F.R`` (:181-185),
``UpperTriangular(A) \\ b`` & friends (:242-270).  A batch is a stack of same-shaped systems.

Array convention = the Julia arrays' memory, stacked (numpy C order):
    bands (nb, n, l+u+1)  <->  per system the column-major (l+u+1) x n ``BandedMatrix.data``
    U     (nb, r, n)      <->  column-major n x r   (``A.fill.args[1]``)
    V     (nb, n, r)      <->  column-major r x n   (``A.fill.args[2]``)
    b     (nb, n)
numpy arrays are host buffers (staged through the library); torch CUDA tensors are passed by device pointer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DimensionMismatch, SSMError, check, c_vp


def _is_torch(a) -> bool:
    return type(a).__module__.split(".")[0] == "torch"


def _as_ptr(a, shape=None, name="array"):
    """(address, keepalive) of a float64 contiguous host ndarray or torch tensor (host or CUDA)."""
    if a is None:
        return None, None
    if _is_torch(a):
        import torch
        if a.dtype != torch.float64:
            raise TypeError(f"{name}: expected float64")
        t = a.contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise DimensionMismatch(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t.data_ptr(), t
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(arr.shape) != tuple(shape):
        raise DimensionMismatch(f"{name}: expected shape {tuple(shape)}, got {tuple(arr.shape)}")
    return arr.ctypes.data, arr


class _TorchOrder:
    """Stream ordering between torch and the library for calls that take torch CUDA tensors (the C ABI only ENQUEUES on its
    context's stream when it is given device pointers).  On entry the context's stream waits for torch's current stream — the
    tensors may still be being produced; on exit torch's current stream waits for the context's — the outputs are ready for
    whatever torch enqueues next, and a temporary from ``.contiguous()`` that is freed afterwards can only be handed out again
    to work that torch's stream orders behind the library's use of it (torch's allocator reuses a block on the stream it was
    allocated on).  ``Tensor.record_stream`` is deliberately NOT used: the allocator would later record an event on the
    context's stream, which may be gone by then.  Nothing happens for host arrays, nor when the context already runs on torch's
    current stream."""

    def __init__(self, ctx, *arrays):
        self.ctx = ctx
        self.ts = [a for a in arrays if a is not None and _is_torch(a) and a.is_cuda]
        self.ext = None

    def __enter__(self):
        if self.ts:
            import torch
            dev = self.ts[0].device
            self.cur = torch.cuda.current_stream(dev)
            sp = self.ctx.stream_ptr
            if self.cur.cuda_stream != sp:
                self.ext = torch.cuda.ExternalStream(sp, device=dev)
                self.ext.wait_stream(self.cur)
        return self

    def __exit__(self, *exc):
        if self.ext is not None:
            self.cur.wait_stream(self.ext)
        return False


class Context:
    """Device + stream + staging workspace (``ssm_ctx``).  ``stream`` is a raw ``cudaStream_t`` (e.g.
    ``torch.cuda.current_stream().cuda_stream``) or None for a private non-blocking stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = c_vp()
        L = _lib.lib()
        if stream is None:
            check(L.ssm_ctx_create(device, C.byref(self._h)), "ssm_ctx_create")
        else:
            check(L.ssm_ctx_create_on_stream(device, c_vp(stream), C.byref(self._h)), "ssm_ctx_create_on_stream")
        self.device = device

    def sync(self):
        check(_lib.lib().ssm_ctx_sync(self._h), "ssm_ctx_sync")

    def set_option(self, key: str, value: int):
        check(_lib.lib().ssm_ctx_set_option(self._h, key.encode(), int(value)), "ssm_ctx_set_option")

    def get_option(self, key: str) -> int:
        v = C.c_int()
        check(_lib.lib().ssm_ctx_get_option(self._h, key.encode(), C.byref(v)), "ssm_ctx_get_option")
        return v.value

    @property
    def launches(self) -> int:
        return int(_lib.lib().ssm_ctx_launch_count(self._h))

    @property
    def stream_ptr(self) -> int:
        """the raw ``cudaStream_t`` this context enqueues on"""
        return int(_lib.lib().ssm_ctx_stream(self._h) or 0)

    def close(self):
        if self._h:
            _lib.lib().ssm_ctx_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: dict[int, Context] = {}


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def pool_stats(device: int = 0) -> dict:
    """The per-device free list of small device blocks (``ssm_pool_stats``): bytes and blocks listed, hits and misses so far."""
    v = [C.c_int64() for _ in range(4)]
    check(_lib.lib().ssm_pool_stats(int(device), *[C.byref(x) for x in v]), "ssm_pool_stats")
    return dict(zip(("bytes", "blocks", "hits", "misses"), (int(x.value) for x in v)))


def pool_trim(device: int = 0) -> int:
    """Give the listed blocks of ``device`` back to the driver; returns the bytes freed (``ssm_pool_trim``)."""
    return int(_lib.lib().ssm_pool_trim(int(device)))


class Vec:
    """Batch of length-n vectors in the device layout (``ssm_vec``)."""

    def __init__(self, ctx: Context, n: int, nb: int, data=None):
        self.ctx, self.n, self.nb = ctx, int(n), int(nb)
        self._h = c_vp()
        check(_lib.lib().ssm_vec_create(ctx._h, self.n, self.nb, C.byref(self._h)), "ssm_vec_create")
        if data is not None:
            self.set(data)

    def set(self, b):
        p, keep = _as_ptr(b, (self.nb, self.n), "b")
        with _TorchOrder(self.ctx, keep):
            check(_lib.lib().ssm_vec_set(self._h, c_vp(p), 0), "ssm_vec_set")
        return self

    def get(self, out=None):
        if out is None:
            out = np.empty((self.nb, self.n))
        p, keep = _as_ptr(out, (self.nb, self.n), "out")
        if keep is not out and not _is_torch(out):
            raise TypeError("out must be a contiguous float64 array")
        with _TorchOrder(self.ctx, keep):
            check(_lib.lib().ssm_vec_get(self._h, c_vp(p), 0), "ssm_vec_get")
        return out

    def copy(self) -> "Vec":
        v = Vec(self.ctx, self.n, self.nb)
        check(_lib.lib().ssm_vec_copy(v._h, self._h), "ssm_vec_copy")
        return v

    def generate(self, seed: int, array_id: int = 3):
        check(_lib.lib().ssm_vec_generate(self._h, seed, array_id), "ssm_vec_generate")
        return self

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_vec_destroy(self._h)
        except Exception:
            pass


def _to_vec(ctx, n, nb, b) -> tuple[Vec, bool]:
    if isinstance(b, Vec):
        if b.n != n or b.nb != nb:
            raise DimensionMismatch(f"vector batch is {b.nb} x {b.n}, matrix batch is {nb} x {n}")
        return b, True
    arr = b
    if not _is_torch(b):
        arr = np.asarray(b, dtype=np.float64)
        if arr.ndim == 1:
            arr = arr[None, :]
    if tuple(arr.shape) != (nb, n):
        raise DimensionMismatch(f"right-hand side has shape {tuple(arr.shape)}, expected {(nb, n)}")
    return Vec(ctx, n, nb, arr), False


class AlmostBandedMatrix:
    """``bands + triu(U*V, u+1)`` (AlmostBandedMatrix.jl:7-18, getindex :84-90), a batch of nb systems."""

    def __init__(self, bands, fill, lu, ctx: Context | None = None):
        l, u = (int(x) for x in lu)
        U, V = fill
        single = (np.ndim(bands) == 2) if not _is_torch(bands) else (bands.dim() == 2)
        if single:
            bands, U, V = bands[None], U[None], V[None]
        nb, n, w = tuple(bands.shape)
        if w != l + u + 1:
            raise DimensionMismatch(f"band storage has {w} rows, bandwidths ({l},{u}) need {l + u + 1}")
        r = int(U.shape[1])
        if tuple(U.shape) != (nb, r, n) or tuple(V.shape) != (nb, n, r):
            # reference: error("Data and fill must be compatible size") (:14-16)
            raise DimensionMismatch("Data and fill must be compatible size")
        self.ctx = ctx or default_context()
        self.n, self.l, self.u, self.r, self.nb, self.single = n, l, u, r, nb, single
        self._h = c_vp()
        check(_lib.lib().ssm_ab_create(self.ctx._h, n, l, u, r, nb, C.byref(self._h)), "ssm_ab_create")
        pb, kb = _as_ptr(bands, (nb, n, w), "bands")
        pu, ku = _as_ptr(U, (nb, r, n), "U")
        pv, kv = _as_ptr(V, (nb, n, r), "V")
        with _TorchOrder(self.ctx, kb, ku, kv):
            check(_lib.lib().ssm_ab_set(self._h, c_vp(pb), 0, c_vp(pu), 0, c_vp(pv), 0), "ssm_ab_set")

    @classmethod
    def _empty(cls, ctx, n, l, u, r, nb):
        self = cls.__new__(cls)
        self.ctx, self.n, self.l, self.u, self.r, self.nb, self.single = ctx, n, l, u, r, nb, False
        self._h = c_vp()
        check(_lib.lib().ssm_ab_create(ctx._h, n, l, u, r, nb, C.byref(self._h)), "ssm_ab_create")
        return self

    @classmethod
    def generate(cls, n, l, u, r, nb, seed, dist=_lib.SSM_DIST_DD, ctx: Context | None = None):
        """Device-side synthetic operands; bit-identical to ``oracle.ab_generate``."""
        self = cls._empty(ctx or default_context(), n, l, u, r, nb)
        check(_lib.lib().ssm_ab_generate(self._h, seed, dist), "ssm_ab_generate")
        return self

    @classmethod
    def from_vcat(cls, a, Bd, lbub, ctx: Context | None = None):
        """``AlmostBandedMatrix(Vcat(a, Bd))`` (AlmostBandedMatrix.jl:48-49, 298-305, 316-337): ``a`` (nb, n, r) = per system the
        memory of the r x n dense boundary rows, ``Bd`` (nb, n, lb+ub+1) = band storage of the (n-r) x n banded block with
        bandwidths ``(lb, ub)``.  almostbandwidths = (max(lb+r, r-1), ub-r) (a negative upper bandwidth is stored as 0)."""
        lb, ub = (int(v) for v in lbub)
        a = np.ascontiguousarray(a, dtype=np.float64); Bd = np.ascontiguousarray(Bd, dtype=np.float64)
        single = a.ndim == 2
        if single:
            a, Bd = a[None], Bd[None]
        nb, n, r = a.shape
        if tuple(Bd.shape) != (nb, n, lb + ub + 1):
            raise DimensionMismatch(f"banded block storage has shape {Bd.shape}, expected {(nb, n, lb + ub + 1)}")
        l, u = C.c_int(), C.c_int()
        check(_lib.lib().ssm_ab_vcat_bandwidths(r, lb, ub, C.byref(l), C.byref(u)), "ssm_ab_vcat_bandwidths")
        self = cls._empty(ctx or default_context(), n, l.value, u.value, r, nb)
        self.single = single
        check(_lib.lib().ssm_ab_set_vcat(self._h, c_vp(a.ctypes.data), 0, c_vp(Bd.ctypes.data), 0, lb, ub), "ssm_ab_set_vcat")
        return self

    def resize(self, n: int) -> "AlmostBandedMatrix":
        """``resize(A, n, n)`` (AlmostBandedMatrix.jl:338-348): the same bands and fill in a larger matrix, everything new zero; on the device."""
        new = type(self).__new__(type(self))
        new.ctx, new.n, new.l, new.u, new.r, new.nb, new.single = self.ctx, int(n), self.l, self.u, self.r, self.nb, self.single
        new._h = c_vp()
        check(_lib.lib().ssm_ab_resize(self._h, int(n), C.byref(new._h)), "ssm_ab_resize")
        return new

    def set_rows(self, k0: int, k1: int, bands, fill):
        """Rows ``k0 <= i < k1`` (0-based) of every matrix — ``A[i, i-l..i+u]``, ``U[i,:]``, ``V[:,i]`` — from full arrays in the constructor's
        layout; the other rows keep their contents.  The copy ``resizedata!`` makes after ``resize`` (AlmostBandedMatrix.jl:373-389)."""
        U, V = fill
        if self.single and ((np.ndim(bands) == 2) if not _is_torch(bands) else (bands.dim() == 2)):
            bands, U, V = bands[None], U[None], V[None]
        nb, n, w, r = self.nb, self.n, self.l + self.u + 1, self.r
        if tuple(bands.shape) != (nb, n, w) or tuple(U.shape) != (nb, r, n) or tuple(V.shape) != (nb, n, r):
            raise DimensionMismatch("Data and fill must be compatible size")
        pb, kb = _as_ptr(bands, (nb, n, w), "bands")
        pu, ku = _as_ptr(U, (nb, r, n), "U")
        pv, kv = _as_ptr(V, (nb, n, r), "V")
        with _TorchOrder(self.ctx, kb, ku, kv):
            check(_lib.lib().ssm_ab_set_rows(self._h, int(k0), int(k1), c_vp(pb), 0, c_vp(pu), 0, c_vp(pv), 0), "ssm_ab_set_rows")
        return self

    # -- reference accessors ----------------------------------------------------------------------
    @property
    def shape(self):
        return (self.n, self.n)

    def almostbandwidths(self):
        return (self.l, self.u)

    def almostbandedrank(self):
        return self.r

    def _get(self):
        bands = np.empty((self.nb, self.n, self.l + self.u + 1)); U = np.empty((self.nb, self.r, self.n)); V = np.empty((self.nb, self.n, self.r))
        check(_lib.lib().ssm_ab_get(self._h, c_vp(bands.ctypes.data), 0, c_vp(U.ctypes.data), 0, c_vp(V.ctypes.data), 0), "ssm_ab_get")
        return bands, U, V

    def bandpart(self):
        return self._get()[0]

    def fillpart(self):
        _, U, V = self._get()
        return U, V

    # -- solves ------------------------------------------------------------------------------------
    def solve(self, b, out: "Vec | None" = None):
        """``A \\ b`` (:125 ``_factorize = qr`` then ``ldiv!``): fused factor + solve.  Returns x (or b in place)."""
        v, inplace = _to_vec(self.ctx, self.n, self.nb, b)
        if out is not None:
            check(_lib.lib().ssm_ab_solve(self._h, v._h, out._h), "ssm_ab_solve")
            return out
        check(_lib.lib().ssm_ab_solve(self._h, v._h, v._h), "ssm_ab_solve")
        if inplace:
            return v
        x = v.get()
        return x[0] if self.single else x

    def residual(self, x, b):
        """relres per system: ``||b - A x|| / ||b||`` with ``A`` applied through its structure."""
        xv, _ = _to_vec(self.ctx, self.n, self.nb, x)
        bv, _ = _to_vec(self.ctx, self.n, self.nb, b)
        out = np.empty(self.nb)
        check(_lib.lib().ssm_ab_residual(self._h, xv._h, bv._h, c_vp(out.ctypes.data)), "ssm_ab_residual")
        return out

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_ab_destroy(self._h)
        except Exception:
            pass


def almostbandwidths(A):
    return A.almostbandwidths()


def almostbandedrank(A):
    return A.almostbandedrank()


def bandpart(A):
    return A.bandpart()


def fillpart(A):
    return A.fillpart()


class _PackedQ:
    """``QRPackedQ(bandpart(F.factors), F.tau)`` (:181): ``lmul`` applies Q, ``adjoint().lmul`` applies Q'."""

    def __init__(self, F, adj=False):
        self.F, self.adj = F, adj

    def adjoint(self):
        return _PackedQ(self.F, not self.adj)

    T = property(adjoint)

    def lmul(self, b):
        F = self.F
        v, inplace = _to_vec(F.ctx, F.n, F.nb, b)
        fn = _lib.lib().ssm_abqr_lmul_qt if self.adj else _lib.lib().ssm_abqr_lmul_q
        check(fn(F._h, v._h), "lmul!")
        if inplace:
            return v
        x = v.get()
        return x[0] if F.single else x

    __matmul__ = lmul


class _UpperR:
    """``UpperTriangular(view(F.factors, 1:n, 1:n))`` (:182-185); ``solve`` = ``ldiv!(R, b)`` (:215-248)."""

    def __init__(self, F):
        self.F = F

    def solve(self, b):
        F = self.F
        v, inplace = _to_vec(F.ctx, F.n, F.nb, b)
        check(_lib.lib().ssm_abqr_upper_ldiv(F._h, v._h), "ssm_abqr_upper_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if F.single else x


class AlmostBandedQR:
    """``QR{Float64,<:AlmostBandedMatrix}`` (:174-179): ``.factors`` (widened bands (l, l+u) + updated fill), ``.tau``."""

    def __init__(self, A: AlmostBandedMatrix, handle=None, ncols=None):
        self.ctx, self.n, self.l, self.u, self.r, self.nb, self.single = A.ctx, A.n, A.l, A.u, A.r, A.nb, A.single
        self._A = A      # V is shared with the operand batch
        self._h = handle if handle is not None else c_vp()
        self.ncols = A.n - 1 if ncols is None else int(ncols)      # min(m - 1, n) for a real square matrix (:137)
        check(_lib.lib().ssm_ab_qr_cols(A._h, self.ncols, C.byref(self._h)), "ssm_ab_qr_cols")

    @classmethod
    def from_factors(cls, factors, U, V, tau, lu, ncols=None, ctx: Context | None = None):
        """A factor object from the reference's own fields (``ssm_abqr_set``): ``factors`` (nb, n, 2l+u+1) widened band storage, ``U`` (nb, r, n),
        ``V`` (nb, n, r), ``tau`` (nb, n) — what ``get()`` / ``factors`` / ``tau`` return.  For a QR that was not made by this library instance."""
        l, u = (int(x) for x in lu)
        factors = np.ascontiguousarray(factors, dtype=np.float64); U = np.ascontiguousarray(U, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64); tau = np.ascontiguousarray(tau, dtype=np.float64)
        single = factors.ndim == 2
        if single:
            factors, U, V, tau = factors[None], U[None], V[None], tau[None]
        nb, n, w = factors.shape
        r = int(U.shape[1])
        if w != 2 * l + u + 1 or U.shape != (nb, r, n) or V.shape != (nb, n, r) or tau.shape != (nb, n):
            raise DimensionMismatch("Data and fill must be compatible size")
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self.n, self.l, self.u, self.r, self.nb, self.single, self._A, self._V = n, l, u, r, nb, single, None, V
        self.ncols = n - 1 if ncols is None else int(ncols)
        self._h = c_vp()
        P = lambda a: c_vp(a.ctypes.data)
        check(_lib.lib().ssm_abqr_set(self.ctx._h, n, l, u, r, nb, P(factors), 0, P(U), 0, P(V), 0, P(tau), 0, self.ncols, C.byref(self._h)), "ssm_abqr_set")
        return self

    def refactor(self, A: AlmostBandedMatrix):
        """qr into the existing factor storage (no allocation): steady-state use and benchmarking."""
        check(_lib.lib().ssm_ab_qr_cols(A._h, self.ncols, C.byref(self._h)), "ssm_ab_qr_cols")
        self._A = A
        return self

    def continue_to(self, ncols: int):
        """Form reflectors ``self.ncols .. ncols-1`` in place (``ssm_abqr_continue``); afterwards equal to ``qr(A, ncols=ncols)``."""
        check(_lib.lib().ssm_abqr_continue(self._h, self._A._h, int(ncols)), "ssm_abqr_continue")
        self.ncols = int(ncols)
        return self

    def resize(self, A: AlmostBandedMatrix):
        """Grow this partial factorisation to the size of ``A`` — the operand it was made from after ``A.resize(n)`` + ``set_rows`` (the
        adaptive-QR growth step, AlmostBandedMatrix.jl:357-394); afterwards equal to ``qr(A, ncols=self.ncols)``; ``continue_to`` goes on."""
        check(_lib.lib().ssm_abqr_resize(self._h, A._h), "ssm_abqr_resize")
        self._A, self.n = A, A.n
        return self

    def get(self):
        n, l, u, r, nb = self.n, self.l, self.u, self.r, self.nb
        F = np.zeros((nb, n, 2 * l + u + 1)); U = np.empty((nb, r, n)); tau = np.empty((nb, n))
        check(_lib.lib().ssm_abqr_get(self._h, c_vp(F.ctypes.data), 0, c_vp(U.ctypes.data), 0, c_vp(tau.ctypes.data), 0), "ssm_abqr_get")
        return F, U, tau

    @property
    def factors(self):
        """(bands (nb,n,2l+u+1) with bandwidths (l, l+u), U (nb,r,n), V (nb,n,r))"""
        F, U, _ = self.get()
        return F, U, (self._A.fillpart()[1] if self._A is not None else self._V)

    @property
    def tau(self):
        return self.get()[2]

    @property
    def Q(self):
        return _PackedQ(self)

    @property
    def R(self):
        return _UpperR(self)

    def ldiv(self, b):
        """``ldiv!(F, b)`` (:187-192): b <- R \\ (Q'b)."""
        v, inplace = _to_vec(self.ctx, self.n, self.nb, b)
        check(_lib.lib().ssm_abqr_ldiv(self._h, v._h), "ssm_abqr_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if self.single else x

    def ldiv_mat(self, B):
        """``ldiv!(F, B::AbstractMatrix)`` (:194-199): B has shape (nb, nrhs, n) = per system the memory of a Julia n x nrhs
        Matrix; every column is solved in one pass over the factors per group of 4 / 8 columns.  Returns an array like B."""
        arr = np.ascontiguousarray(B, dtype=np.float64)
        if self.single and arr.ndim == 2:
            arr = arr[None]
        if arr.ndim != 3 or arr.shape[0] != self.nb or arr.shape[2] != self.n:
            raise DimensionMismatch(f"right-hand side has shape {arr.shape}, expected ({self.nb}, nrhs, {self.n})")
        nrhs = int(arr.shape[1])
        h = c_vp()
        L = _lib.lib()
        check(L.ssm_mat_create(self.ctx._h, self.n, nrhs, self.nb, C.byref(h)), "ssm_mat_create")
        try:
            check(L.ssm_mat_set(h, c_vp(arr.ctypes.data), 0), "ssm_mat_set")
            check(L.ssm_abqr_ldiv_mat(self._h, h), "ssm_abqr_ldiv_mat")
            out = np.empty_like(arr)
            check(L.ssm_mat_get(h, c_vp(out.ctypes.data), 0), "ssm_mat_get")
        finally:
            L.ssm_mat_destroy(h)
        return out[0] if (self.single and np.ndim(B) == 2) else out

    def solve(self, b, out: "Vec | None" = None):
        """``F \\ b`` = ``ldiv!(F, copy(b))``; with ``out`` (a Vec) nothing is copied back to the host."""
        if out is None:
            return self.ldiv(b.copy() if isinstance(b, Vec) else b)
        v, _ = _to_vec(self.ctx, self.n, self.nb, b)
        check(_lib.lib().ssm_abqr_solve(self._h, v._h, out._h), "ssm_abqr_solve")
        return out

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_abqr_destroy(self._h)
        except Exception:
            pass


def qr(A: AlmostBandedMatrix, ncols=None) -> AlmostBandedQR:
    """``qr(A)`` / ``factorize(A)`` (:125-133).  ``A`` is not mutated (test_almostbanded.jl:146).
    ``ncols``: ``_almostbanded_qr!(A, tau, ncols)`` (:139-172) — only the first ``ncols`` reflectors; the rest of ``factors`` is
    the partially updated trailing matrix."""
    return AlmostBandedQR(A, ncols=ncols)


factorize = qr


def ldiv(F, b):
    """``ldiv!(F, b)`` for a factorisation, or a triangular wrapper's solve."""
    return F.ldiv(b) if hasattr(F, "ldiv") else F.solve(b)


def solve(A, b):
    """``A \\ b``."""
    return A.solve(b)


class _Tri:
    def __init__(self, A: AlmostBandedMatrix, upper: bool, unit: bool):
        self.A, self.upper, self.unit = A, upper, unit

    def solve(self, b):
        A = self.A
        v, inplace = _to_vec(A.ctx, A.n, A.nb, b)
        check(_lib.lib().ssm_ab_tri_ldiv(A._h, v._h, int(self.upper), int(self.unit)), "ssm_ab_tri_ldiv")
        if inplace:
            return v
        x = v.get()
        return x[0] if A.single else x


def UpperTriangular(A):          # AlmostBandedMatrix.jl:242-248
    return _Tri(A, True, False)


def UnitUpperTriangular(A):      # :250-256
    return _Tri(A, True, True)


def LowerTriangular(A):          # :258-263
    return _Tri(A, False, False)


def UnitLowerTriangular(A):      # :265-270
    return _Tri(A, False, True)


def solve_host(n, l, u, r, bands, U, V, b, x=None, ctx: Context | None = None):
    """End-to-end host-buffer path (``ssm_ab_solve_host``): H2D + pack + qr + ldiv! + D2H, chunk pipelined."""
    ctx = ctx or default_context()
    nb = int(b.shape[0])
    if x is None:
        x = np.empty((nb, n))
    pb, kb = _as_ptr(bands, (nb, n, l + u + 1), "bands")
    pu, ku = _as_ptr(U, (nb, r, n), "U")
    pv, kv = _as_ptr(V, (nb, n, r), "V")
    pr, kr = _as_ptr(b, (nb, n), "b")
    px, kx = _as_ptr(x, (nb, n), "x")
    check(_lib.lib().ssm_ab_solve_host(ctx._h, n, l, u, r, nb, c_vp(pb), c_vp(pu), c_vp(pv), c_vp(pr), c_vp(px)), "ssm_ab_solve_host")
    return x
