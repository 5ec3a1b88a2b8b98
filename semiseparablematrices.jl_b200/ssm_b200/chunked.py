"""One LARGE AlmostBandedMatrix system solved by the chunked n-axis path (kernel family K10, csrc/abc_kernels.cu).

``LargeAlmostBandedMatrix(bands, (U, V), (l, u)).solve(b)`` is ``A \\ b`` (AlmostBandedMatrix.jl:125-133,187-199) for a
single system; arrays use the reference memory: bands (n, l+u+1), U (r, n), V (n, r), b (n)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DimensionMismatch, check, c_vp
from .almostbanded import Context, _TorchOrder, _as_ptr, _is_torch, default_context


class LargeAlmostBandedMatrix:
    def __init__(self, bands, fill, lu, ctx: Context | None = None):
        l, u = (int(v) for v in lu)
        U, V = fill
        n, w = tuple(bands.shape)
        if w != l + u + 1:
            raise DimensionMismatch(f"band storage has {w} rows, bandwidths ({l},{u}) need {l + u + 1}")
        r = int(U.shape[0])
        if tuple(U.shape) != (r, n) or tuple(V.shape) != (n, r):
            raise DimensionMismatch("Data and fill must be compatible size")
        self._init(ctx or default_context(), n, l, u, r)
        pb, kb = _as_ptr(bands, (n, w), "bands")
        pu, ku = _as_ptr(U, (r, n), "U")
        pv, kv = _as_ptr(V, (n, r), "V")
        with _TorchOrder(self.ctx, kb, ku, kv):
            check(_lib.lib().ssm_abc_set(self._h, c_vp(pb), c_vp(pu), c_vp(pv)), "ssm_abc_set")

    def _init(self, ctx, n, l, u, r):
        self.ctx, self.n, self.l, self.u, self.r = ctx, int(n), l, u, r
        self._h = c_vp()
        check(_lib.lib().ssm_abc_create(ctx._h, self.n, l, u, r, C.byref(self._h)), "ssm_abc_create")

    @classmethod
    def generate(cls, n, l, u, r, seed, dist=_lib.SSM_DIST_BC, ctx: Context | None = None):
        self = cls.__new__(cls)
        self._init(ctx or default_context(), n, l, u, r)
        check(_lib.lib().ssm_abc_generate(self._h, seed, dist), "ssm_abc_generate")
        return self

    def generate_rhs(self, b_dev, seed):
        """fill a torch CUDA float64 tensor of length n with the shared generator's b (array id 3)"""
        with _TorchOrder(self.ctx, b_dev):
            check(_lib.lib().ssm_abc_vec_generate(self._h, c_vp(b_dev.data_ptr()), seed), "ssm_abc_vec_generate")
        self.ctx.sync()
        return b_dev

    def solve(self, b, out=None, nchunks: int = 0, sync: bool = True):
        """x = A \\ b.  b / out: numpy (host) or torch CUDA tensors of length n.  With device tensors the work is only
        enqueued on the context's stream; ``sync`` waits for it (pass False when timing with events on that stream)."""
        pb, kb = _as_ptr(b, (self.n,), "b")
        if out is None:
            if _is_torch(b):
                import torch
                out = torch.empty_like(b)
            else:
                out = np.empty(self.n)
        po, ko = _as_ptr(out, (self.n,), "out")
        with _TorchOrder(self.ctx, kb, ko):
            check(_lib.lib().ssm_abc_solve(self._h, c_vp(pb), c_vp(po), int(nchunks)), "ssm_abc_solve")
        if sync and _is_torch(out):
            self.ctx.sync()
        return out

    def residual(self, x, b) -> float:
        px, kx = _as_ptr(x, (self.n,), "x")
        pb, kb = _as_ptr(b, (self.n,), "b")
        out = C.c_double()
        with _TorchOrder(self.ctx, kx, kb):
            check(_lib.lib().ssm_abc_residual(self._h, c_vp(px), c_vp(pb), C.byref(out)), "ssm_abc_residual")
        return out.value

    @property
    def chunks(self) -> int:
        return int(_lib.lib().ssm_abc_chunks(self._h))

    def _debug(self, what: int, count: int):
        out = np.empty(count)
        check(_lib.lib().ssm_abc_debug_get(self._h, what, c_vp(out.ctypes.data), count), "ssm_abc_debug_get")
        return out

    def __del__(self):
        try:
            if self._h:
                _lib.lib().ssm_abc_destroy(self._h)
        except Exception:
            pass
