"""ctypes binding of libssmb200.so — the Python twin of the Julia `ccall` glue (julia/SemiseparableB200.jl).

The product path has NO CPU fallback: if the CUDA library is missing or cannot be loaded this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent.parent          # semiseparablematrices.jl_b200/
LIB_PATH = Path(os.environ.get("SSM_B200_LIB", _PKG / "lib" / "libssmb200.so"))

c_dp = C.POINTER(C.c_double)
c_vp = C.c_void_p
i64 = C.c_int64
u64 = C.c_uint64

SSM_DIST_DD, SSM_DIST_BC = 0, 1


class SSMError(RuntimeError):
    pass


class DimensionMismatch(SSMError, ValueError):
    """Mirror of Julia's DimensionMismatch (AlmostBandedMatrix.jl:310, semiseparableqr.jl:46)."""


_lib = None

_SIGS = {
    # name: (restype, [argtypes])
    "ssm_version": (C.c_int, []),
    "ssm_last_error": (C.c_char_p, []),
    "ssm_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "ssm_ctx_create_on_stream": (C.c_int, [C.c_int, c_vp, C.POINTER(c_vp)]),
    "ssm_ctx_destroy": (C.c_int, [c_vp]),
    "ssm_ctx_sync": (C.c_int, [c_vp]),
    "ssm_ctx_set_option": (C.c_int, [c_vp, C.c_char_p, C.c_int]),
    "ssm_ctx_get_option": (C.c_int, [c_vp, C.c_char_p, C.POINTER(C.c_int)]),
    "ssm_ctx_launch_count": (i64, [c_vp]),
    "ssm_ctx_stream": (c_vp, [c_vp]),
    "ssm_pool_trim": (i64, [C.c_int]),
    "ssm_pool_stats": (C.c_int, [C.c_int, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]),
    "ssm_vec_create": (C.c_int, [c_vp, C.c_int, i64, C.POINTER(c_vp)]),
    "ssm_vec_destroy": (C.c_int, [c_vp]),
    "ssm_vec_set": (C.c_int, [c_vp, c_vp, i64]),
    "ssm_vec_get": (C.c_int, [c_vp, c_vp, i64]),
    "ssm_vec_copy": (C.c_int, [c_vp, c_vp]),
    "ssm_vec_generate": (C.c_int, [c_vp, u64, C.c_int]),
    "ssm_ab_create": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, C.POINTER(c_vp)]),
    "ssm_ab_destroy": (C.c_int, [c_vp]),
    "ssm_ab_set": (C.c_int, [c_vp, c_vp, i64, c_vp, i64, c_vp, i64]),
    "ssm_ab_get": (C.c_int, [c_vp, c_vp, i64, c_vp, i64, c_vp, i64]),
    "ssm_ab_generate": (C.c_int, [c_vp, u64, C.c_int]),
    "ssm_ab_resize": (C.c_int, [c_vp, C.c_int, C.POINTER(c_vp)]),
    "ssm_ab_set_rows": (C.c_int, [c_vp, C.c_int, C.c_int, c_vp, i64, c_vp, i64, c_vp, i64]),
    "ssm_ab_vcat_bandwidths": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ssm_ab_set_vcat": (C.c_int, [c_vp, c_vp, i64, c_vp, i64, C.c_int, C.c_int]),
    "ssm_ab_qr": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "ssm_ab_qr_cols": (C.c_int, [c_vp, C.c_int, C.POINTER(c_vp)]),
    "ssm_abqr_continue": (C.c_int, [c_vp, c_vp, C.c_int]),
    "ssm_abqr_resize": (C.c_int, [c_vp, c_vp]),
    "ssm_abqr_destroy": (C.c_int, [c_vp]),
    "ssm_abqr_get": (C.c_int, [c_vp, c_vp, i64, c_vp, i64, c_vp, i64]),
    "ssm_abqr_set": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, c_vp, i64, c_vp, i64, c_vp, i64, c_vp, i64, C.c_int, C.POINTER(c_vp)]),
    "ssm_abqr_ldiv": (C.c_int, [c_vp, c_vp]),
    "ssm_abqr_solve": (C.c_int, [c_vp, c_vp, c_vp]),
    "ssm_mat_create": (C.c_int, [c_vp, C.c_int, C.c_int, i64, C.POINTER(c_vp)]),
    "ssm_mat_destroy": (C.c_int, [c_vp]),
    "ssm_mat_set": (C.c_int, [c_vp, c_vp, i64]),
    "ssm_mat_get": (C.c_int, [c_vp, c_vp, i64]),
    "ssm_abqr_ldiv_mat": (C.c_int, [c_vp, c_vp]),
    "ssm_abqr_lmul_qt": (C.c_int, [c_vp, c_vp]),
    "ssm_abqr_lmul_q": (C.c_int, [c_vp, c_vp]),
    "ssm_abqr_upper_ldiv": (C.c_int, [c_vp, c_vp]),
    "ssm_ab_solve": (C.c_int, [c_vp, c_vp, c_vp]),
    "ssm_ab_tri_ldiv": (C.c_int, [c_vp, c_vp, C.c_int, C.c_int]),
    "ssm_ab_residual": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ssm_ab_solve_host": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ssm_bps_create": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, C.POINTER(c_vp)]),
    "ssm_bps_destroy": (C.c_int, [c_vp]),
    "ssm_bps_set": (C.c_int, [c_vp, c_vp, i64, c_vp, c_vp, i64, c_vp, c_vp, i64]),
    "ssm_bps_generate": (C.c_int, [c_vp, u64]),
    "ssm_bps_qr": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "ssm_bpsqr_destroy": (C.c_int, [c_vp]),
    "ssm_bpsqr_get": (C.c_int, [c_vp, c_vp, i64, c_vp, c_vp, i64, c_vp, c_vp, i64, c_vp, i64]),
    "ssm_bpsqr_lmul_qt": (C.c_int, [c_vp, c_vp]),
    "ssm_bpsqr_upper_ldiv": (C.c_int, [c_vp, c_vp]),
    "ssm_bpsqr_ldiv": (C.c_int, [c_vp, c_vp]),
    "ssm_bpsqr_solve": (C.c_int, [c_vp, c_vp, c_vp]),
    "ssm_bps_get": (C.c_int, [c_vp, c_vp, i64, c_vp, c_vp, i64, c_vp, c_vp, i64]),
    "ssm_bpsqr_rq_mul": (C.c_int, [c_vp, C.POINTER(c_vp)]),
    "ssm_bps_mul": (C.c_int, [c_vp, c_vp, c_vp]),
    "ssm_bps_residual": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ssm_bps_upper_ldiv": (C.c_int, [c_vp, c_vp]),
    "ssm_shard_range": (C.c_int, [i64, C.c_int, C.c_int, C.POINTER(i64), C.POINTER(i64)]),
    "ssm_ab_solve_host_sharded": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int, i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "ssm_abc_create": (C.c_int, [c_vp, i64, C.c_int, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "ssm_abc_destroy": (C.c_int, [c_vp]),
    "ssm_abc_set": (C.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "ssm_abc_generate": (C.c_int, [c_vp, u64, C.c_int]),
    "ssm_abc_vec_generate": (C.c_int, [c_vp, c_vp, u64]),
    "ssm_abc_solve": (C.c_int, [c_vp, c_vp, c_vp, C.c_int]),
    "ssm_abc_residual": (C.c_int, [c_vp, c_vp, c_vp, C.POINTER(C.c_double)]),
    "ssm_abc_chunks": (C.c_int, [c_vp]),
    "ssm_abc_debug_get": (C.c_int, [c_vp, C.c_int, c_vp, i64]),
}


def exported_symbols():
    """Every entry point include/ssm_b200.h declares (checked by tests/test_abi.py)."""
    return sorted(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise SSMError(
                f"{LIB_PATH} not found — build it with `python __graft_entry__.py build` (nvcc, sm_100a). "
                "There is no CPU fallback.")
        _lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    msg = lib().ssm_last_error().decode("utf-8", "replace")
    if rc == -2:
        raise DimensionMismatch(msg or what)
    if rc < 0:
        raise SSMError(f"{what}: {msg} (code {rc})" if what else f"{msg} (code {rc})")
    raise SSMError(f"CUDA failure in {what}: {msg}")
