"""ssm_b200 — B200-native batched structured QR (AlmostBanded / BandedPlusSemiseparable) behind the
reference's operator names.  Thin host layer over libssmb200.so (C ABI in include/ssm_b200.h)."""
from ._lib import DimensionMismatch, SSMError, SSM_DIST_BC, SSM_DIST_DD, exported_symbols, lib  # noqa: F401
from .almostbanded import (AlmostBandedMatrix, AlmostBandedQR, Context, LowerTriangular,  # noqa: F401
                           UnitLowerTriangular, UnitUpperTriangular, UpperTriangular, Vec, almostbandedrank,
                           almostbandwidths, bandpart, default_context, factorize, fillpart, ldiv, pool_stats, pool_trim, qr, solve,
                           solve_host)
