// fastmath.cuh — fp64 reciprocal / reciprocal square root without the IEEE slow paths (device only).
#pragma once

namespace ssm {

// fp64 reciprocal / reciprocal square root without the IEEE slow paths: hardware seed (MUFU, ~2^-22) + two Newton steps
// (relative error ~1 ulp).  The chunked path returns a solution, not reference-identical factors, so 1-2 ulp is immaterial.
__device__ __forceinline__ double fast_rcp(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a, y, 1.0);
    y = fma(y, e, y);
    e = fma(-a, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
}

// sqrt(a) to ~1 ulp from the refined reciprocal square root (a > 0 and finite; a == 0 gives NaN, callers mask it)
__device__ __forceinline__ double fast_sqrt_from_rsqrt(double a, double rs) {
    const double nrm = a * rs;
    return fma(fma(-nrm, nrm, a), 0.5 * rs, nrm);
}

}  // namespace ssm
