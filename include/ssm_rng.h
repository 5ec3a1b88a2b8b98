/* ssm_rng.h — counter-based synthetic-input generator shared by the CUDA library
 * (device side, kernel K11) and the CPU oracle (host side), so both see bit-identical
 * inputs without moving data.  Specification: SURVEY.md §8(d) "Synthetic inputs".
 *
 *   u(seed, s, a, e) = (splitmix64(splitmix64(seed + s) + (a << 40) + e) >> 11) * 2^-53  in [0,1)
 *     s = system index in the batch, a = array id, e = linear element index of that array
 *     in the REFERENCE memory layout (column-major band storage / column-major dense).
 *
 * Array ids.  AlmostBanded: 0 bands, 1 U, 2 V, 3 b.   BPS: 0 B, 1 U, 2 V, 3 W, 4 S, 5 b.
 */
#ifndef SSM_RNG_H
#define SSM_RNG_H
#include <stdint.h>

#if defined(__CUDACC__)
#define SSM_HD __host__ __device__ __forceinline__
#else
#define SSM_HD static inline
#endif

SSM_HD uint64_t ssm_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* per-system key; hoist out of element loops */
SSM_HD uint64_t ssm_syskey(uint64_t seed, uint64_t s) { return ssm_splitmix64(seed + s); }

SSM_HD double ssm_unif_key(uint64_t key, uint64_t a, uint64_t e) {
    uint64_t z = ssm_splitmix64(key + (a << 40) + e);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

SSM_HD double ssm_unif(uint64_t seed, uint64_t s, uint64_t a, uint64_t e) {
    return ssm_unif_key(ssm_syskey(seed, s), a, e);
}

/* distributions */
enum { SSM_DIST_DD = 0, /* strictly diagonally dominant band, small fill (throughput configs) */
       SSM_DIST_BC = 1  /* boundary-row / spectral form: U=[I_r;0], V dense decaying rows   */ };

/* ---- AlmostBanded element generators, all indices 0-based ------------------------------
 * bands element e = d + (l+u+1)*j  holds A[k,j] with k = j + d - u  (d = 0..l+u)          */
SSM_HD double ssm_ab_band_elem(uint64_t key, int dist, int n, int l, int u, int r, int d, int j) {
    int k = j + d - u;
    if (k < 0 || k >= n) return 0.0;                 /* outside the matrix: stored zero      */
    double x = ssm_unif_key(key, 0, (uint64_t)d + (uint64_t)(l + u + 1) * (uint64_t)j);
    (void)r; (void)dist;
    if (k == j) return (double)(l + u + 2) + x;      /* diagonal                             */
    return 2.0 * x - 1.0;
}
/* U element e = i + n*q */
SSM_HD double ssm_ab_U_elem(uint64_t key, int dist, int n, int r, int i, int q) {
    if (dist == SSM_DIST_BC) return (i == q) ? 1.0 : 0.0;
    (void)r;
    return 2.0 * ssm_unif_key(key, 1, (uint64_t)i + (uint64_t)n * (uint64_t)q) - 1.0;
}
/* V element e = q + r*j */
SSM_HD double ssm_ab_V_elem(uint64_t key, int dist, int n, int r, int q, int j) {
    if (dist == SSM_DIST_BC) {                       /* (+-1)^{(q-1) j} / (1+j)^2, 1-based q,j */
        double jj = (double)(j + 1);
        double sgn = (((uint64_t)q * (uint64_t)(j + 1)) & 1ull) ? -1.0 : 1.0;
        return sgn / ((1.0 + jj) * (1.0 + jj));
    }
    double x = 2.0 * ssm_unif_key(key, 2, (uint64_t)q + (uint64_t)r * (uint64_t)j) - 1.0;
    return x / ((double)r * (double)n);
}
SSM_HD double ssm_ab_b_elem(uint64_t key, int i) { return 2.0 * ssm_unif_key(key, 3, (uint64_t)i) - 1.0; }

/* ---- BPS element generators (dist dd only) ---------------------------------------------
 * B element e = d + (l+m+1)*j holds A[k,j], k = j + d - m                                  */
SSM_HD double ssm_bps_band_elem(uint64_t key, int n, int l, int m, int d, int j) {
    int k = j + d - m;
    if (k < 0 || k >= n) return 0.0;
    double x = ssm_unif_key(key, 0, (uint64_t)d + (uint64_t)(l + m + 1) * (uint64_t)j);
    if (k == j) return (double)(l + m + 2) + x;
    return 2.0 * x - 1.0;
}
/* generators: arr = 1 U, 2 V (n x r, scale 1/sqrt(r n)); 3 W, 4 S (n x p, scale 1/sqrt(p n)); e = i + n*q */
SSM_HD double ssm_bps_gen_elem(uint64_t key, int arr, int n, int cols, int i, int q, double scale) {
    (void)cols;
    return (2.0 * ssm_unif_key(key, (uint64_t)arr, (uint64_t)i + (uint64_t)n * (uint64_t)q) - 1.0) * scale;
}
SSM_HD double ssm_bps_b_elem(uint64_t key, int i) { return 2.0 * ssm_unif_key(key, 5, (uint64_t)i) - 1.0; }

#endif /* SSM_RNG_H */
