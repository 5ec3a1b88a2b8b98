// generic_emul.cpp — DEBUGGING HARNESS, never shipped and never loaded by the product.
// Compiles the runtime-shape device routine of csrc/ab_generic.cuh for the HOST (one lane of one tile), so that the CPU-only test tier can
// check what the CUDA kernels instantiate: here the second entry of the sweep that continues a partial factorisation (ssm_abqr_continue).
#include <cmath>
using std::fma; using std::sqrt; using std::copysign;
#include "../../semiseparablematrices.jl_b200/csrc/ab_generic.cuh"

// arrays in the device layout of ONE tile: [np][fields][32], lane 0 is the system
extern "C" int generic_qr(int n, int l, int u, int r, long np, const double *AU, const double *V, double *RU, double *QT, int ncols, int kstart) {
    ssm::AbArgs a{AU, V, RU, QT, nullptr, nullptr, n, (int64_t)np, ncols};
    ssm::generic_ab_qr(a, l, u, r, false, true, 0, 0, kstart);
    return 0;
}

// growth fix-up of a partial factorisation (ssm_abqr_resize): RU / QT hold k0 reflectors formed at size n0, AU / V are the grown operand (size n)
extern "C" int generic_grow_fix(int n, int l, int u, int r, long np, const double *AU, const double *V, double *RU, double *QT, int n0, int k0) {
    ssm::AbArgs a{AU, V, RU, QT, nullptr, nullptr, n, (int64_t)np, k0};
    ssm::generic_ab_grow_fix(a, l, u, r, n0, k0, 0, 0);
    return 0;
}
