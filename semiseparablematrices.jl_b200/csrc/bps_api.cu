// bps_api.cu — BandedPlusSemiseparable entry points (placeholder until the BPS kernels land in this round)
#include "ssm_internal.cuh"
using namespace ssm;
#define NOTYET { set_error("%s: BPS path not built yet", __func__); return SSM_E_UNSUPPORTED; }
extern "C" {
int ssm_bps_create(ssm_ctx *, int, int, int, int, int, int64_t, ssm_bps **) NOTYET
int ssm_bps_destroy(ssm_bps *) { return 0; }
int ssm_bps_set(ssm_bps *, const double *, int64_t, const double *, const double *, int64_t, const double *, const double *, int64_t) NOTYET
int ssm_bps_generate(ssm_bps *, uint64_t) NOTYET
int ssm_bps_qr(const ssm_bps *, ssm_bpsqr **) NOTYET
int ssm_bpsqr_destroy(ssm_bpsqr *) { return 0; }
int ssm_bpsqr_get(const ssm_bpsqr *, double *, int64_t, double *, double *, int64_t, double *, double *, int64_t, double *, int64_t) NOTYET
int ssm_bpsqr_lmul_qt(const ssm_bpsqr *, ssm_vec *) NOTYET
int ssm_bpsqr_upper_ldiv(const ssm_bpsqr *, ssm_vec *) NOTYET
int ssm_bpsqr_ldiv(const ssm_bpsqr *, ssm_vec *) NOTYET
int ssm_bps_mul(const ssm_bps *, const ssm_vec *, ssm_vec *) NOTYET
int ssm_bps_residual(const ssm_bps *, const ssm_vec *, const ssm_vec *, double *) NOTYET
int ssm_bps_upper_ldiv(const ssm_bps *, ssm_vec *) NOTYET
}
