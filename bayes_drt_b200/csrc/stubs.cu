// Entry points declared in include/bdrt.h whose kernels are not in this build yet: they fail loudly.
#include "common.cuh"
extern "C" void bdrt_newton_default_opts(bdrt_newton_opts* o) { if (o) { o->max_iter = 40; o->gtol = 1e-9; o->fd_step = 1e-6; } }
extern "C" int bdrt_map_newton(bdrt_ctx* ctx, const bdrt_series_data*, const bdrt_newton_opts*, double*, double*, double*, int*, int*) {
  if (!ctx) return BDRT_E_NULL; BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED, "bdrt_map_newton: not implemented in this build"); }
extern "C" int bdrt_qp_bound(bdrt_ctx* ctx, const double*, const double*, const double*, int, int, double*, double*, int*) {
  if (!ctx) return BDRT_E_NULL; BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED, "bdrt_qp_bound: not implemented in this build"); }
extern "C" void bdrt_ridge_default_opts(bdrt_ridge_opts* o) { if (o) { memset(o, 0, sizeof(*o)); o->nonneg = 1; o->max_iter = 20; o->xtol = 1e-3; o->hl_beta = 2.5; o->lambda_0 = 1e-2; o->reg_ord[2] = 1.0; o->epsilon = 1.0; o->fit_inductance = 1; } }
extern "C" int bdrt_ridge_fit(bdrt_ctx* ctx, const bdrt_ridge_opts*, const double*, const double*, int, const double*, const double*, const double*, const double*, int, int, int, double*, double*, int*, int*) {
  if (!ctx) return BDRT_E_NULL; BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED, "bdrt_ridge_fit: not implemented in this build"); }
