// Entry points declared in include/bdrt.h whose kernels are not in this build yet: they fail loudly.
#include "common.cuh"
extern "C" void bdrt_newton_default_opts(bdrt_newton_opts* o) { if (o) { o->max_iter = 40; o->gtol = 1e-9; o->fd_step = 1e-6; } }
extern "C" int bdrt_map_newton(bdrt_ctx* ctx, const bdrt_series_data*, const bdrt_newton_opts*, double*, double*, double*, int*, int*) {
  if (!ctx) return BDRT_E_NULL; BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED, "bdrt_map_newton: not implemented in this build"); }
