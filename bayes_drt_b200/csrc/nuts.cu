// Host side of nuts: the kernel template lives in nuts_kernel.cuh and is instantiated in nuts_t{0,1,2}.cu.
#include "nuts_kernel.cuh"

extern template __global__ void nuts_kernel<0, 0, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<0, 1, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<0, 2, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<0, 3, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 0, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 0, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 1, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 1, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 2, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 2, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 3, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<1, 3, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 0, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 0, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 1, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 1, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 2, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 2, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 3, 0>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);
extern template __global__ void nuts_kernel<2, 3, 1>(BdrtModel, bdrt_nuts_opts, const double*, double*, double*, long long*, int*, int*, double*, int*, double*, double*, int, int);

extern "C" void bdrt_nuts_default_opts(bdrt_nuts_opts* o) {
  if (!o) return;
  o->chains = 2;  // inversion.py:1079
  o->warmup = 200;
  o->samples = 200;
  o->max_treedepth = 10;
  o->adapt_delta = 0.9;  // inversion.py:1221
  o->adapt_t0 = 10;
  o->adapt_gamma = 0.05;
  o->adapt_kappa = 0.75;
  o->seed = 1234;  // inversion.py:1075
  o->spectrum_offset = 0;
  o->spectrum_ids = nullptr;
}

extern "C" int bdrt_nuts(bdrt_ctx* ctx, const bdrt_series_data* data, const bdrt_nuts_opts* opts, const double* u0,
                         double* draws, double* stepsize, long long* n_leapfrog, int* n_divergent, int* n_maxdepth,
                         double* accept) {
  if (!ctx) return BDRT_E_NULL;
  if (!data || !opts || !u0) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_nuts: null pointer");
  if (opts->chains < 1 || opts->warmup < 0 || opts->samples < 0) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad chains/warmup/samples");
  if (opts->max_treedepth < 1 || opts->max_treedepth > MAXDEPTH)
    BDRT_FAIL(ctx, BDRT_E_SIZE, "max_treedepth must be in 1..%d", MAXDEPTH);
  const int D = bdrt_num_params(data);
  const int Dpad = (D + 223) / 224 * 224;  // zero-padded work vectors: the hot sweeps of the kernel have no bounds checks
  const long long n_work = (long long)data->B * opts->chains;
  if (n_work > 2000000000LL) BDRT_FAIL(ctx, BDRT_E_SIZE, "too many chains in one call");
  const long long groups = data->per_spectrum_grid ? data->B : (n_work + NSLOT - 1) / NSLOT;
  const int max_ctas = 2 * ctx->sm_count;  // scratch is sized for the two-CTAs-per-SM (Toeplitz) plan
  const int grid_max = n_work == 0 ? 0 : (int)(groups < max_ctas ? groups : max_ctas);
  int grid = grid_max;
  const size_t gvec_bytes = (size_t)grid_max * NSLOT * NV * Dpad * sizeof(double);
  const size_t ck_bytes = (size_t)grid_max * NSLOT * 2 * MAXDEPTH * Dpad * sizeof(double);
  BdrtModel m;
  void* extra = nullptr;
  int rc = bdrt_model_prepare(ctx, data, &m, 256 + gvec_bytes + ck_bytes, &extra, 2);
  if (rc) return rc;
  if (n_work == 0) return BDRT_OK;
  if (m.wmode && grid > (n_work + NSLOT - 1) / NSLOT) grid = (int)((n_work + NSLOT - 1) / NSLOT);  // 8 chains per CTA
  int* queue = (int*)extra;
  double* gvec = (double*)((char*)extra + 256);
  double* ckpt = (double*)((char*)extra + 256 + gvec_bytes);
  BDRT_CUDA(ctx, cudaMemsetAsync(queue, 0, 256, ctx->stream));
  const BdrtPlan pl = bdrt_plan(ctx, m, Dpad, NV, 2);
  if (grid > ctx->sm_count * pl.ctas_per_sm) grid = ctx->sm_count * pl.ctas_per_sm;
  BDRT_LAUNCH_SOLVER(ctx, m, nuts_kernel, grid, pl.smem, m, *opts, u0, draws, stepsize, n_leapfrog, n_divergent, n_maxdepth,
              accept, queue, gvec, ckpt, pl.nvec, Dpad);
  return BDRT_OK;
}
