// Host side of lbfgs: the kernel template lives in lbfgs_kernel.cuh and is instantiated in lbfgs_t{0,1,2}.cu.
#include "lbfgs_kernel.cuh"

extern template __global__ void lbfgs_kernel<0, 0, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<0, 1, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<0, 2, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<0, 3, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 0, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 0, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 1, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 1, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 2, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 2, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 3, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<1, 3, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 0, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 0, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 1, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 1, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 2, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 2, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 3, 0>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);
extern template __global__ void lbfgs_kernel<2, 3, 1>(BdrtModel, bdrt_lbfgs_opts, double*, double*, int*, int*, int*, int*, double*, double*, int, int);

extern "C" void bdrt_lbfgs_default_opts(bdrt_lbfgs_opts* o) {
  if (!o) return;
  o->max_iter = 2000;  // Stan's own default; the reference passes iter=50000 (inversion.py:1076)
  o->history = 5;
  o->init_alpha = 1e-3;
  o->tol_obj = 1e-12;
  o->tol_rel_obj = 1e4;
  o->tol_grad = 1e-8;
  o->tol_rel_grad = 1e7;
  o->tol_param = 1e-8;
}

extern "C" int bdrt_map_lbfgs(bdrt_ctx* ctx, const bdrt_series_data* data, const bdrt_lbfgs_opts* opts, double* u,
                              double* lp, int* iters, int* n_eval, int* status) {
  if (!ctx) return BDRT_E_NULL;
  if (!u || !opts) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_map_lbfgs: null pointer");
  if (opts->history < 1 || opts->history > MAXHIST) BDRT_FAIL(ctx, BDRT_E_SIZE, "history must be in 1..%d", MAXHIST);
  if (opts->max_iter < 1) BDRT_FAIL(ctx, BDRT_E_SIZE, "max_iter must be >= 1");
  BdrtModel m;
  {
    // sizes first (model_prepare needs the scratch size)
    if (!data) BDRT_FAIL(ctx, BDRT_E_NULL, "null data");
  }
  const int D = bdrt_num_params(data);
  const int Dpad = (D + 223) / 224 * 224;  // zero-padded work vectors: the sweeps of the kernel have no bounds checks
  const int max_ctas = 2 * ctx->sm_count;  // scratch is sized for the two-CTAs-per-SM (Toeplitz) plan
  const int groups = data->per_spectrum_grid ? data->B : (data->B + NSLOT - 1) / NSLOT;
  const int grid_max = data->B == 0 ? 0 : (groups < max_ctas ? groups : max_ctas);
  int grid = grid_max;
  const size_t hist_bytes = (size_t)grid_max * NSLOT * 2 * opts->history * Dpad * sizeof(double);
  const size_t gvec_bytes = (size_t)grid_max * NSLOT * 5 * Dpad * sizeof(double);
  void* extra = nullptr;
  int rc = bdrt_model_prepare(ctx, data, &m, 256 + hist_bytes + gvec_bytes, &extra, 3);
  if (rc) return rc;
  if (data->B == 0) return BDRT_OK;
  if (m.wmode && grid > (data->B + NSLOT - 1) / NSLOT) grid = (data->B + NSLOT - 1) / NSLOT;  // 8 spectra per CTA
  int* queue = (int*)extra;
  double* hist = (double*)((char*)extra + 256);
  double* gvec = (double*)((char*)extra + 256 + hist_bytes);
  BDRT_CUDA(ctx, cudaMemsetAsync(queue, 0, 256, ctx->stream));
  // how many of the 5 work vectors per slot fit in shared memory
  const BdrtPlan pl = bdrt_plan(ctx, m, Dpad, 5, 2);
  if (grid > ctx->sm_count * pl.ctas_per_sm) grid = ctx->sm_count * pl.ctas_per_sm;
  BDRT_LAUNCH(ctx, m, lbfgs_kernel, grid, pl.smem, m, *opts, u, lp, iters, n_eval, status, queue, hist, gvec, pl.nvec,
              Dpad);
  return BDRT_OK;
}
