// FP64 throughput micro-benchmarks: the roofline denominators for the solver kernels.
// MEASURED_PEAKS.json (driver-written) has HBM and bf16 figures only; the inversion hot path is FP64-compute bound
// (SURVEY.md section 8d), so bench.py measures the FP64 FMA-pipe peak and the FP64 tensor (DMMA) peak here.
#include "common.cuh"

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double a, double b) {
  double c[8][2];
#pragma unroll
  for (int r = 0; r < 8; ++r) { c[r][0] = threadIdx.x; c[r][1] = r; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[r][0]), "+d"(c[r][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int r = 0; r < 8; ++r) s += c[r][0] + c[r][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int bdrt_peak_fp64(bdrt_ctx* ctx, double* dfma_tflops, double* dmma_tflops) {
  if (!ctx) return BDRT_E_NULL;
  if (!dfma_tflops || !dmma_tflops) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_peak_fp64: null pointer");
  const int ctas = ctx->sm_count * 4, thr = 256;
  int rc = bdrt_ws_reserve(ctx, (size_t)ctas * thr * sizeof(double));
  if (rc) return rc;
  double* out = (double*)ctx->ws;
  cudaEvent_t e0, e1;
  BDRT_CUDA(ctx, cudaEventCreate(&e0));
  BDRT_CUDA(ctx, cudaEventCreate(&e1));
  float ms;
  const int it_f = 4096, it_m = 2048;
  for (int rep = 0; rep < 2; ++rep) {  // first repetition warms up
    BDRT_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    dfma_peak_kernel<<<ctas, thr, 0, ctx->stream>>>(out, it_f, 0.999999, 1e-9);
    BDRT_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    BDRT_CUDA(ctx, cudaEventSynchronize(e1));
    BDRT_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    *dfma_tflops = 2.0 * 64.0 * it_f * (double)ctas * thr / (ms * 1e-3) / 1e12;
    BDRT_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
    dmma_peak_kernel<<<ctas, thr, 0, ctx->stream>>>(out, it_m, 0.5, 1e-9);
    BDRT_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
    BDRT_CUDA(ctx, cudaEventSynchronize(e1));
    BDRT_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
    // one m8n8k4 DMMA = 8*8*4 FMA = 512 flop per warp
    *dmma_tflops = 512.0 * 8.0 * it_m * (double)ctas * (thr / 32) / (ms * 1e-3) / 1e12;
  }
  ctx->launches += 4;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return BDRT_OK;
}
