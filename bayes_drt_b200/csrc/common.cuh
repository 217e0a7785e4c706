// Shared host/device helpers of libbdrt (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bdrt.h"

struct bdrt_ctx {
  int device;
  cudaStream_t stream;
  char err[512];
  long long launches;
  // scratch workspace (grown lazily; the library never allocates result memory)
  void* ws;
  size_t ws_bytes;
  int sm_count;
  int smem_optin;  // max dynamic shared memory per block (opt-in), bytes
  int smem_per_sm; // shared memory per SM, bytes
  unsigned long long* dbg_clk;  // profiling builds (-DBDRT_PHASE_CLOCKS): [16] device counters
};

#define BDRT_FAIL(ctx, code, ...)                             \
  do {                                                        \
    snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);    \
    return (code);                                            \
  } while (0)

#define BDRT_CUDA(ctx, call)                                                                            \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) {                                                                            \
      snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call,              \
               cudaGetErrorString(e_));                                                                 \
      return (int)e_;                                                                                   \
    }                                                                                                   \
  } while (0)

// grow-only scratch workspace
static inline int bdrt_ws_reserve(bdrt_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return 0;
  if (ctx->ws) {
    BDRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    BDRT_CUDA(ctx, cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  BDRT_CUDA(ctx, cudaMalloc(&ctx->ws, bytes));
  ctx->ws_bytes = bytes;
  return 0;
}

#ifdef __CUDACC__
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;  // xor butterfly: bitwise identical in every lane
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_and(int v) { return __all_sync(0xffffffffu, v); }
__device__ __forceinline__ double bcast(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif
