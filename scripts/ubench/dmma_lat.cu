// Micro-benchmark: latency / throughput of mma.sync.m8n8k4.f64 (DMMA) and DFMA per warp on sm_100a, as a function of
// the number of independent accumulation chains per warp and of warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/dmma_lat scripts/ubench/dmma_lat.cu && gpurun_out/dmma_lat
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int CH>
__global__ void k_dmma(double* out, long long* clk, int iters, double a, double b) {
  double c[CH][2];
#pragma unroll
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int CH>
__global__ void k_dfma(double* out, long long* clk, int iters, double a, double b) {
  double c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) c[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

// dependent shared-memory loads (pointer chase) and shuffles
__global__ void k_lds(double* out, long long* clk, int iters) {
  __shared__ double buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = (double)((i * 7 + 3) & 1023);
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) idx = (int)buf[idx];
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = idx;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void k_shfl(double* out, long long* clk, int iters) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v += __shfl_xor_sync(0xffffffffu, v, 1 + (it & 15));
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void k_mufu(double* out, long long* clk, int iters) {  // dependent exp / log / rcp chains in FP64
  double v = 1.0 + threadIdx.x * 1e-3, w = v, r = v;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = exp(v * 1e-3);
  long long t1 = clock64();
  for (int it = 0; it < iters; ++it) w = log(w + 2.0);
  long long t2 = clock64();
  for (int it = 0; it < iters; ++it) r = __drcp_rn(r + 1.5);
  long long t3 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + w + r;
  if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; }
}

template <int CH>
void run_dmma(double* out, long long* clk, int warps, int blocks_per_sm, int nsm) {
  const int iters = 2000;
  k_dmma<CH><<<nsm * blocks_per_sm, warps * 32>>>(out, clk, iters, 1.0000001, 0.5);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("DMMA chains=%d warps/CTA=%2d CTAs/SM=%d : %.1f clk per DMMA per warp, %.2f clk per DMMA per SM\n", CH, warps,
         blocks_per_sm, (double)h / (iters * CH), (double)h / (iters * CH) / (warps * blocks_per_sm));
}
template <int CH>
void run_dfma(double* out, long long* clk, int warps, int nsm) {
  const int iters = 4000;
  k_dfma<CH><<<nsm, warps * 32>>>(out, clk, iters, 1.0000001, 0.5);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost);
  printf("DFMA chains=%d warps/CTA=%2d : %.2f clk per DFMA per warp, %.3f clk per warp-DFMA per SM\n", CH, warps,
         (double)h / (iters * CH), (double)h / (iters * CH) / warps);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int nsm = p.multiProcessorCount;
  double* out; long long* clk;
  cudaMalloc(&out, sizeof(double) * nsm * 4 * 1024); cudaMalloc(&clk, 64);
  for (int w : {1, 4, 8, 16}) {
    run_dmma<1>(out, clk, w, 1, nsm); run_dmma<2>(out, clk, w, 1, nsm); run_dmma<4>(out, clk, w, 1, nsm);
    run_dmma<6>(out, clk, w, 1, nsm); run_dmma<8>(out, clk, w, 1, nsm);
  }
  for (int w : {1, 4, 8, 16}) { run_dfma<1>(out, clk, w, nsm); run_dfma<4>(out, clk, w, nsm); run_dfma<8>(out, clk, w, nsm); }
  long long h[3];
  k_lds<<<nsm, 32>>>(out, clk, 4000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("dependent LDS.64 + cvt: %.1f clk\n", (double)h[0] / 4000);
  k_shfl<<<nsm, 32>>>(out, clk, 4000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 8, cudaMemcpyDeviceToHost);
  printf("dependent 64-bit shfl + dadd: %.1f clk\n", (double)h[0] / 4000);
  for (int w : {1, 16}) {
    k_mufu<<<nsm, 32 * w>>>(out, clk, 1000); cudaDeviceSynchronize(); cudaMemcpy(h, clk, 24, cudaMemcpyDeviceToHost);
    printf("warps=%d dependent exp: %.0f clk, log: %.0f clk, drcp: %.0f clk\n", w, h[0] / 1000.0, h[1] / 1000.0, h[2] / 1000.0);
  }
  return 0;
}
