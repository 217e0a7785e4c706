// On-device posterior summaries of HMC draws: per-parameter mean and percentiles.
//
// Replaces the numpy reductions the reference applies to StanFit4Model draws: np.mean(axis=0) in
// Inverter._extract_parameter (bayes_drt/inversion.py:2514-2519) and np.percentile(axis=0) (default 'linear'
// interpolation) in coef_percentile / predict_Z / predict_sigma (:2560, :2702, :3096).
//
// Layout: draws [G, S, P] (G = spectra, S = merged post-warm-up draws of all chains, P = parameters, row-major).
// Every draw is read from HBM exactly once (S * P * 8 bytes per spectrum; 0.79 MB at S = 400, P = 246).  One CTA takes
// up to 32 consecutive parameters of one spectrum: rows are read coalesced (256 B contiguous) and transposed into
// shared memory, then every warp sorts columns with a bitonic network (S padded to a power of two with +inf) and reads
// the percentiles off the sorted column; the mean is accumulated in the original order of the draws.
// The sort runs in registers (warp_sort: shuffles between lanes, compare-exchanges inside a lane) for columns of up to
// 1024 draws.  Measured (B200, 4000 x 400 x 246 draws = 3.15 GB): 14.5 ms = 220 GB/s -- bound by the sort's
// compare-exchange instructions (61 % of the ncu samples), not by HBM (3.4 % of the measured copy bandwidth); the
// first version sorted in shared memory in 24.6 ms.  It is < 0.1 % of the sampling time that produces the draws.
#include "common.cuh"

#define SUM_COLS 32
#define SUM_THREADS 256

// Bitonic sort of 32 * E values held by one warp, E per lane, entirely in registers: lane owns the sorted positions
// lane * E .. lane * E + E - 1 at the end.  Stages that pair positions inside a lane are register compare-exchanges,
// stages that pair lanes are one shuffle per element.
template <int E>
__device__ __forceinline__ void warp_sort(double (&v)[E], int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= E) {  // partner position is in lane ^ (j / E), same register
        const int dl = j / E;
        const bool up = ((lane * E) & k) == 0;            // k >= 2E here: the direction is uniform over the lane
        const bool keep_min = ((lane & dl) == 0) == up;
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const double o = __shfl_xor_sync(0xffffffffu, v[r], dl);
          const bool take = keep_min ? (o < v[r]) : (o > v[r]);  // plain compares: NaN columns are handled by the caller
          v[r] = take ? o : v[r];
        }
      } else {  // both positions in this lane
#pragma unroll
        for (int r = 0; r < E; ++r) {
          const int l = r ^ j;
          if (l > r) {
            const bool up = (((lane * E) + r) & k) == 0;
            const double a = v[r], b = v[l];
            const bool sw = (a > b) == up;
            v[r] = sw ? b : a;
            v[l] = sw ? a : b;
          }
        }
      }
    }
  }
}

template <int E>
__device__ __forceinline__ void sort_column(double* col, int lane) {
  double v[E];
#pragma unroll
  for (int r = 0; r < E; ++r) v[r] = col[r * 32 + lane];  // any input order will do: conflict-free interleaved read
  warp_sort<E>(v, lane);
  __syncwarp();
#pragma unroll
  for (int r = 0; r < E; ++r) col[lane * E + r] = v[r];
  __syncwarp();
}

// E = S2 / 32 values per lane for the register sort (0: shared-memory sort for longer columns)
template <int E>
__global__ void __launch_bounds__(SUM_THREADS)
summarize_kernel(const double* __restrict__ draws, int G, int S, int P, int S2, int cols,
                 const double* __restrict__ probs, int nq, double* __restrict__ mean, double* __restrict__ quant) {
  extern __shared__ __align__(16) double sm[];  // [cols][S2 + 1], cols = parameters per tile (power of two <= 32)
  const int ld = S2 + 1;
  const int tiles = (P + cols - 1) / cols;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lc = threadIdx.x % cols, lr = threadIdx.x / cols, rows_per_pass = SUM_THREADS / cols;
  for (long long t = blockIdx.x; t < (long long)G * tiles; t += gridDim.x) {
    const int g = (int)(t / tiles), p0 = (int)(t % tiles) * cols;
    const int pc = lc + p0;
    const double* base = draws + (long long)g * S * P;
    __syncthreads();
    // coalesced load + transpose: consecutive threads read consecutive parameters of one draw
    for (int s = lr; s < S2; s += rows_per_pass)
      sm[lc * ld + s] = (s < S && pc < P) ? base[(long long)s * P + pc] : INFINITY;
    __syncthreads();
    // each warp handles columns warp, warp + 8, ...
    for (int c = warp; c < cols; c += SUM_THREADS / 32) {
      if (p0 + c >= P) continue;  // warp-uniform
      double* col = sm + c * ld;
      double acc = 0.0;
      for (int s = lane; s < S; s += 32) acc += col[s];
      acc = warp_sum(acc);
      // sort col[0 .. S2): in registers when the column fits (S2 <= 1024), else with a shared-memory bitonic network
      __syncwarp();
      if (E > 0) sort_column<(E > 0 ? E : 1)>(col, lane);
      else {
        for (int k = 2; k <= S2; k <<= 1) {
          for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < S2; i += 32) {
              const int l = i ^ j;
              if (l > i) {
                const double a = col[i], b = col[l];
                const bool up = (i & k) == 0;
                if ((a > b) == up) {
                  col[i] = b;
                  col[l] = a;
                }
              }
            }
            __syncwarp();
          }
        }
      }
      if (lane == 0 && mean) mean[(long long)g * P + p0 + c] = acc / S;
      // np.percentile(..., interpolation='linear'): virtual index q (S - 1)
      for (int q = lane; q < nq; q += 32) {
        const double h = probs[q] * (S - 1);
        int lo = (int)floor(h);
        lo = lo < 0 ? 0 : (lo > S - 1 ? S - 1 : lo);
        const int hi = lo + 1 > S - 1 ? S - 1 : lo + 1;
        const double fr = h - lo;
        // numpy's lerp: a + (b - a) * t for t < 0.5, b - (b - a) * (1 - t) otherwise
        const double a = col[lo], b = col[hi];
        double v = fr < 0.5 ? a + (b - a) * fr : b - (b - a) * (1.0 - fr);
        if (isnan(acc)) v = acc;  // a NaN among the draws: np.percentile returns NaN
        quant[((long long)q * G + g) * P + p0 + c] = v;
      }
      __syncwarp();
    }
  }
}

// draws [G, S, P] (device); probs_host [nq] in [0, 1] (host); mean [G, P] (device, may be NULL); quant [nq, G, P] (device,
// may be NULL when nq == 0).
extern "C" int bdrt_summarize(bdrt_ctx* ctx, const double* draws, int G, int S, int P, const double* probs_host, int nq,
                              double* mean, double* quant) {
  if (!ctx) return BDRT_E_NULL;
  if (!draws || (nq > 0 && (!probs_host || !quant))) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_summarize: null pointer");
  if (G < 0 || S < 1 || P < 1 || nq < 0 || nq > 64) BDRT_FAIL(ctx, BDRT_E_SIZE, "bdrt_summarize: bad sizes");
  for (int q = 0; q < nq; ++q)
    if (!(probs_host[q] >= 0.0 && probs_host[q] <= 1.0))
      BDRT_FAIL(ctx, BDRT_E_SIZE, "Percentiles must be in the range [0, 100]");  // numpy's message
  if (G == 0) return BDRT_OK;
  int S2 = 32;  // at least one value per lane
  while (S2 < S) S2 <<= 1;
  int cols = SUM_COLS;  // fewer parameters per tile for long chains, so that the columns still fit in shared memory
  while (cols > 1 && (size_t)cols * (S2 + 1) * sizeof(double) > (size_t)ctx->smem_optin / 2) cols >>= 1;
  const size_t smem = (size_t)cols * (S2 + 1) * sizeof(double);
  if (smem > (size_t)ctx->smem_optin)
    BDRT_FAIL(ctx, BDRT_E_SMEM, "bdrt_summarize: %d draws per parameter exceed the shared-memory sort (max %d)", S,
              (int)(ctx->smem_optin / sizeof(double)) / 2);
  int rc = bdrt_ws_reserve(ctx, 64 * sizeof(double));
  if (rc) return rc;
  double* dprobs = (double*)ctx->ws;
  if (nq) BDRT_CUDA(ctx, cudaMemcpyAsync(dprobs, probs_host, nq * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
#define SUM_LAUNCH(EE)                                                                                          \
  do {                                                                                                         \
    BDRT_CUDA(ctx, cudaFuncSetAttribute(summarize_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    summarize_kernel<EE><<<(int)grid, SUM_THREADS, smem, ctx->stream>>>(draws, G, S, P, S2, cols, dprobs, nq, mean, quant); \
  } while (0)
  const long long tiles = (long long)G * ((P + cols - 1) / cols);
  int per_sm = (int)((size_t)ctx->smem_per_sm / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > tiles) grid = tiles;
  switch (S2) {
    case 32: SUM_LAUNCH(1); break;
    case 64: SUM_LAUNCH(2); break;
    case 128: SUM_LAUNCH(4); break;
    case 256: SUM_LAUNCH(8); break;
    case 512: SUM_LAUNCH(16); break;
    case 1024: SUM_LAUNCH(32); break;
    default: SUM_LAUNCH(0); break;
  }
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Sampler diagnostics on the device: split R-hat and rank-normalised split-chain bulk ESS (Vehtari et al. 2021) of every
// column of draws [G, chains * n, P] (chain-major along the draw axis) -- what pystan prints after
// StanModel.sampling (bayes_drt/inversion.py:1218-1221) and what the benchmark's ESS/s metric is made of.
// One warp per (spectrum, parameter) column: the column and its normal scores live in shared memory; ranks by counting
// (average rank for ties, like scipy.stats.rankdata), autocovariances lag-parallel over the lanes, Geyer's initial
// monotone sequence by lane 0.  Same estimator as oracle/nuts.py: ess_bulk (the tests compare the two).
// ---------------------------------------------------------------------------------------------------------------------
#define DIAG_WARPS 4

__global__ void __launch_bounds__(DIAG_WARPS * 32)
diag_kernel(const double* __restrict__ draws, int G, int chains, int n, int P, double* __restrict__ rhat,
            double* __restrict__ ess) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = n / 2, m = 2 * chains, S = m * h;  // split chains; an odd n drops its last draw
  double* x = sm + (size_t)warp * 2 * S;  // values, later rho[t]
  double* z = x + S;                       // normal scores, centred per split chain
  const long long ncol = (long long)G * P;
  for (long long col = (long long)blockIdx.x * DIAG_WARPS + warp; col < ncol; col += (long long)gridDim.x * DIAG_WARPS) {
    const int g = (int)(col / P), p = (int)(col - (long long)g * P);
    const double* base = draws + (long long)g * chains * n * P + p;
    __syncwarp();
    for (int s = lane; s < S; s += 32) {  // split chain k = s / h: chain k / 2, half k & 1
      const int k = s / h, i = s - k * h;
      x[s] = base[(long long)((k >> 1) * n + (k & 1) * h + i) * P];
    }
    __syncwarp();
    // ---- classic split R-hat on the raw values
    double sumv = 0.0, summ = 0.0, summ2 = 0.0;
    for (int k = 0; k < m; ++k) {
      double a = 0.0;
      for (int i = lane; i < h; i += 32) a += x[k * h + i];
      const double mu = warp_sum(a) / h;
      double v = 0.0;
      for (int i = lane; i < h; i += 32) { const double d = x[k * h + i] - mu; v = fma(d, d, v); }
      sumv += warp_sum(v) / (h - 1.0);
      summ += mu;
      summ2 = fma(mu, mu, summ2);
    }
    {
      const double W = sumv / m, mm = summ / m;
      const double Bn = (summ2 - m * mm * mm) / (m - 1.0);  // variance of the split-chain means
      if (lane == 0 && rhat) rhat[col] = sqrt(((h - 1.0) / h * W + Bn) / W);
    }
    // ---- rank-normalise
    for (int s = lane; s < S; s += 32) {
      const double v = x[s];
      int less = 0, eq = 0;
      for (int j = 0; j < S; ++j) {
        const double o = x[j];
        less += o < v;
        eq += o == v;
      }
      const double r = less + 0.5 * (eq + 1);
      z[s] = normcdfinv((r - 0.375) / (S + 0.25));
    }
    __syncwarp();
    sumv = 0.0; summ = 0.0; summ2 = 0.0;
    for (int k = 0; k < m; ++k) {  // centre every split chain
      double a = 0.0;
      for (int i = lane; i < h; i += 32) a += z[k * h + i];
      const double mu = warp_sum(a) / h;
      double v = 0.0;
      for (int i = lane; i < h; i += 32) {
        const double d = z[k * h + i] - mu;
        z[k * h + i] = d;
        v = fma(d, d, v);
      }
      sumv += warp_sum(v) / (h - 1.0);  // acov_k[0] h / (h - 1)
      summ += mu;
      summ2 = fma(mu, mu, summ2);
    }
    __syncwarp();
    const double Wz = sumv / m, mz = summ / m;
    const double var_plus = Wz * (h - 1.0) / h + (summ2 - m * mz * mz) / (m - 1.0);
    // rho_t = 1 - (W - mean_k acov_k[t]) / var_plus, lag-parallel
    for (int t = lane; t < h; t += 32) {
      double a = 0.0;
      for (int k = 0; k < m; ++k) {
        const double* zk = z + k * h;
        for (int i = 0; i + t < h; ++i) a = fma(zk[i], zk[i + t], a);
      }
      x[t] = t == 0 ? 1.0 : 1.0 - (Wz - a / (h * (double)m)) / var_plus;
    }
    __syncwarp();
    if (lane == 0 && ess) {
      double tau = -1.0, prev = INFINITY;
      for (int t = 0; t + 1 < h; t += 2) {
        double pair = x[t] + x[t + 1];
        if (pair < 0.0) break;
        pair = fmin(pair, prev);
        prev = pair;
        tau += 2.0 * pair;
      }
      tau = fmax(tau, 1.0 / log10((double)S));
      ess[col] = S / tau;
    }
  }
}

// draws [G, chains * n, P] (device, draw index = chain * n + i); rhat [G, P], ess_bulk [G, P] (device; either may be
// NULL).  n >= 4, chains >= 1.
extern "C" int bdrt_diagnostics(bdrt_ctx* ctx, const double* draws, int G, int chains, int n, int P, double* rhat,
                                double* ess_bulk) {
  if (!ctx) return BDRT_E_NULL;
  if (!draws) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_diagnostics: null pointer");
  if (G < 0 || chains < 1 || n < 4 || P < 1) BDRT_FAIL(ctx, BDRT_E_SIZE, "bdrt_diagnostics: bad sizes");
  if (G == 0 || (!rhat && !ess_bulk)) return BDRT_OK;
  const size_t S = (size_t)2 * chains * (n / 2);
  const size_t smem = DIAG_WARPS * 2 * S * sizeof(double);
  if (smem > (size_t)ctx->smem_optin)
    BDRT_FAIL(ctx, BDRT_E_SMEM, "bdrt_diagnostics: %zu draws per parameter exceed the shared-memory buffers (max %zu)", S,
              (size_t)ctx->smem_optin / (DIAG_WARPS * 2 * sizeof(double)));
  BDRT_CUDA(ctx, cudaFuncSetAttribute(diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ncol = (long long)G * P;
  int per_sm = (int)((size_t)ctx->smem_per_sm / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  long long grid = (long long)ctx->sm_count * per_sm;
  if (grid > (ncol + DIAG_WARPS - 1) / DIAG_WARPS) grid = (ncol + DIAG_WARPS - 1) / DIAG_WARPS;
  diag_kernel<<<(int)grid, DIAG_WARPS * 32, smem, ctx->stream>>>(draws, G, chains, n, P, rhat, ess_bulk);
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}
