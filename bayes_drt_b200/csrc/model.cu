// Context management, model preparation, and the log_prob / constrain entry points.
#include <stdlib.h>

#include "engine.cuh"

extern "C" int bdrt_version(void) { return BDRT_VERSION; }

extern "C" int bdrt_ctx_create(int device, void* stream, bdrt_ctx** out) {
  if (!out) return BDRT_E_NULL;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  bdrt_ctx* c = new bdrt_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = (int)prop.sharedMemPerBlockOptin;
  c->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
  *out = c;
  return BDRT_OK;
}

extern "C" int bdrt_ctx_destroy(bdrt_ctx* ctx) {
  if (!ctx) return BDRT_E_NULL;
  cudaSetDevice(ctx->device);
  if (ctx->ws) cudaFree(ctx->ws);
  delete ctx;
  return BDRT_OK;
}

#ifdef BDRT_PHASE_CLOCKS
// profiling builds only (not part of include/bdrt.h): read and reset the per-phase clock sums
extern "C" int bdrt_debug_phase_clocks(bdrt_ctx* ctx, unsigned long long* out16_host) {
  if (!ctx->dbg_clk) return BDRT_E_NULL;
  BDRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  BDRT_CUDA(ctx, cudaMemcpy(out16_host, ctx->dbg_clk, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  BDRT_CUDA(ctx, cudaMemset(ctx->dbg_clk, 0, 16 * sizeof(unsigned long long)));
  return BDRT_OK;
}
#endif

extern "C" const char* bdrt_last_error(const bdrt_ctx* ctx) { return ctx ? ctx->err : "null context"; }
extern "C" long long bdrt_launch_count(const bdrt_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int bdrt_num_params(const bdrt_series_data* d) {
  if (!d) return BDRT_E_NULL;
  if ((d->model & 15) == BDRT_MODEL_SERIES_PARALLEL) return 2 * (d->K + d->Kp) + 12;
  if ((d->model & 15) == BDRT_MODEL_SERIES_2PARALLEL) return 2 * (d->K + d->Kp + d->Kp2) + 15;
  return 2 * d->K + 9 + ((d->model & BDRT_MODEL_OUTLIERS) ? 2 * d->Nf : 0);
}
extern "C" int bdrt_num_outputs(const bdrt_series_data* d) {
  if (!d) return BDRT_E_NULL;
  if ((d->model & 15) == BDRT_MODEL_SERIES_PARALLEL) return d->K + d->Kp + 6 + 2 * d->Nf;
  if ((d->model & 15) == BDRT_MODEL_SERIES_2PARALLEL) return d->K + d->Kp + d->Kp2 + 6 + 2 * d->Nf;
  return d->K + 6 + 2 * d->Nf + ((d->model & BDRT_MODEL_OUTLIERS) ? d->Nf : 0);
}

// ---------------------------------------------------------------------------------------------------------------------
// Band extraction of the dense penalty matrices L0/L1/L2 (K x K each): at the reference's default epsilon the entries
// decay like exp(-(n-m)^2), so only |n-m| <= bw are kept (everything dropped is < 1e-18 of the largest entry, i.e.
// invisible in FP64 next to the reference's dense L @ x).  info[0] = bw, info[1] = 1 when every L_j is Toeplitz.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void band_prep_kernel(const double* L, int K, double* Lb, int* info) {
  const int j = blockIdx.x;
  const double* Lj = L + (long long)j * K * K;
  __shared__ double s_red[32];
  __shared__ double s_max;
  double mx = 0.0;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) mx = fmax(mx, fabs(Lj[i]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) v = fmax(v, s_red[w]);
    s_max = v;
  }
  __syncthreads();
  mx = s_max;
  int bw = 0, toep = 1;
  const int r = K / 2;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    const int n = i / K, c = i - n * K;
    const int d = c - n;
    const double v = Lj[i];
    if (fabs(v) > 1e-18 * mx) bw = max(bw, abs(d));
    if (abs(d) <= MAXBW) {
      Lb[((long long)j * K + n) * LBW + d + MAXBW] = v;
      const int rc = r + d;
      if (rc >= 0 && rc < K) {
        if (fabs(v - Lj[r * K + rc]) > 1e-13 * mx) toep = 0;
      } else if (fabs(v) > 1e-18 * mx) {
        toep = 0;
      }
    }
  }
  // zero the out-of-range band slots
  for (int i = threadIdx.x; i < K * LBW; i += blockDim.x) {
    const int n = i / LBW, d = i - n * LBW - MAXBW;
    if (n + d < 0 || n + d >= K) Lb[((long long)j * K + n) * LBW + d + MAXBW] = 0.0;
  }
  if (bw) atomicMax(&info[0], bw);
  if (!toep) atomicAnd(&info[1], 0);
}

// Is the stacked kernel matrix [A_re; A_im] Toeplitz in each part (shared log-uniform grid with the measurement
// frequencies on the basis grid -- the case the reference special-cases too, matrices.py:145-242)?  info[2] &= yes.
__global__ void toep_check_kernel(const double* A0, long long stride, int Nf, int K, int* info) {
  const double* A = A0 + blockIdx.x * stride;  // one block per grid
  __shared__ double s_red[32];
  __shared__ double s_max;
  double mx = 0.0;
  for (int i = threadIdx.x; i < 2 * Nf * K; i += blockDim.x) mx = fmax(mx, fabs(A[i]));
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int w = 0; w < (blockDim.x + 31) / 32; ++w) v = fmax(v, s_red[w]);
    s_max = v;
  }
  __syncthreads();
  // entries of a diagonal must agree to 1e-12 relative (the table then reproduces every entry to that accuracy, far
  // inside the 1e-10 parity tolerance of the matrices themselves); entries below 1e-16 of the largest are noise
  const double floor_ = 1e-16 * s_max;
  int ok = 1;
  for (int i = threadIdx.x; i < 2 * Nf * K; i += blockDim.x) {
    const int r = i / K, c = i - r * K;
    const int rl = r >= Nf ? r - Nf : r;
    if (rl + 1 < Nf && c + 1 < K) {
      const double a = A[i], b = A[i + K + 1];
      if (!(fabs(a - b) <= 1e-12 * fmax(fabs(a), fabs(b)) + floor_)) ok = 0;
    }
  }
  if (!ok) atomicAnd(&info[2], 0);
}

static int bdrt_check_data(bdrt_ctx* ctx, const bdrt_series_data* d, int* nd_out) {
  if (!d) BDRT_FAIL(ctx, BDRT_E_NULL, "null data");
  if (!d->A || !d->Z || !d->freq || !d->L) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_series_data: null matrix pointer");
  const int base = d->model & 15;
  if (base != BDRT_MODEL_SERIES && base != BDRT_MODEL_SERIES_PARALLEL && base != BDRT_MODEL_PARALLEL &&
      base != BDRT_MODEL_SERIES_2PARALLEL)
    BDRT_FAIL(ctx, BDRT_E_MODEL, "unknown model id %d", d->model);
  if (d->model & ~(15 | BDRT_MODEL_POS | BDRT_MODEL_OUTLIERS)) BDRT_FAIL(ctx, BDRT_E_MODEL, "unknown model flags");
  const int nd = base == BDRT_MODEL_SERIES_PARALLEL ? 2 : (base == BDRT_MODEL_SERIES_2PARALLEL ? 3 : 1);
  if (base != BDRT_MODEL_SERIES && (d->model & BDRT_MODEL_OUTLIERS))
    BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED,
              "Parallel_outliers / Series-Parallel*_outliers are dimensionally inconsistent as shipped by the reference "
              "and are not implemented");
  if (d->Nf < 2 || d->K < 3 || d->B < 0 || d->Nf > 4096 || d->K > 4096) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad Nf/K/B");
  if (nd >= 2) {
    if (!d->Ap || !d->Lp) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_series_data: null Ap / Lp for a Series-Parallel model");
    if (d->Kp < 3 || d->Kp > 4096) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad Kp");
    if (!(d->x_sum_invscale >= 0) || !(d->xp_scale > 0)) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad x_sum_invscale / xp_scale");
  }
  if (nd == 3) {
    if (!d->Ap2 || !d->Lp2) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_series_data: null Ap2 / Lp2 for a Series-2Parallel model");
    if (d->Kp2 < 3 || d->Kp2 > 4096) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad Kp2");
    if (!(d->xp2_scale > 0)) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad xp2_scale");
  }
  if (!(d->sigma_min >= 0) || !(d->ups_alpha > 0) || !(d->ups_beta > 0) || !(d->induc_scale > 0))
    BDRT_FAIL(ctx, BDRT_E_SIZE, "model constants must be positive");
  *nd_out = nd;
  return BDRT_OK;
}

// workspace layout shared by the analysis and by every solver call: [info (256 B) | Lb of every distribution | extra]
static size_t bdrt_lb_offsets(const bdrt_series_data* d, int nd, size_t* lb_off) {
  const int Ks[MAXD] = {d->K, d->Kp, d->Kp2};
  size_t head = 256;
  for (int i = 0; i < nd; ++i) {
    lb_off[i] = head;
    head += ((size_t)3 * Ks[i] * LBW * sizeof(double) + 255) & ~(size_t)255;
  }
  return head;
}

// The structure of the matrices (bandwidth / Toeplitz flags / taps): device kernels, then two read-backs that wait
// for the stream.  Done once per problem by bdrt_series_analyze, or per call when the caller passes no info.
static int bdrt_analyze(bdrt_ctx* ctx, const bdrt_series_data* d, int nd, bdrt_series_info* out) {
  const int Ks[MAXD] = {d->K, d->Kp, d->Kp2};
  const double* As[MAXD] = {d->A, d->Ap, d->Ap2};
  const double* Ls[MAXD] = {d->L, d->Lp, d->Lp2};
  size_t lb_off[MAXD];
  const size_t head = bdrt_lb_offsets(d, nd, lb_off);
  int rc = bdrt_ws_reserve(ctx, head);
  if (rc) return rc;
  int* info = (int*)ctx->ws;
  const int init[12] = {0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 0, 0};  // per distribution i: info[4i] = bw, info[4i+1] = L Toeplitz
  BDRT_CUDA(ctx, cudaMemcpyAsync(info, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
  for (int i = 0; i < nd; ++i) {
    double* Lb = (double*)((char*)ctx->ws + lb_off[i]);
    band_prep_kernel<<<3, 256, 0, ctx->stream>>>(Ls[i], Ks[i], Lb, info + 4 * i);
    ctx->launches++;
    if (d->B > 0) {
      const long long stride = d->per_spectrum_grid ? (long long)2 * d->Nf * Ks[i] : 0;
      toep_check_kernel<<<d->per_spectrum_grid ? d->B : 1, 512, 0, ctx->stream>>>(As[i], stride, d->Nf, Ks[i], info);
      ctx->launches++;
    }
  }
  BDRT_CUDA(ctx, cudaGetLastError());
  int hinfo[12];
  BDRT_CUDA(ctx, cudaMemcpyAsync(hinfo, info, sizeof(hinfo), cudaMemcpyDeviceToHost, ctx->stream));
  BDRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memset(out, 0, sizeof(*out));
  out->toepA = hinfo[2];
  for (int i = 0; i < nd; ++i) {
    out->bw[i] = hinfo[4 * i];
    out->toepL[i] = hinfo[4 * i + 1];
    if (hinfo[4 * i] <= MAXBW && Ks[i] >= 2 * FBW + 1) {  // taps = row K/2 of the banded copies
      const double* Lb = (const double*)((char*)ctx->ws + lb_off[i]);
      double row[3][LBW];
      for (int j = 0; j < 3; ++j)
        BDRT_CUDA(ctx, cudaMemcpyAsync(row[j], Lb + ((size_t)j * Ks[i] + Ks[i] / 2) * LBW, LBW * sizeof(double),
                                       cudaMemcpyDeviceToHost, ctx->stream));
      BDRT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      for (int j = 0; j < 3; ++j)
        for (int t = 0; t < 2 * FBW + 1; ++t) out->taps[i][j][t] = row[j][MAXBW - FBW + t];
    }
  }
  out->valid = 1;
  return BDRT_OK;
}

extern "C" int bdrt_series_analyze(bdrt_ctx* ctx, const bdrt_series_data* data, bdrt_series_info* info) {
  if (!ctx) return BDRT_E_NULL;
  if (!info) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_series_analyze: null info");
  int nd = 0;
  int rc = bdrt_check_data(ctx, data, &nd);
  if (rc) return rc;
  return bdrt_analyze(ctx, data, nd, info);
}

// Global-dense mode: the padded, scaled copy of one distribution's kernel matrices in the layout the dense products
// index (rows [re: nfp | im: nfp], pitch lda, zeros in the padding), one per grid.
__global__ void pad_A_kernel(const double* __restrict__ A, long long A_stride, double* __restrict__ Ag, long long ngrid,
                             long long per, int Nf, int nfp, int n2p, int K, int lda, double ascale) {
  const long long tot = ngrid * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
    const long long s = i / per;
    const int e = (int)(i - s * per);
    const int rp = e / lda, c = e - rp * lda;
    const int p = rp >= nfp, r = rp - p * nfp;
    Ag[i] = (rp < n2p && r < Nf && c < K) ? A[s * A_stride + ((long long)p * Nf + r) * K + c] * ascale : 0.0;
  }
}

int bdrt_model_prepare(bdrt_ctx* ctx, const bdrt_series_data* d, BdrtModel* m, size_t extra_ws_bytes,
                       void** extra_ws, int allow_wmode) {
  int nd = 0;
  int rc = bdrt_check_data(ctx, d, &nd);
  if (rc) return rc;
  const int base = d->model & 15;
  memset(m, 0, sizeof(*m));
  m->flags = ((d->model & BDRT_MODEL_POS) ? F_POS : 0) | ((d->model & BDRT_MODEL_OUTLIERS) ? F_OUT : 0);
  m->ND = nd;
  m->Nf = d->Nf;
  m->B = d->B;
  const int Ks[MAXD] = {d->K, d->Kp, d->Kp2};
  const double* As[MAXD] = {d->A, d->Ap, d->Ap2};
  const double* Ls[MAXD] = {d->L, d->Lp, d->Lp2};
  for (int i = 0; i < nd; ++i) {
    m->d[i].K = Ks[i];
    m->d[i].A = As[i];
    m->d[i].A_stride = d->per_spectrum_grid ? (long long)2 * d->Nf * Ks[i] : 0;
    m->d[i].ascale = 1.0;
  }
  m->d[0].pos = (d->model & BDRT_MODEL_POS) ? 1 : 0;
  if (base == BDRT_MODEL_PARALLEL) {
    m->d[0].pos = 1;  // vector<lower=0>[K] x  (Parallel_modelcode.txt:27)
    m->d[0].par = 1;
    m->flags |= F_POS;
  }
  if (nd >= 2) {
    m->d[1].par = 1;
    m->d[1].pos = 1;                 // vector<lower=0>[Kp] xp_raw
    m->d[1].ascale = d->xp_scale;    // xp = xp_raw * xp_scale
    m->x_sum_invscale = d->x_sum_invscale;
  }
  if (nd == 3) {
    m->d[2].par = 1;
    m->d[2].pos = 1;
    m->d[2].ascale = d->xp2_scale;
  }
  m->freq = d->freq;
  m->f_stride = d->per_spectrum_grid ? d->Nf : 0;
  m->Z = d->Z;
  m->sigma_min2 = d->sigma_min * d->sigma_min;
  m->ups_alpha = d->ups_alpha;
  m->ups_beta = d->ups_beta;
  m->induc_scale = d->induc_scale;
  m->so_lambda = d->sigma_out_lambda;
  m->so_alpha = d->sigma_out_alpha;
  m->so_beta = d->sigma_out_beta;
  // structure of the matrices: from the caller (bdrt_series_analyze, no synchronisation here) or found now
  bdrt_series_info local;
  const bdrt_series_info* si = d->info;
  if (!si || !si->valid) {
    rc = bdrt_analyze(ctx, d, nd, &local);
    if (rc) return rc;
    si = &local;
  }
  int bw = 0;
  for (int i = 0; i < nd; ++i) {
    if (si->bw[i] > MAXBW)
      BDRT_FAIL(ctx, BDRT_E_UNSUPPORTED,
                "penalty matrices L0/L1/L2 are not banded within %d off-diagonals (found %d): epsilon is too small "
                "relative to the basis spacing for this build",
                MAXBW, si->bw[i]);
    if (si->bw[i] > bw) bw = si->bw[i];
  }
  m->bw = bw;
  for (int i = 0; i < nd; ++i) m->d[i].toepL = si->toepL[i] && (Ks[i] >= 2 * bw + 1);
  // BDRT_FORCE_DENSE=1 keeps the dense-resident A path even for Toeplitz grids (used by the tests to cover both)
  const char* fd = getenv("BDRT_FORCE_DENSE");
  m->toepA = si->toepA && !(fd && fd[0] == '1');
#ifdef BDRT_PHASE_CLOCKS
  if (!ctx->dbg_clk) {
    BDRT_CUDA(ctx, cudaMalloc(&ctx->dbg_clk, 16 * sizeof(unsigned long long)));
    BDRT_CUDA(ctx, cudaMemset(ctx->dbg_clk, 0, 16 * sizeof(unsigned long long)));
  }
  m->dbg_clk = ctx->dbg_clk;
#endif
  // which Toeplitz engine: warp mode (no CTA rendezvous, per-slot tables) or the cooperative products (8 warps in
  // lockstep -- they share instruction-cache lines, which the 250 kB L-BFGS kernel needs more than it minds the barriers:
  // measured +8 % on the benchmark batch, +14 % at uniform work; NUTS is 2 % faster in warp mode).  BDRT_COOP=1 /
  // BDRT_WARP=1 override the default where the kernel has both instantiations (A/B measurements, tests).
  const char* fc = getenv("BDRT_COOP");
  const char* fw = getenv("BDRT_WARP");
  const bool force_coop = fc && fc[0] == '1', force_warp = fw && fw[0] == '1';
  m->wmode = m->toepA && allow_wmode && !((allow_wmode == 1 || allow_wmode == 3) && force_coop);
  if (allow_wmode == 3 && !d->per_spectrum_grid && !force_warp) m->wmode = 0;
  m->pslot = m->wmode && d->per_spectrum_grid;
  {  // BDRT_WSYNC=1: synchronised warp mode in the solver kernels (A/B measurements)
    const char* fs = getenv("BDRT_WSYNC");
    m->wsync = m->wmode && fs && fs[0] == '1';
  }
  // register-tiled per-slot phases: every distribution has Toeplitz L within FBW off-diagonals and K <= 128
  // (BDRT_FORCE_GENERIC=1 keeps the generic per-slot code, for tests)
  const char* fg = getenv("BDRT_FORCE_GENERIC");
  m->fast = m->toepA && bw <= FBW && !(fg && fg[0] == '1');
  for (int i = 0; i < nd; ++i) m->fast = m->fast && m->d[i].toepL && Ks[i] <= 128 && Ks[i] >= 2 * FBW + 1;
  if (m->fast) {  // taps into the kernel parameter bank
    for (int i = 0; i < nd; ++i)
      for (int j = 0; j < 3; ++j)
        for (int t = 0; t < 2 * FBW + 1; ++t) m->d[i].tapc[j][t] = si->taps[i][j][t];
  }
  int eng = bdrt_model_layout(m);
  // dense operands that do not fit next to the engine's working set (two or three distributions on general grids,
  // matrices.py:243-263): keep padded copies in the workspace and let the products read them through L1 / L2
  // (BDRT_FORCE_GDENSE=1 takes this path whenever the operands are dense, for tests)
  const char* fgd = getenv("BDRT_FORCE_GDENSE");
  if (!m->toepA && ((size_t)eng * 8 > (size_t)ctx->smem_optin || (fgd && fgd[0] == '1'))) {
    m->gdense = 1;
    eng = bdrt_model_layout(m);
  }
  if ((size_t)eng * 8 > (size_t)ctx->smem_optin)
    BDRT_FAIL(ctx, BDRT_E_SMEM, "problem needs %zu B of shared memory per CTA, device offers %d", (size_t)eng * 8,
              ctx->smem_optin);
  size_t lb_off[MAXD], ag_off[MAXD];
  size_t head = bdrt_lb_offsets(d, nd, lb_off);
  const long long ngrid = d->per_spectrum_grid ? d->B : 1;
  if (m->gdense) {
    for (int i = 0; i < nd; ++i) {
      ag_off[i] = head;
      m->d[i].Ag_stride = d->per_spectrum_grid ? (long long)m->n2p * m->d[i].lda + 8 : 0;
      head += ((size_t)ngrid * ((size_t)m->n2p * m->d[i].lda + 8) * sizeof(double) + 255) & ~(size_t)255;
    }
  }
  rc = bdrt_ws_reserve(ctx, head + extra_ws_bytes);
  if (rc) return rc;
  // banded copies of the penalty matrices (device work only; the flags it recomputes are not read back)
  for (int i = 0; i < nd; ++i) {
    double* Lb = (double*)((char*)ctx->ws + lb_off[i]);
    band_prep_kernel<<<3, 256, 0, ctx->stream>>>(Ls[i], Ks[i], Lb, (int*)ctx->ws + 4 * i);
    ctx->launches++;
    m->d[i].Lb = Lb;
  }
  if (m->gdense && d->B > 0) {
    for (int i = 0; i < nd; ++i) {
      double* Ag = (double*)((char*)ctx->ws + ag_off[i]);
      const long long per = (long long)m->n2p * m->d[i].lda + 8;
      const long long tot = ngrid * per;
      const int nblk = (int)((tot + 255) / 256 < 65535LL * 16 ? (tot + 255) / 256 : 65535LL * 16);
      pad_A_kernel<<<nblk, 256, 0, ctx->stream>>>(m->d[i].A, m->d[i].A_stride, Ag, ngrid, per, m->Nf, m->nfp, m->n2p,
                                                   m->d[i].K, m->d[i].lda, m->d[i].ascale);
      ctx->launches++;
      m->d[i].Ag = Ag;
    }
  }
  BDRT_CUDA(ctx, cudaGetLastError());
  if (extra_ws) *extra_ws = (char*)ctx->ws + head;
  return BDRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// log_prob + gradient test hook
// ---------------------------------------------------------------------------------------------------------------------
template <int TOEP, int MK, int FAST>
__global__ void __launch_bounds__(NTHREADS, TOEP ? 2 : 1)
logpost_kernel(BdrtModel m, const double* u, const int* spec, int n_cols, int jacobian, double* lp, double* grad,
               int cta_per_spec) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (TOEP == 2 && m.pslot) {
    // warp mode with per-spectrum grids: every warp walks its own columns and reloads its tables when the spectrum changes
    engine_load(m, sm, 0);
    int loaded = -1;
    for (int c = blockIdx.x * NSLOT + warp; c < n_cols; c += gridDim.x * NSLOT) {
      const int b = spec ? spec[c] : c % m.B;
      if (b != loaded) {
        engine_load_slot(m, sm, b);
        loaded = b;
      }
      const double v = engine_eval<TOEP, MK, FAST>(m, sm, true, u + (long long)c * m.D, grad + (long long)c * m.D,
                                                   m.Z + (long long)b * m.N2, jacobian);
      if (lane == 0) lp[c] = v;
    }
  } else if (!cta_per_spec) {
    engine_load(m, sm, 0);
    const int ngroups = (n_cols + NSLOT - 1) / NSLOT;
    for (int gidx = blockIdx.x; gidx < ngroups; gidx += gridDim.x) {
      const int c = gidx * NSLOT + warp;
      const bool active = c < n_cols;
      const int b = active ? (spec ? spec[c] : c % m.B) : 0;
      const double v = engine_eval<TOEP, MK, FAST>(m, sm, active, u + (long long)c * m.D, grad + (long long)c * m.D,
                                   m.Z + (long long)b * m.N2, jacobian);
      if (active && lane == 0) lp[c] = v;
    }
  } else {
    // per-spectrum grids: the 8 slots of a CTA must share A, so one CTA handles the columns of one spectrum
    for (int b = blockIdx.x; b < m.B; b += gridDim.x) {
      engine_load(m, sm, b);
      // columns of spectrum b are those with spec[c] == b; spec == NULL means c % B == b
      int c = (spec ? 0 : b) - (spec ? 1 : m.B);
      bool more = true;
      while (more) {
        // every warp scans for the next (warp+1)-th matching column after c: simple sequential scan, uniform per CTA
        int found[NSLOT];
        int nf = 0;
        int cc = c;
        while (nf < NSLOT) {
          cc += spec ? 1 : m.B;
          if (cc >= n_cols) break;
          if (!spec || spec[cc] == b) found[nf++] = cc;
        }
        more = (nf == NSLOT);
        c = nf ? found[nf - 1] : c;
        const bool active = warp < nf;
        const int col = active ? found[warp] : 0;
        const double v = engine_eval<TOEP, MK, FAST>(m, sm, active, u + (long long)col * m.D, grad + (long long)col * m.D,
                                     m.Z + (long long)b * m.N2, jacobian);
        if (active && lane == 0) lp[col] = v;
      }
      cta_sync();
    }
  }
}

extern "C" int bdrt_logpost_grad(bdrt_ctx* ctx, const bdrt_series_data* data, const double* u, const int* spec,
                                 int n_cols, int jacobian, double* lp, double* grad) {
  if (!ctx) return BDRT_E_NULL;
  if (!u || !lp || !grad) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_logpost_grad: null pointer");
  if (n_cols < 0) BDRT_FAIL(ctx, BDRT_E_SIZE, "n_cols < 0");
  BdrtModel m;
  int rc = bdrt_model_prepare(ctx, data, &m, 0, nullptr, 1);
  if (rc) return rc;
  if (n_cols == 0 || data->B == 0) return BDRT_OK;
  const BdrtPlan pl = bdrt_plan(ctx, m, 2, 0, 0);
  const int slots = ctx->sm_count * pl.ctas_per_sm;
  int grid;
  if (data->per_spectrum_grid && !m.wmode)
    grid = data->B < slots ? data->B : slots;
  else {
    const int ngroups = (n_cols + NSLOT - 1) / NSLOT;
    grid = ngroups < slots ? ngroups : slots;
  }
  BDRT_LAUNCH(ctx, m, logpost_kernel, grid, pl.smem, m, u, spec, n_cols, jacobian, lp, grad, data->per_spectrum_grid);
  return BDRT_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// constrain: unconstrained -> the constrained / transformed parameters the reference reads back
// (Series_modelcode.txt:37-50; Inverter._extract_parameter, inversion.py:2494-2519)
// one warp per point; plain FP64 FMA dot products (this is a read-out, not the hot loop)
// ---------------------------------------------------------------------------------------------------------------------
__global__ void constrain_kernel(BdrtModel m, const double* u, const int* spec, int n, double* out, int P) {
  const int lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n) return;
  const int Nf = m.Nf;
  const bool outl = m.flags & F_OUT;
  const double* uc = u + c * m.D;
  double* o = out + c * P;
  const int b = spec ? spec[c] : (int)(c % m.B);
  const double* f = m.freq + (long long)b * m.f_stride;
  int KT = 0;  // total number of coefficients
  for (int dd = 0; dd < m.ND; ++dd) {
    const BdrtDist& D = m.d[dd];
    for (int k = lane; k < D.K; k += 32) o[KT + k] = (D.pos ? exp(uc[D.off_x + k]) : uc[D.off_x + k]) * D.ascale;
    KT += D.K;
  }
  const double Rinf = 100.0 * exp(uc[0]), induc = exp(uc[1]) * m.induc_scale;
  const double sr = 0.05 * exp(uc[m.off_err]), ap = 0.05 * exp(uc[m.off_err + 1]), are = 0.05 * exp(uc[m.off_err + 2]),
               aim = 0.05 * exp(uc[m.off_err + 3]);
  if (lane == 0) {
    o[KT] = Rinf;
    o[KT + 1] = induc;
    o[KT + 2] = sr;
    o[KT + 3] = ap;
    o[KT + 4] = are;
    o[KT + 5] = aim;
  }
  __syncwarp();
  for (int nn = lane; nn < Nf; nn += 32) {
    double zz[MAXD][2];
    for (int dd = 0; dd < m.ND; ++dd) {
      const BdrtDist& D = m.d[dd];
      const double* A = D.A + (long long)b * D.A_stride;
      double zre = 0, zim = 0;
      for (int k = 0; k < D.K; ++k) {
        const double xk = (D.pos ? exp(uc[D.off_x + k]) : uc[D.off_x + k]) * D.ascale;
        zre = fma(A[(long long)nn * D.K + k], xk, zre);
        zim = fma(A[(long long)(Nf + nn) * D.K + k], xk, zim);
      }
      zz[dd][0] = zre;
      zz[dd][1] = zim;
    }
    double zre = Rinf, zim = induc * 2.0 * M_PI * f[nn];
    for (int dd = 0; dd < m.ND; ++dd) {
      if (m.d[dd].par) {  // Z_p = 1 / Y  (Parallel_modelcode.txt:46-50, Series-Parallel_modelcode.txt:63-66)
        const double Yr = zz[dd][0], Yi = zz[dd][1], iM = 1.0 / (Yr * Yr + Yi * Yi);
        zre += Yr * iM;
        zim -= Yi * iM;
      } else {
        zre += zz[dd][0];
        zim += zz[dd][1];
      }
    }
    double common = (are * zre) * (are * zre) + (aim * zim) * (aim * zim);
    if (outl) {
      const double so = 0.05 * exp(uc[m.off_so + nn]) * exp(uc[m.off_so + Nf + nn]);
      o[KT + 6 + 2 * Nf + nn] = so;
      common += so * so;
    }
    o[KT + 6 + nn] = sqrt(m.sigma_min2 + sr * sr + (ap * zre) * (ap * zre) + common);
    o[KT + 6 + Nf + nn] = sqrt(m.sigma_min2 + sr * sr + (ap * zim) * (ap * zim) + common);
  }
}

extern "C" int bdrt_constrain(bdrt_ctx* ctx, const bdrt_series_data* data, const double* u, const int* spec, int n,
                              double* out) {
  if (!ctx) return BDRT_E_NULL;
  if (!u || !out) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_constrain: null pointer");
  BdrtModel m;
  int rc = bdrt_model_prepare(ctx, data, &m, 0, nullptr);
  if (rc) return rc;
  if (n <= 0) return BDRT_OK;
  const int P = bdrt_num_outputs(data);
  constrain_kernel<<<(n + 3) / 4, 128, 0, ctx->stream>>>(m, u, spec, n, out, P);
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}
