// Batched hyper-parametric ridge fit: bound-constrained QP + hyper-lambda fixed-point loop.
//
// Replaces cvxopt.solvers.qp inside Inverter._convex_opt (bayes_drt/inversion.py:1043-1067) and the hyper-lambda
// loop of Inverter.ridge_fit (inversion.py:489-753; lambda updates :947-954 and :973-983).
//
// The QP  min 1/2 c'Pc + q'c  s.t. c >= lb  (P = WA'WA + sum_o frac_o Lam_o^1/2 Pen_o Lam_o^1/2, strictly convex, simple
// bounds) is solved EXACTLY by block principal pivoting: every pivot step is one dense Cholesky solve on the free set.
// Mapping: one CTA per spectrum (persistent, grid-stride); the (K+2)^2 working matrix lives in shared memory with an odd
// row stride; bound variables are kept in the system as identity rows/columns so the factorisation never changes size.
// The hyper loop's stop test reproduces the reference's numpy semantics (mean|dc/c| with 0/0 = NaN -> not converged).
#include "common.cuh"

#define RT 256  // threads per CTA

struct RidgeArgs {
  bdrt_ridge_opts o;
  const double *WA_re, *WA_im, *WZ_re, *WZ_im, *Pen, *Lmat;
  long long wa_stride;  // 0: shared WA
  int B, Nf, K, n, ld;
  double* coef;
  double* lam;
  int* iters;
  int* converged;
  double* scratch;  // per CTA: G0 [n*n] | Pg [n*n]
};

namespace {

__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < RT / 32; ++i) s += red[i];
  return s;
}
__device__ double block_max(double v, double* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = red[0];
  for (int i = 1; i < RT / 32; ++i) s = fmax(s, red[i]);
  return s;
}

// In-place Cholesky (lower) of the n x n matrix M (row stride ld) by the whole CTA, then solve M x = r by warp 0.
// r is overwritten by x.  Returns nothing; a non-positive pivot is replaced by a tiny positive number (flagged).
__device__ void chol_solve(double* M, int n, int ld, double* r, int* flag) {
  const int tid = threadIdx.x;
  for (int j = 0; j < n; ++j) {
    if (tid == 0) {
      double d = M[j * ld + j];
      if (!(d > 0.0)) { d = 1e-300; *flag = 1; }
      M[j * ld + j] = sqrt(d);
    }
    __syncthreads();
    const double dj = 1.0 / M[j * ld + j];
    for (int i = j + 1 + tid; i < n; i += RT) M[i * ld + j] *= dj;
    __syncthreads();
    const int mrem = n - j - 1;
    for (int idx = tid; idx < mrem * mrem; idx += RT) {
      const int ii = idx / mrem, kk = idx - ii * mrem;
      if (kk <= ii) {
        const int i = j + 1 + ii, k = j + 1 + kk;
        M[i * ld + k] = fma(-M[i * ld + j], M[k * ld + j], M[i * ld + k]);
      }
    }
    __syncthreads();
  }
  if (tid < 32) {
    // forward: L z = r
    for (int j = 0; j < n; ++j) {
      const double zj = r[j] / M[j * ld + j];
      __syncwarp();
      if (tid == 0) r[j] = zj;
      for (int i = j + 1 + tid; i < n; i += 32) r[i] = fma(-M[i * ld + j], zj, r[i]);
      __syncwarp();
    }
    // backward: L' x = z
    for (int j = n - 1; j >= 0; --j) {
      const double xj = r[j] / M[j * ld + j];
      __syncwarp();
      if (tid == 0) r[j] = xj;
      for (int i = tid; i < j; i += 32) r[i] = fma(-M[j * ld + i], xj, r[i]);
      __syncwarp();
    }
  }
  __syncthreads();
}

}  // namespace

__global__ void __launch_bounds__(RT) ridge_kernel(RidgeArgs a) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x, n = a.n, ld = a.ld, K = a.K, Nf = a.Nf;
  double* M = sm;                 // n * ld
  double* q = M + n * ld;         // n
  double* x = q + n;              // n   current coefficients
  double* prev = x + n;           // n
  double* rhs = prev + n;         // n
  double* y = rhs + n;            // n
  double* lb = y + n;             // n
  double* lam = lb + n;           // 3 n
  double* red = lam + 3 * n;      // 32
  int* F = (int*)(red + 32);      // n  free-set mask
  int* ictl = F + n;              // [0] flag, [1] nv, [2] tries, [3] ninf, [4] max violating index
  double* G0 = a.scratch + (long long)blockIdx.x * 2 * n * n;
  double* Pg = G0 + (long long)n * n;
  const bool integral = a.o.penalty == 1;
  bool have_G0 = false;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const double* WAr = a.WA_re + (long long)b * a.wa_stride;
    const double* WAi = a.WA_im + (long long)b * a.wa_stride;
    const double* WZr = a.WZ_re + (long long)b * Nf;
    const double* WZi = a.WZ_im + (long long)b * Nf;
    // ---- G0 = WA_re' WA_re + WA_im' WA_im  (inversion.py:1045), staged through the (still unused) M region
    if (a.wa_stride != 0 || !have_G0) {
      for (int part = 0; part < 2; ++part) {
        const double* W = part ? WAi : WAr;
        __syncthreads();
        for (int i = tid; i < Nf * n; i += RT) M[i] = W[i];
        __syncthreads();
        for (int idx = tid; idx < n * n; idx += RT) {
          const int i = idx / n, j = idx - i * n;
          double s = part ? G0[idx] : 0.0;
          for (int r = 0; r < Nf; ++r) s = fma(M[r * n + i], M[r * n + j], s);
          G0[idx] = s;
        }
      }
      have_G0 = true;
    }
    // ---- q = -WA_re' WZ_re - WA_im' WZ_im + L1_vec  (inversion.py:1046, :450-452)
    for (int i = tid; i < n; i += RT) {
      double s = 0.0;
      for (int r = 0; r < Nf; ++r) s += WAr[r * n + i] * WZr[r] + WAi[r * n + i] * WZi[r];
      const double l1 = (i < 2) ? 0.0 : sqrt(M_PI) / a.o.epsilon * a.o.L1_penalty;
      q[i] = -s + l1;
      x[i] = 1e-6;  // inversion.py:495
      lb[i] = (a.o.nonneg || i < 2) ? 0.0 : -10.0;  // inversion.py:1054-1064
      F[i] = 0;
      for (int o = 0; o < 3; ++o) lam[o * n + i] = a.o.lambda_0;
    }
    __syncthreads();
    double qm = 0.0;
    for (int i = tid; i < n; i += RT) qm = fmax(qm, fabs(q[i]));
    const double qinf = fmax(block_max(qm, red), 1e-300);

    int it = 0, conv = 0, n_hyper = 0;
    while (it < a.o.max_iter) {
      for (int i = tid; i < n; i += RT) prev[i] = x[i];
      __syncthreads();
      // ---- lambda update from the previous coefficients
      for (int o = 0; o < 3; ++o) {
        if (!(a.o.reg_ord[o] > 0.0)) continue;
        if (!integral) {
          // lam_k = 1 / ((L_o c)_k^2/(beta-1) + 1/lambda_0), lam[0:2] = 1   (inversion.py:947-954)
          const double* Lo = a.Lmat + (long long)o * K * n;
          const int w = tid >> 5, l = tid & 31;
          for (int k = w; k < K; k += RT / 32) {
            double s = 0.0;
            for (int j = l; j < n; j += 32) s = fma(Lo[(long long)k * n + j], prev[j], s);
            s = warp_sum(s);
            if (l == 0) {
              if (a.o.hl_fbeta > 0.0) y[2 + k] = s * s;  // second pass below needs max_k (L c)_k^2
              else lam[o * n + 2 + k] = 1.0 / (s * s / (a.o.hl_beta - 1.0) + 1.0 / a.o.lambda_0);
            }
          }
          if (a.o.hl_fbeta > 0.0) {
            // lam_k = lambda_0 / ((L_o c)_k^2 / (max_k (L_o c)_k^2 * hl_fbeta) + 1)   (inversion.py:956-964)
            __syncthreads();
            double mx = 0.0;
            for (int k = tid; k < K; k += RT) mx = fmax(mx, y[2 + k]);
            mx = block_max(mx, red);
            for (int k = tid; k < K; k += RT) lam[o * n + 2 + k] = a.o.lambda_0 / (y[2 + k] / (mx * a.o.hl_fbeta) + 1.0);
          }
          if (tid < 2) lam[o * n + tid] = 1.0;
        } else {
          // closed form of the integral penalty (inversion.py:973-983), coefficient factors 100 / 10 / 1 (:680-687)
          const double* Mo = a.Pen + (long long)o * n * n;
          const double factor = (o == 0) ? 100.0 : (o == 1 ? 10.0 : 1.0);
          for (int j = tid; j < n; j += RT) rhs[j] = sqrt(lam[o * n + j]);  // previous Lambda^1/2
          __syncthreads();
          for (int j = tid; j < n; j += RT) {
            const double cj = factor * prev[j];
            double C = 0.0;
            for (int i = 0; i < n; ++i)
              if (i != j) C = fma(factor * prev[i] * rhs[i], Mo[(long long)i * n + j], C);
            C *= cj;
            const double aa = a.o.hl_beta / 2.0, bb = 0.5 * (2.0 * aa - 2.0) / a.o.lambda_0;
            const double d = cj * cj * Mo[(long long)j * n + j] + 2.0 * bb;
            const double sg = (C > 0.0) ? 1.0 : (C < 0.0 ? -1.0 : 0.0);
            double lv = (C * C - sg * C * sqrt(4.0 * d * (2.0 * aa - 2.0) + C * C) + 2.0 * d * (2.0 * aa - 2.0)) /
                        (2.0 * d * d);
            if (lv <= 0.0) lv = 1e-15;  // inversion.py:689
            y[j] = lv;
          }
          __syncthreads();
          for (int j = tid; j < n; j += RT) lam[o * n + j] = y[j];
        }
        __syncthreads();
      }
      // ---- P = G0 + sum_o frac_o Lam_o^1/2 Pen_o Lam_o^1/2   (inversion.py:695-700)
      for (int idx = tid; idx < n * n; idx += RT) {
        const int i = idx / n, j = idx - i * n;
        double s = G0[idx];
        for (int o = 0; o < 3; ++o)
          if (a.o.reg_ord[o] > 0.0)
            s += a.o.reg_ord[o] * sqrt(lam[o * n + i]) * a.Pen[(long long)o * n * n + idx] * sqrt(lam[o * n + j]);
        Pg[idx] = s;
      }
      __syncthreads();
      // ---- QP by block principal pivoting, warm-started from the previous free set
      if (tid == 0) { ictl[2] = 3; ictl[3] = n + 1; }
      for (int pit = 0; pit < 500; ++pit) {
        for (int idx = tid; idx < n * n; idx += RT) {
          const int i = idx / n, j = idx - i * n;
          M[i * ld + j] = (F[i] && F[j]) ? Pg[idx] : (i == j ? 1.0 : 0.0);
        }
        for (int i = tid; i < n; i += RT) {
          double r;
          if (F[i]) {
            r = -q[i];
            for (int j = 0; j < n; ++j)
              if (!F[j] && lb[j] != 0.0) r = fma(-Pg[(long long)i * n + j], lb[j], r);
          } else {
            r = lb[i];
          }
          rhs[i] = r;
        }
        if (tid == 0) ictl[0] = 0;
        __syncthreads();
        chol_solve(M, n, ld, rhs, &ictl[0]);
        // y = P x + q on the bound set, violations
        double xm = 0.0;
        for (int i = tid; i < n; i += RT) xm = fmax(xm, fabs(rhs[i]));
        const double tol_x = 1e-14 * fmax(block_max(xm, red), 1e-300), tol_y = 1e-12 * qinf;
        int myv = 0, mymax = -1;
        for (int i = tid; i < n; i += RT) {
          double yi = 0.0;
          if (!F[i]) {
            yi = q[i];
            for (int j = 0; j < n; ++j) yi = fma(Pg[(long long)i * n + j], rhs[j], yi);
          }
          y[i] = yi;
          const int v = F[i] ? (rhs[i] < lb[i] - tol_x) : (yi < -tol_y);
          if (v) { ++myv; mymax = i; }
          F[i] = F[i] | (v << 1);  // bit 1: violation flag, consumed by the exchange step below
        }
        const double nv = block_sum((double)myv, red);
        const double vmax = block_max((double)mymax, red);
        __syncthreads();
        if (nv == 0.0) {
          for (int i = tid; i < n; i += RT) F[i] &= 1;
          __syncthreads();
          break;
        }
        int mode;  // 0: exchange all, 1: exchange only the largest violating index
        {
          const int inv = (int)nv;
          if (inv < ictl[3]) mode = 0;
          else if (ictl[2] >= 1) mode = 0;
          else mode = 1;
          __syncthreads();
          if (tid == 0) {
            if (inv < ictl[3]) { ictl[3] = inv; ictl[2] = 3; }
            else if (ictl[2] >= 1) ictl[2] -= 1;
          }
        }
        for (int i = tid; i < n; i += RT) {
          const int v = (F[i] >> 1) & 1, f = F[i] & 1;
          F[i] = (v && (mode == 0 || i == (int)vmax)) ? (f ^ 1) : f;
        }
        __syncthreads();
      }
      for (int i = tid; i < n; i += RT) x[i] = rhs[i];
      __syncthreads();
      ++n_hyper;
      // ---- stop test with numpy semantics: mean(|(c - prev)/prev|) < xtol; NaN (0/0) compares false  (:730-736)
      double dsum = 0.0;
      for (int i = tid; i < n; i += RT) {
        double d = fabs((x[i] - prev[i]) / prev[i]);
        if (i == 1 && !a.o.fit_inductance) d = 0.0;
        dsum += d;
      }
      dsum = block_sum(dsum, red);
      if (dsum / n < a.o.xtol) { conv = 1; break; }
      ++it;
    }
    for (int i = tid; i < n; i += RT) {
      a.coef[(long long)b * n + i] = x[i];
      for (int o = 0; o < 3; ++o) a.lam[((long long)b * 3 + o) * n + i] = lam[o * n + i];
    }
    if (tid == 0) {
      if (a.iters) a.iters[b] = n_hyper;
      if (a.converged) a.converged[b] = conv;
    }
    __syncthreads();
  }
}

// stand-alone batched QP (test hook for the parity of the solver itself)
__global__ void __launch_bounds__(RT) qp_kernel(const double* P, const double* qv, const double* lbv, int B, int n, int ld,
                                                double* xo, double* kkt, int* iters) {
  extern __shared__ double sm[];
  const int tid = threadIdx.x;
  double* M = sm;
  double* q = M + n * ld;
  double* rhs = q + n;
  double* lb = rhs + n;
  double* red = lb + n;
  int* F = (int*)(red + 32);
  int* ictl = F + n;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const double* Pg = P + (long long)b * n * n;
    double qm = 0.0;
    for (int i = tid; i < n; i += RT) {
      q[i] = qv[(long long)b * n + i];
      lb[i] = lbv[i];
      F[i] = 0;
      qm = fmax(qm, fabs(q[i]));
    }
    const double qinf = fmax(block_max(qm, red), 1e-300);
    if (tid == 0) { ictl[2] = 3; ictl[3] = n + 1; }
    __syncthreads();
    int pit = 0;
    double res = 0.0;
    for (pit = 1; pit <= 500; ++pit) {
      for (int idx = tid; idx < n * n; idx += RT) {
        const int i = idx / n, j = idx - i * n;
        M[i * ld + j] = (F[i] && F[j]) ? Pg[idx] : (i == j ? 1.0 : 0.0);
      }
      for (int i = tid; i < n; i += RT) {
        double r;
        if (F[i]) {
          r = -q[i];
          for (int j = 0; j < n; ++j)
            if (!F[j] && lb[j] != 0.0) r = fma(-Pg[(long long)i * n + j], lb[j], r);
        } else
          r = lb[i];
        rhs[i] = r;
      }
      if (tid == 0) ictl[0] = 0;
      __syncthreads();
      chol_solve(M, n, ld, rhs, &ictl[0]);
      double xm = 0.0;
      for (int i = tid; i < n; i += RT) xm = fmax(xm, fabs(rhs[i]));
      const double tol_x = 1e-14 * fmax(block_max(xm, red), 1e-300), tol_y = 1e-12 * qinf;
      int myv = 0, mymax = -1;
      double myres = 0.0;
      for (int i = tid; i < n; i += RT) {
        double yi = q[i];
        for (int j = 0; j < n; ++j) yi = fma(Pg[(long long)i * n + j], rhs[j], yi);
        const int v = F[i] ? (rhs[i] < lb[i] - tol_x) : (yi < -tol_y);
        // KKT residual: |gradient| on the free set, negative part of the multiplier on the bound set, bound violation
        myres = fmax(myres, F[i] ? fabs(yi) : fmax(0.0, -yi));
        myres = fmax(myres, fmax(0.0, lb[i] - rhs[i]));
        if (v) { ++myv; mymax = i; }
        F[i] = F[i] | (v << 1);
      }
      const double nv = block_sum((double)myv, red);
      const double vmax = block_max((double)mymax, red);
      res = block_max(myres, red);
      __syncthreads();
      if (nv == 0.0) {
        for (int i = tid; i < n; i += RT) F[i] &= 1;
        __syncthreads();
        break;
      }
      int mode;
      {
        const int inv = (int)nv;
        if (inv < ictl[3]) mode = 0;
        else if (ictl[2] >= 1) mode = 0;
        else mode = 1;
        __syncthreads();
        if (tid == 0) {
          if (inv < ictl[3]) { ictl[3] = inv; ictl[2] = 3; }
          else if (ictl[2] >= 1) ictl[2] -= 1;
        }
      }
      for (int i = tid; i < n; i += RT) {
        const int v = (F[i] >> 1) & 1, f = F[i] & 1;
        F[i] = (v && (mode == 0 || i == (int)vmax)) ? (f ^ 1) : f;
      }
      __syncthreads();
    }
    for (int i = tid; i < n; i += RT) xo[(long long)b * n + i] = rhs[i];
    if (tid == 0) {
      if (kkt) kkt[b] = res;
      if (iters) iters[b] = pit;
    }
    __syncthreads();
  }
}

extern "C" void bdrt_ridge_default_opts(bdrt_ridge_opts* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->penalty = 0;  // 'discrete'  (inversion.py:144)
  o->nonneg = 1;
  o->max_iter = 20;
  o->xtol = 1e-3;
  o->hl_beta = 2.5;
  o->lambda_0 = 1e-2;
  o->reg_ord[2] = 1.0;  // reg_ord = 2
  o->L1_penalty = 0.0;
  o->epsilon = 1.0;
  o->fit_inductance = 1;
  o->hl_fbeta = 0.0;  // off: the hl_beta rule
}

static size_t ridge_smem(int n, int ld) { return ((size_t)n * ld + 9 * n + 32) * sizeof(double) + ((size_t)n + 8) * sizeof(int); }

extern "C" int bdrt_ridge_fit(bdrt_ctx* ctx, const bdrt_ridge_opts* opts, const double* WA_re, const double* WA_im,
                              int per_spectrum_W, const double* WZ_re, const double* WZ_im, const double* Pen,
                              const double* Lmat, int B, int Nf, int K, double* coef, double* lam, int* iters,
                              int* converged) {
  if (!ctx) return BDRT_E_NULL;
  if (!opts || !WA_re || !WA_im || !WZ_re || !WZ_im || !Pen || !coef || !lam)
    BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_ridge_fit: null pointer");
  if (opts->penalty != 0 && opts->penalty != 1) BDRT_FAIL(ctx, BDRT_E_MODEL, "penalty must be 0 (discrete) or 1 (integral)");
  if (opts->penalty == 0 && !Lmat) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_ridge_fit: Lmat is required for the discrete penalty");
  if (opts->penalty == 0 && !(opts->hl_beta > 1.0))
    BDRT_FAIL(ctx, BDRT_E_SIZE, "hl_beta must be greater than 1 for penalty 'discrete'");  // inversion.py:286-288
  if (opts->penalty == 1 && !(opts->hl_beta > 2.0))
    BDRT_FAIL(ctx, BDRT_E_SIZE, "hl_beta must be greater than 2 for penalty 'integral'");  // inversion.py:289-291
  if (B < 0 || Nf < 1 || K < 1 || opts->max_iter < 1) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad sizes");
  if (B == 0) return BDRT_OK;
  const int n = K + 2, ld = n | 1;
  if ((long long)Nf * n > (long long)n * ld) BDRT_FAIL(ctx, BDRT_E_SIZE, "Nf > K+2 is not supported by the staging buffer");
  const size_t smem = ridge_smem(n, ld);
  if (smem > (size_t)ctx->smem_optin) BDRT_FAIL(ctx, BDRT_E_SMEM, "K too large for the shared-memory QP solver");
  int per_sm = (int)((size_t)ctx->smem_optin / smem);
  if (per_sm > 2) per_sm = 2;
  int grid = ctx->sm_count * per_sm;
  if (grid > B) grid = B;
  int rc = bdrt_ws_reserve(ctx, (size_t)grid * 2 * n * n * sizeof(double));
  if (rc) return rc;
  RidgeArgs a;
  a.o = *opts;
  a.WA_re = WA_re; a.WA_im = WA_im; a.WZ_re = WZ_re; a.WZ_im = WZ_im; a.Pen = Pen; a.Lmat = Lmat;
  a.wa_stride = per_spectrum_W ? (long long)Nf * n : 0;
  a.B = B; a.Nf = Nf; a.K = K; a.n = n; a.ld = ld;
  a.coef = coef; a.lam = lam; a.iters = iters; a.converged = converged;
  a.scratch = (double*)ctx->ws;
  BDRT_CUDA(ctx, cudaFuncSetAttribute(ridge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ridge_kernel<<<grid, RT, smem, ctx->stream>>>(a);
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}

extern "C" int bdrt_qp_bound(bdrt_ctx* ctx, const double* P, const double* q, const double* lb, int B, int n, double* x,
                             double* kkt, int* iters) {
  if (!ctx) return BDRT_E_NULL;
  if (!P || !q || !lb || !x) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_qp_bound: null pointer");
  if (B < 0 || n < 1) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad sizes");
  if (B == 0) return BDRT_OK;
  const int ld = n | 1;
  const size_t smem = ((size_t)n * ld + 3 * n + 32) * sizeof(double) + ((size_t)n + 8) * sizeof(int);
  if (smem > (size_t)ctx->smem_optin) BDRT_FAIL(ctx, BDRT_E_SMEM, "n too large for the shared-memory QP solver");
  int per_sm = (int)((size_t)ctx->smem_optin / smem);
  if (per_sm > 2) per_sm = 2;
  int grid = ctx->sm_count * per_sm;
  if (grid > B) grid = B;
  BDRT_CUDA(ctx, cudaFuncSetAttribute(qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qp_kernel<<<grid, RT, smem, ctx->stream>>>(P, q, lb, B, n, ld, x, kkt, iters);
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}
