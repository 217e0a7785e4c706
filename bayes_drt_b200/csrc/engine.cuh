// Fused log-posterior + analytic-gradient engine for the 'Series' family of Stan programs
// (bayes_drt/stan_model_files/Series_modelcode.txt:24-69, Series_pos_modelcode.txt:27,
//  Series_outliers_modelcode.txt:22-72, Series_pos_outliers_modelcode.txt:25).
//
// Execution model (B200): one persistent CTA of 8 warps per SM.  The CTA owns NSLOT = 8 "column slots"; slot s is
// driven by warp s, which runs its own copy of the calling algorithm (L-BFGS, NUTS, ...) with ordinary warp-uniform
// control flow.  Whenever the algorithms need log p and its gradient they all call engine_eval(): the stacked kernel
// matrix A (2Nf x K, FP64) stays resident in shared memory for the lifetime of the CTA and the two dense products
//     Zhat = A   X   (2Nf x K) (K x 8)     and     GX = A^T V   (K x 2Nf) (2Nf x 8)
// are done cooperatively by all 8 warps for all 8 slots at once on the FP64 tensor cores
// (mma.sync.m8n8k4.f64: the 8 slots are exactly the N=8 columns of the DMMA tile), while everything that is
// per-slot and O(Nf + K) -- error model, residual weights, banded derivative stencils, hyper-priors, chain rule to
// the unconstrained space -- is done by the slot's own warp with shuffle reductions.  Four CTA barriers per
// evaluation; no global-memory traffic besides the slot's own u / grad vectors and its spectrum Z.
//
// Shared-memory strides are chosen so that every DMMA fragment load is bank-conflict free without any swizzle:
//   A [row][lda], lda % 16 in {4, 12}: the A-operand fragment of A (8 rows x 4 cols) and of A^T (4 rows x 8 cols) both
//   hit 16 distinct 8-byte banks per half-warp;  X / V [slot][ldxv] with the same rule for the B-operand fragment.
#pragma once
#include "common.cuh"

#define NSLOT 8
#define NWARP 8
#define NTHREADS (NWARP * 32)
#define MAXBW 24
#define LBW (2 * MAXBW + 1)
#define LOG_015 (-1.8971199848858813)  // log(0.15)

#define F_POS 1
#define F_OUT 2

struct BdrtModel {
  int flags, Nf, K, N2, D, B;
  int lda, ldxv, ldzg;
  int kpad4, kpad8, n2pad4, n2pad8;
  int off_so, off_ups, off_d;
  int bw, toeplitz;
  const double* A;
  long long A_stride;  // per-spectrum stride (0: shared)
  const double* freq;
  long long f_stride;
  const double* Z;   // [B, N2]
  const double* Lb;  // [3][K][LBW] banded copies of the scaled L0, L1, L2
  double sigma_min2, ups_alpha, ups_beta, induc_scale, so_lambda, so_alpha, so_beta;
  // shared-memory carve-up, in doubles
  int oA, oXV, oZG, oSt, oTap, oOm, oUser;
};

static inline int bdrt_pad_stride(int n) {  // smallest s >= n with s % 16 in {4, 12}
  int s = n;
  while ((s % 16) != 4 && (s % 16) != 12) ++s;
  return s;
}

// Fills the derived fields of m (everything but the pointers / scalars).  Returns doubles of engine smem.
static inline int bdrt_model_layout(BdrtModel* m) {
  m->N2 = 2 * m->Nf;
  m->kpad4 = (m->K + 3) / 4 * 4;
  m->kpad8 = (m->K + 7) / 8 * 8;
  m->n2pad4 = (m->N2 + 3) / 4 * 4;
  m->n2pad8 = (m->N2 + 7) / 8 * 8;
  m->lda = bdrt_pad_stride(m->K);
  int mx = m->kpad4 > m->n2pad4 ? m->kpad4 : m->n2pad4;
  m->ldxv = bdrt_pad_stride(mx);
  int mz = m->kpad8 > m->n2pad8 ? m->kpad8 : m->n2pad8;
  m->ldzg = mz + 4;  // % 8 == 4
  m->off_so = 6 + m->K;
  m->off_ups = 6 + m->K + ((m->flags & F_OUT) ? 2 * m->Nf : 0);
  m->off_d = m->off_ups + m->K;
  m->D = m->off_d + 3;
  int o = 0;
  m->oA = o;   o += m->n2pad8 * m->lda + 8;
  m->oXV = o;  o += NSLOT * m->ldxv;
  m->oZG = o;  o += NSLOT * m->ldzg;
  m->oSt = o;  o += NSLOT * 4 * m->K;
  m->oTap = o; o += 3 * LBW;
  m->oOm = o;  o += m->Nf;
  o = (o + 1) & ~1;
  m->oUser = o;
  return o;
}

#ifdef __CUDACC__
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cta_sync() {
  __syncwarp();
  asm volatile("bar.sync 0;" ::: "memory");
}

// Cooperative load of the resident operands.  Called by all threads once (or once per spectrum when the grid is
// per-spectrum); ends with a CTA barrier.
__device__ inline void engine_load(const BdrtModel& m, double* sm, long long spec) {
  double* sA = sm + m.oA;
  const double* gA = m.A + spec * m.A_stride;
  const int tid = threadIdx.x;
  const int rows = m.n2pad8;
  for (int i = tid; i < rows * m.lda + 8; i += NTHREADS) {
    const int r = i / m.lda, c = i - r * m.lda;
    sA[i] = (r < m.N2 && c < m.K) ? gA[(long long)r * m.K + c] : 0.0;  // coalesced along c
  }
  for (int i = tid; i < NSLOT * m.ldxv; i += NTHREADS) sm[m.oXV + i] = 0.0;
  for (int i = tid; i < NSLOT * m.ldzg; i += NTHREADS) sm[m.oZG + i] = 0.0;
  // Toeplitz taps: row K/2 of the banded copies
  for (int i = tid; i < 3 * LBW; i += NTHREADS) {
    const int j = i / LBW, d = i - j * LBW;
    sm[m.oTap + i] = m.Lb[((long long)j * m.K + m.K / 2) * LBW + d];
  }
  const double* f = m.freq + spec * m.f_stride;
  for (int i = tid; i < m.Nf; i += NTHREADS) sm[m.oOm + i] = 2.0 * M_PI * f[i];
  cta_sync();
}

__device__ __forceinline__ double tap_at(const BdrtModel& m, const double* sTap, int j, int row, int d) {
  // L_j[row][row + d]
  return m.toeplitz ? sTap[j * LBW + d + MAXBW] : __ldg(m.Lb + ((long long)j * m.K + row) * LBW + d + MAXBW);
}

// log p(u) and d/du for the slot of the calling warp; all NWARP warps of the CTA must call it together.
//   active : this slot has a point to evaluate (inactive slots only help with the matrix products)
//   u, grad: the slot's D-vectors (generic pointers: shared or global)
//   Zs     : the slot's stacked scaled spectrum [N2] (global)
// Returns lp (non-finite lp or gradient entries must be checked by the caller).
//   nact/snap: optional CTA-wide "slots still working" counter; *snap receives its value at a point where no warp can
//           be modifying it (between the first and last barrier), so every warp of the CTA reads the same value.
__device__ inline double engine_eval(const BdrtModel& m, double* sm, bool active, const double* u, double* grad,
                                     const double* Zs, int jacobian, const volatile int* nact = nullptr,
                                     int* snap = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp;
  const int K = m.K, Nf = m.Nf;
  const bool pos = m.flags & F_POS, outl = m.flags & F_OUT;
  double* sA = sm + m.oA;
  double* sX = sm + m.oXV + slot * m.ldxv;
  double* sZ = sm + m.oZG + slot * m.ldzg;
  double* sW = sm + m.oSt + slot * 4 * K;  // w0 | w1 | w2 | ups
  const double* sTap = sm + m.oTap;
  const double* sOm = sm + m.oOm;
  const double jac = jacobian ? 1.0 : 0.0;

  double lp = 0.0;
  double rinf_raw = 0, ind_raw = 0, sr_raw = 0, ap_raw = 0, are_raw = 0, aim_raw = 0;
  // ---------------------------------------------------------------- phase 1: x -> smem, priors, stencils (per slot)
  if (active) {
    rinf_raw = exp(u[0]);
    ind_raw = exp(u[1]);
    sr_raw = exp(u[2 + K]);
    ap_raw = exp(u[3 + K]);
    are_raw = exp(u[4 + K]);
    aim_raw = exp(u[5 + K]);
    const double d0 = exp(u[m.off_d]), d1 = exp(u[m.off_d + 1]), d2 = exp(u[m.off_d + 2]);
    double ujac = 0.0;
    for (int k = lane; k < m.kpad4; k += 32) {
      double xv = 0.0;
      if (k < K) {
        const double uk = u[2 + k];
        xv = pos ? exp(uk) : uk;
        if (pos) ujac += uk;
      }
      sX[k] = xv;
    }
    __syncwarp();
    double sa0 = 0, sa1 = 0, sa2 = 0;
    const int bw = m.bw;
    for (int k = lane; k < K; k += 32) {
      double a0 = 0, a1 = 0, a2 = 0;
      const int dlo = (k - bw < 0) ? -k : -bw, dhi = (k + bw > K - 1) ? (K - 1 - k) : bw;
      for (int d = dlo; d <= dhi; ++d) {
        const double xv = sX[k + d];
        a0 = fma(tap_at(m, sTap, 0, k, d), xv, a0);
        a1 = fma(tap_at(m, sTap, 1, k, d), xv, a1);
        a2 = fma(tap_at(m, sTap, 2, k, d), xv, a2);
      }
      const double uk = u[m.off_ups + k];
      ujac += uk;
      const double ups_raw = exp(uk), ups = 0.15 * ups_raw;
      const double iu = 1.0 / ups, iu2 = iu * iu;
      const double q2 = d0 * a0 * a0 + d1 * a1 * a1 + d2 * a2 * a2;
      // q ~ normal(0, ups): -1/2 q^2/ups^2 - log ups ;  ups_raw ~ inv_gamma(alpha, beta)
      lp += -0.5 * q2 * iu2 - (LOG_015 + uk) - (m.ups_alpha + 1.0) * uk - m.ups_beta / ups_raw;
      sa0 = fma(a0 * a0, iu2, sa0);
      sa1 = fma(a1 * a1, iu2, sa1);
      sa2 = fma(a2 * a2, iu2, sa2);
      sW[k] = d0 * a0 * iu2;
      sW[K + k] = d1 * a1 * iu2;
      sW[2 * K + k] = d2 * a2 * iu2;
      sW[3 * K + k] = ups;
      grad[m.off_ups + k] = q2 * iu2 * iu - iu;  // d lp / d ups_k without the dups terms (finished below)
    }
    __syncwarp();
    const double* su = sW + 3 * K;
    for (int k = lane; k < K; k += 32) {
      // dups_j = 0.5 - 0.25 (ups_j + ups_{j+2}) / ups_{j+1},  j = 0..K-3   (Series_modelcode.txt:51-53)
      double gu = grad[m.off_ups + k];
      const double uk = su[k];
      if (k + 2 < K) {  // k is the left point of dups_k
        const double e = 0.5 - 0.25 * (uk + su[k + 2]) / su[k + 1];
        gu += e * 0.25 / su[k + 1];
        lp += -0.5 * e * e;
      }
      if (k >= 1 && k + 1 < K) {  // middle point of dups_{k-1}
        const double sum = su[k - 1] + su[k + 1];
        const double e = 0.5 - 0.25 * sum / uk;
        gu -= e * 0.25 * sum / (uk * uk);
      }
      if (k >= 2) {  // right point of dups_{k-2}
        const double e = 0.5 - 0.25 * (su[k - 2] + uk) / su[k - 1];
        gu += e * 0.25 / su[k - 1];
      }
      const double ups_raw = uk * (1.0 / 0.15);
      grad[m.off_ups + k] = gu * uk - (m.ups_alpha + 1.0) + m.ups_beta / ups_raw + jac;
      // prior part of d lp / d x_k:  - sum_j (L_j^T w_j)_k
      double acc = 0.0;
      const int dlo = (k - bw < 0) ? -k : -bw, dhi = (k + bw > K - 1) ? (K - 1 - k) : bw;
      for (int d = dlo; d <= dhi; ++d) {  // row n = k + d, column k  ->  offset -d in that row
        const int n = k + d;
        acc = fma(tap_at(m, sTap, 0, n, -d), sW[n], acc);
        acc = fma(tap_at(m, sTap, 1, n, -d), sW[K + n], acc);
        acc = fma(tap_at(m, sTap, 2, n, -d), sW[2 * K + n], acc);
      }
      grad[2 + k] = -acc;
    }
    sa0 = warp_sum(sa0);
    sa1 = warp_sum(sa1);
    sa2 = warp_sum(sa2);
    if (lane == 0) {
      // d_j ~ inv_gamma(5, 5): -6 log d - 5/d
      lp += -6.0 * (u[m.off_d] + u[m.off_d + 1] + u[m.off_d + 2]) - 5.0 / d0 - 5.0 / d1 - 5.0 / d2;
      lp += -0.5 * (rinf_raw * rinf_raw + ind_raw * ind_raw + sr_raw * sr_raw + ap_raw * ap_raw + are_raw * are_raw +
                    aim_raw * aim_raw);
      grad[m.off_d] = -0.5 * sa0 * d0 - 6.0 + 5.0 / d0 + jac;
      grad[m.off_d + 1] = -0.5 * sa1 * d1 - 6.0 + 5.0 / d1 + jac;
      grad[m.off_d + 2] = -0.5 * sa2 * d2 - 6.0 + 5.0 / d2 + jac;
      ujac += u[0] + u[1] + u[2 + K] + u[3 + K] + u[4 + K] + u[5 + K] + u[m.off_d] + u[m.off_d + 1] + u[m.off_d + 2];
    }
    if (outl)
      for (int n = lane; n < 2 * Nf; n += 32) ujac += u[m.off_so + n];
    if (jacobian) lp += ujac;
  } else {
    for (int k = lane; k < m.kpad4; k += 32) sX[k] = 0.0;
  }
  cta_sync();
  if (snap) *snap = *nact;

  // ---------------------------------------------------------------- phase 2: Zhat = A X on the FP64 tensor cores
  const int g = lane >> 2, t = lane & 3;
  {
    const double* bp = sm + m.oXV + g * m.ldxv + t;
    double* zg = sm + m.oZG;
    const int nmt = m.n2pad8 >> 3;
    for (int mt = warp; mt < nmt; mt += 2 * NWARP) {
      const int mt2 = mt + NWARP;
      const bool two = mt2 < nmt;
      const double* a0p = sA + (mt * 8 + g) * m.lda + t;
      const double* a1p = sA + ((two ? mt2 : mt) * 8 + g) * m.lda + t;
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
#pragma unroll 5
      for (int kk = 0; kk < m.kpad4; kk += 4) {
        const double b = bp[kk];
        dmma(c00, c01, a0p[kk], b);
        dmma(c10, c11, a1p[kk], b);
      }
      zg[(2 * t) * m.ldzg + mt * 8 + g] = c00;
      zg[(2 * t + 1) * m.ldzg + mt * 8 + g] = c01;
      if (two) {
        zg[(2 * t) * m.ldzg + mt2 * 8 + g] = c10;
        zg[(2 * t + 1) * m.ldzg + mt2 * 8 + g] = c11;
      }
    }
  }
  cta_sync();

  // ---------------------------------------------------------------- phase 3: error model, residual weights (per slot)
  double* sV = sX;
  if (active) {
    const double Rinf = 100.0 * rinf_raw, induc = ind_raw * m.induc_scale;
    const double sr = 0.05 * sr_raw, ap = 0.05 * ap_raw, are = 0.05 * are_raw, aim = 0.05 * aim_raw;
    const double base = m.sigma_min2 + sr * sr;
    const double ap2 = ap * ap, are2 = are * are, aim2 = aim * aim;
    double Sv = 0, Swv = 0, Sg = 0, Sgz = 0, SGre = 0, SGim = 0;
    for (int n = lane; n < Nf; n += 32) {
      const double om = sOm[n];
      const double zre = sZ[n] + Rinf, zim = sZ[Nf + n] + induc * om;
      double common = are2 * zre * zre + aim2 * zim * zim;
      double so_raw = 0, so_scale = 0, so = 0, usc = 0;
      if (outl) {
        so_raw = exp(u[m.off_so + n]);
        usc = u[m.off_so + Nf + n];
        so_scale = exp(usc);
        so = 0.05 * so_raw * so_scale;  // Series_outliers_modelcode.txt:45
        common += so * so;
        // sigma_out_raw ~ exponential(lambda); sigma_out_scale ~ inv_gamma(alpha, beta)
        lp += -m.so_lambda * so_raw - (m.so_alpha + 1.0) * usc - m.so_beta / so_scale;
      }
      const double s_re = base + ap2 * zre * zre + common, s_im = base + ap2 * zim * zim + common;
      const double i_re = 1.0 / s_re, i_im = 1.0 / s_im;
      const double r_re = Zs[n] - zre, r_im = Zs[Nf + n] - zim;
      lp += -0.5 * (r_re * r_re * i_re + r_im * r_im * i_im) - 0.5 * (log(s_re) + log(s_im));
      const double g_re = 0.5 * r_re * r_re * i_re * i_re - 0.5 * i_re;
      const double g_im = 0.5 * r_im * r_im * i_im * i_im - 0.5 * i_im;
      const double G = g_re + g_im;
      const double v_re = r_re * i_re + 2.0 * zre * (ap2 * g_re + are2 * G);
      const double v_im = r_im * i_im + 2.0 * zim * (ap2 * g_im + aim2 * G);
      sV[n] = v_re;
      sV[Nf + n] = v_im;
      Sv += v_re;
      Swv = fma(om, v_im, Swv);
      Sg += G;
      Sgz += g_re * zre * zre + g_im * zim * zim;
      SGre = fma(G, zre * zre, SGre);
      SGim = fma(G, zim * zim, SGim);
      if (outl) {
        const double dso = 2.0 * so * G * so;  // (d lp/d sigma_out) * sigma_out ; sigma_out = .05 raw scale
        grad[m.off_so + n] = dso - m.so_lambda * so_raw + jac;
        grad[m.off_so + Nf + n] = dso - (m.so_alpha + 1.0) + m.so_beta / so_scale + jac;
      }
    }
    for (int n = m.N2 + lane; n < m.n2pad4; n += 32) sV[n] = 0.0;
    Sv = warp_sum(Sv);
    Swv = warp_sum(Swv);
    Sg = warp_sum(Sg);
    Sgz = warp_sum(Sgz);
    SGre = warp_sum(SGre);
    SGim = warp_sum(SGim);
    if (lane == 0) {
      grad[0] = (100.0 * Sv - rinf_raw) * rinf_raw + jac;
      grad[1] = (m.induc_scale * Swv - ind_raw) * ind_raw + jac;
      grad[2 + K] = (0.1 * sr * Sg - sr_raw) * sr_raw + jac;
      grad[3 + K] = (0.1 * ap * Sgz - ap_raw) * ap_raw + jac;
      grad[4 + K] = (0.1 * are * SGre - are_raw) * are_raw + jac;
      grad[5 + K] = (0.1 * aim * SGim - aim_raw) * aim_raw + jac;
    }
  } else {
    for (int n = lane; n < m.n2pad4; n += 32) sV[n] = 0.0;
  }
  cta_sync();

  // ---------------------------------------------------------------- phase 4: GX = A^T V on the FP64 tensor cores
  {
    const double* bp = sm + m.oXV + g * m.ldxv + t;
    double* zg = sm + m.oZG;
    const int nmt = m.kpad8 >> 3;
    for (int mt = warp; mt < nmt; mt += 2 * NWARP) {
      const int mt2 = mt + NWARP;
      const bool two = mt2 < nmt;
      const double* a0p = sA + t * m.lda + mt * 8 + g;             // A^T[kk0+g][i0+t] = A[i0+t][kk0+g]
      const double* a1p = sA + t * m.lda + (two ? mt2 : mt) * 8 + g;
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
      const int step = 4 * m.lda;
#pragma unroll 5
      for (int i0 = 0, ao = 0; i0 < m.n2pad4; i0 += 4, ao += step) {
        const double b = bp[i0];
        dmma(c00, c01, a0p[ao], b);
        dmma(c10, c11, a1p[ao], b);
      }
      zg[(2 * t) * m.ldzg + mt * 8 + g] = c00;
      zg[(2 * t + 1) * m.ldzg + mt * 8 + g] = c01;
      if (two) {
        zg[(2 * t) * m.ldzg + mt2 * 8 + g] = c10;
        zg[(2 * t + 1) * m.ldzg + mt2 * 8 + g] = c11;
      }
    }
  }
  cta_sync();

  // ---------------------------------------------------------------- phase 5: assemble d lp / d u_x (per slot)
  if (active) {
    for (int k = lane; k < K; k += 32) {
      double gx = sZ[k] + grad[2 + k];
      if (pos) gx = gx * exp(u[2 + k]) + jac;
      grad[2 + k] = gx;
    }
    lp = warp_sum(lp);
    __syncwarp();
  }
  return lp;
}
#endif  // __CUDACC__

// host side (model.cu)
int bdrt_model_prepare(bdrt_ctx* ctx, const bdrt_series_data* data, BdrtModel* m, size_t extra_ws_bytes,
                       void** extra_ws);
