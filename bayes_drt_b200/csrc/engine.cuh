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
  int kpad4, kpad8;
  int nfp;    // Nf rounded up to a multiple of 8: inside the engine the stacked vectors are [re: nfp | im: nfp]
  int n2p;    // 2 * nfp
  int toepA;  // A_re / A_im are Toeplitz (shared log-uniform grid): resident operand = two 1-D tables of length lt
  int lt;     // nfp + kpad8
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
  int xoff;  // offset of the data inside an X/V row (= bw: zero margin for the stencils)
  int ws;    // stride of one stencil scratch vector (K + 2 bw, zero margins)
  int st;    // per-slot scratch size
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
  m->nfp = (m->Nf + 7) / 8 * 8;
  m->n2p = 2 * m->nfp;
  m->lt = m->nfp + m->kpad8;
  m->lda = bdrt_pad_stride(m->K);
  m->xoff = m->bw;
  int kx = m->kpad4 > m->K + m->bw ? m->kpad4 : m->K + m->bw;
  int mx = kx > m->n2p ? kx : m->n2p;
  m->ldxv = bdrt_pad_stride(m->xoff + mx);
  m->ws = m->K + 2 * m->bw;
  // per-slot scratch: W0 | W1 | W2 (ws each) | ups (K) | 1/ups (K) | scalars (16) | sigma_out raw, scale (2 Nf)
  m->st = 3 * m->ws + 2 * m->K + 16 + ((m->flags & F_OUT) ? 2 * m->Nf : 0);
  int mz = m->kpad8 > m->n2p ? m->kpad8 : m->n2p;
  m->ldzg = mz + 4;  // % 8 == 4
  m->off_so = 6 + m->K;
  m->off_ups = 6 + m->K + ((m->flags & F_OUT) ? 2 * m->Nf : 0);
  m->off_d = m->off_ups + m->K;
  m->D = m->off_d + 3;
  int o = 0;
  m->oA = o;   o += m->toepA ? 2 * m->lt : m->n2p * m->lda + 8;
  m->oXV = o;  o += NSLOT * m->ldxv;
  m->oZG = o;  o += NSLOT * m->ldzg;
  m->oSt = o;  o += NSLOT * m->st;
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
  if (m.toepA) {
    // table of part p: T_p[(col - row) + nfp - 1] = A_p[row][col]; first row for col - row >= 0, first column below
    for (int i = tid; i < 2 * m.lt; i += NTHREADS) {
      const int p = i >= m.lt, d = i - p * m.lt - (m.nfp - 1);
      double v = 0.0;
      if (d >= 0 && d < m.K) v = gA[(long long)p * m.Nf * m.K + d];
      else if (d < 0 && -d < m.Nf) v = gA[((long long)p * m.Nf - d) * m.K];
      sA[i] = v;
    }
  } else {
    for (int i = tid; i < m.n2p * m.lda + 8; i += NTHREADS) {
      const int rp = i / m.lda, c = i - rp * m.lda;
      const int p = rp >= m.nfp, r = rp - p * m.nfp;  // padded row -> (part, row)
      sA[i] = (rp < m.n2p && r < m.Nf && c < m.K) ? gA[((long long)p * m.Nf + r) * m.K + c] : 0.0;  // coalesced along c
    }
  }
  for (int i = tid; i < NSLOT * m.ldxv; i += NTHREADS) sm[m.oXV + i] = 0.0;
  for (int i = tid; i < NSLOT * m.ldzg; i += NTHREADS) sm[m.oZG + i] = 0.0;
  for (int i = tid; i < NSLOT * m.st; i += NTHREADS) sm[m.oSt + i] = 0.0;
  // Toeplitz taps: row K/2 of the banded copies
  for (int i = tid; i < 3 * LBW; i += NTHREADS) {
    const int j = i / LBW, d = i - j * LBW;
    sm[m.oTap + i] = m.Lb[((long long)j * m.K + m.K / 2) * LBW + d];
  }
  const double* f = m.freq + spec * m.f_stride;
  for (int i = tid; i < m.Nf; i += NTHREADS) sm[m.oOm + i] = 2.0 * M_PI * f[i];
  cta_sync();
}

// log p(u) and d/du for the slot of the calling warp; all NWARP warps of the CTA must call it together.
//   active : this slot has a point to evaluate (inactive slots only help with the matrix products)
//   u, grad: the slot's D-vectors (generic pointers: shared or global)
//   Zs     : the slot's stacked scaled spectrum [N2] (global)
//   nact/snap: optional CTA-wide "slots still working" counter; *snap receives its value at a point where no warp can
//           be modifying it (between the first and last barrier), so every warp of the CTA reads the same value.
// Returns lp (non-finite lp or gradient entries must be checked by the caller).
template <int TOEP>
__device__ inline double engine_eval(const BdrtModel& m, double* sm, bool active, const double* u, double* grad,
                                     const double* Zs, int jacobian, const volatile int* nact = nullptr,
                                     int* snap = nullptr) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp;
  const int K = m.K, Nf = m.Nf, bw = m.bw;
  const bool pos = m.flags & F_POS, outl = m.flags & F_OUT;
  double* sA = sm + m.oA;
  double* sX = sm + m.oXV + slot * m.ldxv + m.xoff;  // x_k at sX[k], zero margins of width bw on both sides
  double* sZ = sm + m.oZG + slot * m.ldzg;
  double* sW = sm + m.oSt + slot * m.st + bw;         // W_j[k] at sW[j * ws + k], zero margins
  double* sUps = sm + m.oSt + slot * m.st + 3 * m.ws;
  double* sIu = sUps + K;
  double* sTh = sIu + K;   // Rinf_raw, induc_raw, sigma_res_raw, alpha_prop/re/im_raw, d0, d1, d2
  double* sSo = sTh + 16;  // sigma_out_raw [Nf], sigma_out_scale [Nf]
  const double* sTap = sm + m.oTap + MAXBW;  // tap_j[d] at sTap[j * LBW + d]
  const double* sOm = sm + m.oOm;
  const double jac = jacobian ? 1.0 : 0.0;

  double lp = 0.0;
  // ---------------------------------------------------------------- phase 1: transforms, priors, stencils (per slot)
  if (active) {
    // 1a. theta = exp(u) for every lower=0 parameter, scattered to the slot's scratch (one coalesced pass over u)
    double ujac = 0.0;
    for (int i = lane; i < m.D; i += 32) {
      const double ui = u[i];
      const bool isx = (i >= 2) && (i < 2 + K);
      const double th = (isx && !pos) ? ui : exp(ui);
      if (!isx || pos) ujac += ui;
      if (isx)
        sX[i - 2] = th;
      else if (i < 2)
        sTh[i] = th;
      else if (i < 6 + K)
        sTh[i - K] = th;  // 2+K..5+K -> 2..5
      else if (i < m.off_ups)
        sSo[i - m.off_so] = th;
      else if (i < m.off_d) {
        const double ups = 0.15 * th;
        sUps[i - m.off_ups] = ups;
        sIu[i - m.off_ups] = 1.0 / ups;
      } else
        sTh[6 + i - m.off_d] = th;
    }
    {
      const int kend = (m.kpad4 > K + bw) ? m.kpad4 : K + bw;
      for (int k = K + lane; k < kend; k += 32) sX[k] = 0.0;  // right margin (phase 3 of the previous call wrote V here)
    }
    __syncwarp();
    const double d0 = sTh[6], d1 = sTh[7], d2 = sTh[8];
    double sa0 = 0, sa1 = 0, sa2 = 0;
    // 1b. a_j = L_j x (banded), q^2, hyper-priors, d lp / d ups, W_j = d_j a_j / ups^2
    for (int kb = 0; kb < K; kb += 128) {
      double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
      if (m.toeplitz) {
        for (int d = -bw; d <= bw; ++d) {
          const double t0 = sTap[d], t1 = sTap[LBW + d], t2 = sTap[2 * LBW + d];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kb + lane + 32 * j;
            const double xv = (k < K) ? sX[k + d] : 0.0;
            a0[j] = fma(t0, xv, a0[j]);
            a1[j] = fma(t1, xv, a1[j]);
            a2[j] = fma(t2, xv, a2[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kb + lane + 32 * j;
          if (k < K) {
            const double* l0 = m.Lb + (long long)k * LBW + MAXBW;
            const double* l1 = l0 + (long long)K * LBW;
            const double* l2 = l1 + (long long)K * LBW;
            for (int d = -bw; d <= bw; ++d) {
              const double xv = sX[k + d];
              a0[j] = fma(__ldg(l0 + d), xv, a0[j]);
              a1[j] = fma(__ldg(l1 + d), xv, a1[j]);
              a2[j] = fma(__ldg(l2 + d), xv, a2[j]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kb + lane + 32 * j;
        if (k < K) {
          const double ups = sUps[k], iu = sIu[k], iu2 = iu * iu;
          const double uk = u[m.off_ups + k];
          const double q2 = d0 * a0[j] * a0[j] + d1 * a1[j] * a1[j] + d2 * a2[j] * a2[j];
          // q ~ normal(0, ups): -1/2 q^2/ups^2 - log ups ;  ups_raw ~ inv_gamma(alpha, beta): -(alpha+1) log - beta/ups_raw
          lp += -0.5 * q2 * iu2 - (LOG_015 + uk) - (m.ups_alpha + 1.0) * uk - m.ups_beta * 0.15 * iu;
          sa0 = fma(a0[j] * a0[j], iu2, sa0);
          sa1 = fma(a1[j] * a1[j], iu2, sa1);
          sa2 = fma(a2[j] * a2[j], iu2, sa2);
          sW[k] = d0 * a0[j] * iu2;
          sW[m.ws + k] = d1 * a1[j] * iu2;
          sW[2 * m.ws + k] = d2 * a2[j] * iu2;
          // dups_j = 0.5 - 0.25 (ups_j + ups_{j+2}) / ups_{j+1},  j = 0..K-3   (Series_modelcode.txt:51-53)
          double gu = q2 * iu2 * iu - iu;
          if (k + 2 < K) {  // k is the left point of dups_k
            const double e = 0.5 - 0.25 * (ups + sUps[k + 2]) * sIu[k + 1];
            gu += e * 0.25 * sIu[k + 1];
            lp += -0.5 * e * e;
          }
          if (k >= 1 && k + 1 < K) {  // middle point of dups_{k-1}
            const double sum = sUps[k - 1] + sUps[k + 1];
            const double e = 0.5 - 0.25 * sum * iu;
            gu -= e * 0.25 * sum * iu2;
          }
          if (k >= 2) {  // right point of dups_{k-2}
            const double e = 0.5 - 0.25 * (sUps[k - 2] + ups) * sIu[k - 1];
            gu += e * 0.25 * sIu[k - 1];
          }
          grad[m.off_ups + k] = gu * ups - (m.ups_alpha + 1.0) + m.ups_beta * 0.15 * iu + jac;
        }
      }
    }
    __syncwarp();
    // 1c. prior part of d lp / d x:  - sum_j L_j^T W_j
    for (int kb = 0; kb < K; kb += 128) {
      double acc[4] = {0, 0, 0, 0};
      if (m.toeplitz) {
        double b1[4] = {0, 0, 0, 0}, b2[4] = {0, 0, 0, 0};
        const double* w0 = sW + kb + lane;  // lanes past K read the (zero / neighbouring) scratch: never stored
        for (int d = -bw; d <= bw; ++d) {   // row n = k + d, column k -> tap_j[-d]
          const double t0 = sTap[-d], t1 = sTap[LBW - d], t2 = sTap[2 * LBW - d];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int kc = (kb + lane + 32 * j < K) ? 32 * j : 0;
            acc[j] = fma(t0, w0[kc + d], acc[j]);
            b1[j] = fma(t1, w0[m.ws + kc + d], b1[j]);
            b2[j] = fma(t2, w0[2 * m.ws + kc + d], b2[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += b1[j] + b2[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = kb + lane + 32 * j;
          if (k < K) {
            const int dlo = (k - bw < 0) ? -k : -bw, dhi = (k + bw > K - 1) ? (K - 1 - k) : bw;
            for (int d = dlo; d <= dhi; ++d) {
              const double* l0 = m.Lb + (long long)(k + d) * LBW + MAXBW - d;
              acc[j] = fma(__ldg(l0), sW[k + d], acc[j]);
              acc[j] = fma(__ldg(l0 + (long long)K * LBW), sW[m.ws + k + d], acc[j]);
              acc[j] = fma(__ldg(l0 + 2LL * K * LBW), sW[2 * m.ws + k + d], acc[j]);
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kb + lane + 32 * j;
        if (k < K) grad[2 + k] = -acc[j];
      }
    }
    sa0 = warp_sum(sa0);
    sa1 = warp_sum(sa1);
    sa2 = warp_sum(sa2);
    if (lane == 0) {
      // d_j ~ inv_gamma(5, 5): -6 log d - 5/d ; half-normal priors on the six scalar raw parameters
      lp += -6.0 * (u[m.off_d] + u[m.off_d + 1] + u[m.off_d + 2]) - 5.0 / d0 - 5.0 / d1 - 5.0 / d2;
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) ss = fma(sTh[i], sTh[i], ss);
      lp += -0.5 * ss;
      grad[m.off_d] = -0.5 * sa0 * d0 - 6.0 + 5.0 / d0 + jac;
      grad[m.off_d + 1] = -0.5 * sa1 * d1 - 6.0 + 5.0 / d1 + jac;
      grad[m.off_d + 2] = -0.5 * sa2 * d2 - 6.0 + 5.0 / d2 + jac;
    }
    if (jacobian) lp += ujac;
  } else {
    const int kend = (m.kpad4 > K + bw) ? m.kpad4 : K + bw;
    for (int k = lane; k < kend; k += 32) sX[k] = 0.0;
  }
  cta_sync();
  if (snap) *snap = *nact;

  // ---------------------------------------------------------------- phase 2: Zhat = A X on the FP64 tensor cores
  const int g = lane >> 2, t = lane & 3;
  {
    const double* bp = sm + m.oXV + g * m.ldxv + m.xoff + t;
    double* zg = sm + m.oZG;
    const int nmt = m.n2p >> 3;
    for (int mt = warp; mt < nmt; mt += 2 * NWARP) {
      const int mt2 = mt + NWARP;
      const bool two = mt2 < nmt;
      const double *a0p, *a1p;
      if (TOEP) {  // A_p[row][col] = T_p[col - row + nfp - 1]
        const int r0 = mt * 8 + g, r1 = (two ? mt2 : mt) * 8 + g;
        const int p0 = r0 >= m.nfp, p1 = r1 >= m.nfp;
        a0p = sA + p0 * m.lt + (m.nfp - 1) - (r0 - p0 * m.nfp) + t;
        a1p = sA + p1 * m.lt + (m.nfp - 1) - (r1 - p1 * m.nfp) + t;
      } else {
        a0p = sA + (mt * 8 + g) * m.lda + t;
        a1p = sA + ((two ? mt2 : mt) * 8 + g) * m.lda + t;
      }
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
    const double rinf_raw = sTh[0], ind_raw = sTh[1], sr_raw = sTh[2], ap_raw = sTh[3], are_raw = sTh[4],
                 aim_raw = sTh[5];
    const double Rinf = 100.0 * rinf_raw, induc = ind_raw * m.induc_scale;
    const double sr = 0.05 * sr_raw, ap = 0.05 * ap_raw, are = 0.05 * are_raw, aim = 0.05 * aim_raw;
    const double base = m.sigma_min2 + sr * sr;
    const double ap2 = ap * ap, are2 = are * are, aim2 = aim * aim;
    double Sv = 0, Swv = 0, Sg = 0, Sgz = 0, SGre = 0, SGim = 0;
    for (int n = lane; n < Nf; n += 32) {
      const double om = sOm[n];
      const double zre = sZ[n] + Rinf, zim = sZ[m.nfp + n] + induc * om;
      double common = are2 * zre * zre + aim2 * zim * zim;
      double so_raw = 0, so_scale = 0, so = 0;
      if (outl) {
        so_raw = sSo[n];
        so_scale = sSo[Nf + n];
        so = 0.05 * so_raw * so_scale;  // Series_outliers_modelcode.txt:45
        common += so * so;
        // sigma_out_raw ~ exponential(lambda); sigma_out_scale ~ inv_gamma(alpha, beta)
        lp += -m.so_lambda * so_raw - (m.so_alpha + 1.0) * u[m.off_so + Nf + n] - m.so_beta / so_scale;
      }
      const double s_re = base + ap2 * zre * zre + common, s_im = base + ap2 * zim * zim + common;
      const double i_re = 1.0 / s_re, i_im = 1.0 / s_im;
      const double r_re = Zs[n] - zre, r_im = Zs[Nf + n] - zim;
      lp += -0.5 * (r_re * r_re * i_re + r_im * r_im * i_im) - 0.5 * log(s_re * s_im);
      const double g_re = 0.5 * r_re * r_re * i_re * i_re - 0.5 * i_re;
      const double g_im = 0.5 * r_im * r_im * i_im * i_im - 0.5 * i_im;
      const double G = g_re + g_im;
      const double v_re = r_re * i_re + 2.0 * zre * (ap2 * g_re + are2 * G);
      const double v_im = r_im * i_im + 2.0 * zim * (ap2 * g_im + aim2 * G);
      sV[n] = v_re;
      sV[m.nfp + n] = v_im;
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
    for (int n = Nf + lane; n < m.nfp; n += 32) {  // padding rows of both parts
      sV[n] = 0.0;
      sV[m.nfp + n] = 0.0;
    }
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
    for (int n = lane; n < m.n2p; n += 32) sV[n] = 0.0;
  }
  cta_sync();

  // ---------------------------------------------------------------- phase 4: GX = A^T V on the FP64 tensor cores
  {
    const double* bp = sm + m.oXV + g * m.ldxv + m.xoff + t;
    double* zg = sm + m.oZG;
    const int nmt = m.kpad8 >> 3;
    for (int mt = warp; mt < nmt; mt += 2 * NWARP) {
      const int mt2 = mt + NWARP;
      const bool two = mt2 < nmt;
      const int col0 = mt * 8 + g, col1 = (two ? mt2 : mt) * 8 + g;
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
      const int step = TOEP ? -4 : 4 * m.lda;
#pragma unroll
      for (int p = 0; p < 2; ++p) {  // A^T[col][row] = A_p[row][col], rows of part p
        const double* ab = TOEP ? sA + p * m.lt + (m.nfp - 1) - t : sA + (p * m.nfp + t) * m.lda;
        const double *a0p = ab + col0, *a1p = ab + col1;
        const double* bq = bp + p * m.nfp;
#pragma unroll 3
        for (int i0 = 0, ao = 0; i0 < m.nfp; i0 += 4, ao += step) {
          const double b = bq[i0];
          dmma(c00, c01, a0p[ao], b);
          dmma(c10, c11, a1p[ao], b);
        }
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

// Shared-memory plan of a persistent solver kernel: the engine region plus as many of the solver's per-slot work
// vectors (Dpad doubles each, NSLOT slots) as fit.  With Toeplitz-resident A two CTAs share an SM.
struct BdrtPlan {
  int ctas_per_sm, nvec;
  size_t smem;
};
static inline BdrtPlan bdrt_plan(const bdrt_ctx* ctx, const BdrtModel& m, int Dpad, int nvec_max, int head_doubles) {
  BdrtPlan pl;
  pl.ctas_per_sm = m.toepA ? 2 : 1;
  long long budget = ctx->smem_optin;
  if (pl.ctas_per_sm == 2) {
    const long long half = ctx->smem_per_sm / 2 - 1024;  // 1 KB per resident CTA is reserved by the system
    if (half < budget) budget = half;
    if ((long long)(m.oUser + head_doubles) * 8 > budget) {  // engine alone does not fit twice
      pl.ctas_per_sm = 1;
      budget = ctx->smem_optin;
    }
  }
  long long nv = (budget / 8 - m.oUser - head_doubles) / ((long long)NSLOT * Dpad);
  if (nv > nvec_max) nv = nvec_max;
  if (nv < 0) nv = 0;
  pl.nvec = (int)nv;
  pl.smem = ((size_t)m.oUser + head_doubles + (size_t)NSLOT * pl.nvec * Dpad) * sizeof(double);
  return pl;
}

// launch KERNEL<1> (Toeplitz-resident A) or KERNEL<0> (dense-resident A)
#define BDRT_LAUNCH(ctx, m, KERNEL, grid, smem, ...)                                                         \
  do {                                                                                                       \
    if ((m).toepA) {                                                                                         \
      BDRT_CUDA(ctx, cudaFuncSetAttribute(KERNEL<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
      KERNEL<1><<<grid, NTHREADS, smem, (ctx)->stream>>>(__VA_ARGS__);                                       \
    } else {                                                                                                 \
      BDRT_CUDA(ctx, cudaFuncSetAttribute(KERNEL<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
      KERNEL<0><<<grid, NTHREADS, smem, (ctx)->stream>>>(__VA_ARGS__);                                       \
    }                                                                                                        \
    (ctx)->launches++;                                                                                       \
    BDRT_CUDA(ctx, cudaGetLastError());                                                                      \
  } while (0)
