// Fused log-posterior + analytic-gradient engine for the reference's Stan programs
//   'Series' family  (bayes_drt/stan_model_files/Series_modelcode.txt:24-69, Series_pos_modelcode.txt:27,
//                     Series_outliers_modelcode.txt:22-72, Series_pos_outliers_modelcode.txt:25)            ND = 1
//   'Parallel'        (Parallel_modelcode.txt:24-75)                                                         ND = 1
//   'Series-Parallel' (Series-Parallel_modelcode.txt:32-107, Series-Parallel_pos_modelcode.txt:35)         ND = 2
//   'Series-2Parallel' (Series-2Parallel_modelcode.txt:39-129, Series-2Parallel_pos_modelcode.txt:42)      ND = 3
//
// Execution model (B200): persistent CTAs of 8 warps.  A CTA owns NSLOT = 8 "column slots"; slot s is driven by warp s,
// which runs its own copy of the calling algorithm (L-BFGS, NUTS, ...) with ordinary warp-uniform control flow.
// Whenever the algorithms need log p and its gradient they all call engine_eval(): the stacked kernel matrix of every
// distribution stays resident in shared memory for the lifetime of the CTA and the dense products
//     Zhat_d = A_d X_d   (2Nf x K_d) (K_d x 8)     and     G_d = A_d^T V_d   (K_d x 2Nf) (2Nf x 8)
// are done cooperatively by all 8 warps for all 8 slots at once on the FP64 tensor cores
// (mma.sync.m8n8k4.f64: the 8 slots are exactly the N=8 columns of the DMMA tile), while everything that is
// per-slot and O(Nf + K) -- error model, residual weights, banded derivative stencils, hyper-priors, chain rule to
// the unconstrained space -- is done by the slot's own warp with shuffle reductions.  Five CTA barriers per
// evaluation; no global-memory traffic besides the slot's own u / grad vectors and its spectrum Z.
//
// Resident operand, two layouts (template parameter TOEP):
//   dense    A_d [row][lda], lda % 16 in {4, 12}: the A-operand fragment of A (8 rows x 4 cols) and of A^T (4 rows x 8
//            cols) both hit 16 distinct 8-byte banks per half-warp, no swizzle; one CTA per SM (A alone is ~115 kB);
//   Toeplitz when both grids are log-uniform with the same spacing (the reference's own special case,
//            matrices.py:145-242) A_re / A_im only depend on (col - row): two 1-D tables per distribution replace the
//            dense matrix, fragment loads become broadcast table look-ups, and two CTAs fit on an SM.
// Inside the engine the stacked vectors are laid out [re: nfp | im: nfp] with nfp = Nf rounded up to 8, so that neither
// an 8-row tile nor a 4-row k-step straddles the two parts.
//
// Third mode (TOEP == 2, "warp mode"): Toeplitz operands AND warp-private products.  With A_p[r][c] = T_p[c - r] the
// product of ONE slot is itself a small matrix product once the vector is viewed as a Hankel matrix:
//     Zhat_p[n + 8 i] = sum_kk T_p[kk - 8 - 8 i] * x[n + kk - 8]        (M = (part, row group i), N = n in 0..7, K = kk)
//     G[n + 8 i]      = sum_p sum_kk T_p[8 i + 7 - kk] * V_p[n - 7 + kk]  (M = column group i,     N = n,       K = (p, kk))
// so every warp runs mma.sync.m8n8k4.f64 on its own slot (the A fragment is a table look-up shared by all row groups,
// the B fragment a sliding window of the slot's vector) and the CTA-wide rendezvous disappears: no barrier inside an
// evaluation, warps of a CTA are at different phases and the FP64 tensor pipe, the FP64 FMA pipe and the L2 round trips
// of the solvers' own vector code overlap across warps.  Cost: the sliding window pads the K dimension by 7-8 and the
// tiles are not full (Nf = 70, K = 100: 161 DMMA per evaluation and slot against 112.5), on a pipe that is otherwise
// ~15 % busy.  Per-spectrum grids get per-slot tables, so eight different grids share a CTA.
#pragma once
#include "common.cuh"

#define NSLOT 8
#define NWARP 8
#define NTHREADS (NWARP * 32)
#define MAXBW 24
#define LBW (2 * MAXBW + 1)
#define LOG_015 (-1.8971199848858813)  // log(0.15)
#define MAXD 3                          // distributions per model (Series-2Parallel: one series + two parallel)
#define FBW 6                           // stencil half-width of the register-tiled fast path (default epsilon: bw = 6)

#define F_POS 1  // lower=0 coefficients of the series distribution (x = exp(u))
#define F_OUT 2  // outlier error model (Series family only)

#define TPAD 4                          // zero entries in front of a warp-mode table (keeps every fragment load in range)

struct BdrtDist {
  int K, kpad4, kpad8, lda, lt;
  int lt2;                    // warp mode: pitch of one part of the table (% 16 == 4: conflict-free A fragments)
  int oTs;                    // warp mode, per-spectrum grids: offset of the slot's own table inside its scratch
  int off_x, off_ups, off_d;  // offsets of x, ups_raw, d_strength inside the unconstrained vector
  int oA, oTap;               // shared-memory offsets (doubles) of the resident operand and of the stencil taps
  int pos;                    // coefficients are lower=0
  int par;                    // parallel distribution: contributes Z_p = 1 / (A x) (Parallel / Series-Parallel :63-66)
  int toepL;                  // L0/L1/L2 are Toeplitz: use the tap table
  const double* A;            // [2Nf, K] or [B, 2Nf, K]
  long long A_stride;         // per-spectrum stride (0: shared)
  const double* Ag;           // global-dense mode: padded, scaled copy [n2p][lda] (+ 8) per grid, in the workspace
  long long Ag_stride;        // per-spectrum stride of Ag (0: shared)
  const double* Lb;           // [3][K][LBW] banded copies of the scaled L0, L1, L2
  double ascale;              // the resident operand is A * ascale (xp = xp_raw * xp_scale, Series-Parallel :52)
  double tapc[3][2 * FBW + 1]; // Toeplitz taps of L0/L1/L2 for |d| <= FBW, read straight from the parameter bank
};

struct BdrtModel {
  int flags, ND, Nf, N2, D, B;
  BdrtDist d[MAXD];
  int Kmax, ldxv, ldzg;
  int nfp;    // Nf rounded up to a multiple of 8
  int n2p;    // 2 * nfp
  int toepA;  // Toeplitz-resident operands
  int gdense; // dense operands that do not fit in shared memory (two / three distributions on general grids): the products
              // read the padded copies d[].Ag through L1 / L2 instead; oCur holds the CTA's current spectrum
  int oCur;
  int wmode;  // warp mode: Toeplitz operands, warp-private Hankel products (engine_eval<2, ..>), no CTA barriers
  int pslot;  // warp mode with per-spectrum grids: tables and omega live in the slot's scratch
  int wsync;  // warp mode, solver kernels: one CTA barrier at the entry of every evaluation (none inside) keeps the eight
              // warps of a CTA in step through the straight-line engine code, so that they share instruction-cache lines
  int vim;    // offset of the imaginary part inside a V row (nfp; warp mode: nfp + 12, zero gap for the sliding window)
  int xz;     // phase 1 zeroes x[K .. xz)
  int oOmS;   // pslot: offset of the slot's omega [Nf] inside its scratch
#ifdef BDRT_PHASE_CLOCKS
  unsigned long long* dbg_clk;  // [16] per-phase clock sums of all warps (profiling builds only, scripts/gpu_phase_clocks.py)
#endif
  int fast;   // register-tiled per-slot phases: Toeplitz L with bw <= FBW and K <= 128 for every distribution
  int off_err, off_so;
  int bw;
  const double* freq;
  long long f_stride;
  const double* Z;  // [B, N2]
  double sigma_min2, ups_alpha, ups_beta, induc_scale, so_lambda, so_alpha, so_beta, x_sum_invscale;
  // shared-memory carve-up, in doubles
  int oXV, oZG, oSt, oOm, oUser;
  int xoff;  // offset of the data inside an X/V row (= bw: zero margin for the stencils)
  int wm;    // zero margin on both sides of a stencil scratch vector: max(bw, FBW)
  int ws;    // stride of one stencil scratch vector (Kmax + 2 wm, even)
  int kup;   // Kmax rounded up to a multiple of 4
  int sd;    // per-slot, per-distribution scratch: W0 | W1 | W2 (ws each) | ups (kup) | 1/ups (kup)
  int st;    // per-slot scratch size: ND * sd | scalars (16) | sigma_out raw, scale (2 Nf)
};

static inline int bdrt_pad_stride(int n) {  // smallest s >= n with s % 16 in {4, 12}
  int s = n;
  while ((s % 16) != 4 && (s % 16) != 12) ++s;
  return s;
}

// Fills the derived fields of m (everything but the pointers / scalars; needs ND, Nf, d[].K, d[].pos, flags, bw, toepA,
// wmode, pslot).  Returns the doubles of engine shared memory.
static inline int bdrt_model_layout(BdrtModel* m) {
  m->N2 = 2 * m->Nf;
  m->nfp = (m->Nf + 7) / 8 * 8;
  m->n2p = 2 * m->nfp;
  m->xoff = m->bw > FBW ? m->bw : FBW;  // even, so that 4-element windows of a row are 16-byte aligned
  m->xoff += m->xoff & 1;
  if (m->wmode && m->xoff < 8) m->xoff = 8;  // the sliding window of the forward product starts at x[-8]
  m->vim = m->wmode ? m->nfp + 12 : m->nfp;
  m->Kmax = 0;
  int kx = 0, k8 = 0;
  for (int i = 0; i < m->ND; ++i) {
    BdrtDist& d = m->d[i];
    d.kpad4 = (d.K + 3) / 4 * 4;
    d.kpad8 = (d.K + 7) / 8 * 8;
    d.lt = m->nfp + d.kpad8;
    d.lt2 = m->nfp + d.kpad8 + 8;
    while ((d.lt2 & 15) != 4) ++d.lt2;
    d.lda = bdrt_pad_stride(d.K);
    if (d.K > m->Kmax) m->Kmax = d.K;
    int k = d.kpad4 > d.K + m->bw ? d.kpad4 : d.K + m->bw;
    if (m->wmode && k < d.K + 12) k = d.K + 12;            // the forward window reads x up to K + 10
    if (m->fast && k < 128 + FBW + 4) k = 128 + FBW + 4;  // the fast path addresses 4 x 32 entries + window
    if (k > kx) kx = k;
    if (d.kpad8 > k8) k8 = d.kpad8;
  }
  m->xz = kx;
  const int vext = m->wmode ? m->vim + m->nfp + 12 : m->n2p;  // warp mode: zero gap after either part of V
  const int mx = kx > vext ? kx : vext;
  m->ldxv = bdrt_pad_stride(m->xoff + mx);
  m->wm = m->bw > FBW ? m->bw : FBW;
  m->ws = m->Kmax + 2 * m->wm;
  m->ws += m->ws & 1;  // even
  m->kup = (m->Kmax + 3) & ~3;  // a lane's 4 consecutive coefficients never run past the row
  int mz = k8 > m->n2p ? k8 : m->n2p;
  if (m->fast && mz < 128) mz = 128;  // the fast path reads 4 x 32 gradient entries per row
  m->ldzg = mz + 4;  // % 8 == 4
  // The Z / G row of a slot (products of phases 2 and 4, read by phases 3 and 5) lives on top of the stencil scratch of
  // phase 1 (dead by then); phase 1 re-zeroes the margins of its scratch vectors on every evaluation.
  m->sd = 3 * m->ws + 2 * m->kup;
  if (m->sd < m->ldzg) m->sd = m->ldzg;
  m->st = m->ND * m->sd + 16 + ((m->flags & F_OUT) ? 2 * m->Nf : 0);
  m->st += m->st & 1;
  if (m->pslot) {  // per-slot tables and omega
    for (int i = 0; i < m->ND; ++i) {
      m->d[i].oTs = m->st;
      m->st += 2 * m->d[i].lt2;
    }
    m->oOmS = m->st;
    m->st += (m->Nf + 1) & ~1;
  }
  // unconstrained vector, Stan declaration order:
  //   Rinf_raw induc_raw | x_0 .. x_{ND-1} | sigma_res alpha_prop alpha_re alpha_im | [sigma_out_raw sigma_out_scale] |
  //   ups_0 .. ups_{ND-1} | d_0(3) .. d_{ND-1}(3)
  int o = 2;
  for (int i = 0; i < m->ND; ++i) { m->d[i].off_x = o; o += m->d[i].K; }
  m->off_err = o; o += 4;
  m->off_so = o;
  if (m->flags & F_OUT) o += 2 * m->Nf;
  for (int i = 0; i < m->ND; ++i) { m->d[i].off_ups = o; o += m->d[i].K; }
  for (int i = 0; i < m->ND; ++i) { m->d[i].off_d = o; o += 3; }
  m->D = o;
  o = 0;
  for (int i = 0; i < m->ND; ++i) {
    m->d[i].oA = o;
    if (m->wmode) o += m->pslot ? 0 : 2 * m->d[i].lt2;
    else if (!m->gdense) o += m->toepA ? 2 * m->d[i].lt : m->n2p * m->d[i].lda + 8;
  }
  m->oXV = o;  o += m->ND * NSLOT * m->ldxv;
  m->oZG = o;  // (no region of its own: see sd above)
  m->oSt = o;  o += NSLOT * m->st;
  for (int i = 0; i < m->ND; ++i) { m->d[i].oTap = o; o += 3 * LBW; }
  m->oOm = o;  o += m->Nf;
  o = (o + 1) & ~1;
  m->oCur = o;  o += 2;
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
// boundary between two phases of an evaluation: CTA rendezvous for the cooperative products, warp-local in warp mode
template <int TOEP>
__device__ __forceinline__ void phase_sync() {
  if (TOEP == 2) __syncwarp();
  else cta_sync();
}

__device__ __forceinline__ void ld2(const double* p, double& a, double& b) {  // 16-byte aligned shared-memory pair
  const double2 v = *reinterpret_cast<const double2*>(p);
  a = v.x;
  b = v.y;
}
__device__ __forceinline__ void st2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// Branch-free exp for the register-tiled phases: Cody-Waite reduction x = k ln2 + r, |r| <= ln2 / 2, Taylor polynomial of
// degree 13 (truncation 4e-18 relative), scaling by 2^k in two exact steps (so results in the subnormal range are right
// and +-inf / overflow / underflow come out as in libm); NaN propagates.  About 1 ulp, like CUDA's exp(), but without
// its slow-path branch, so the independent calls of an unrolled loop are interleaved by the compiler.
__device__ __forceinline__ double bdrt_exp(double x) {
  const double xc = fmin(fmax(x, -746.0), 710.0);
  const double t = fma(xc, 1.4426950408889634, 6755399441055744.0);  // 1.5 * 2^52: rint(x log2 e) in the low word
  const int ki = __double2loint(t);
  const double kf = t - 6755399441055744.0;
  double r = fma(kf, -6.93147180369123816490e-01, xc);
  r = fma(kf, -1.90821492927058770002e-10, r);
  double q = 1.6059043836821613e-10;        // 1/13!
  q = fma(q, r, 2.08767569878681e-09);      // 1/12!
  q = fma(q, r, 2.505210838544172e-08);     // 1/11!
  q = fma(q, r, 2.755731922398589e-07);     // 1/10!
  q = fma(q, r, 2.7557319223985893e-06);    // 1/9!
  q = fma(q, r, 2.48015873015873e-05);      // 1/8!
  q = fma(q, r, 1.984126984126984e-04);     // 1/7!
  q = fma(q, r, 1.388888888888889e-03);     // 1/6!
  q = fma(q, r, 8.333333333333333e-03);     // 1/5!
  q = fma(q, r, 4.1666666666666664e-02);    // 1/4!
  q = fma(q, r, 1.6666666666666666e-01);    // 1/3!
  q = fma(q, r, 0.5);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const int k1 = ki >> 1, k2 = ki - k1;
  const double s1 = __hiloint2double((1023 + k1) << 20, 0), s2 = __hiloint2double((1023 + k2) << 20, 0);
  const double y = q * s1 * s2;
  return x != x ? x : y;
}
// Branch-free reciprocal: hardware seed (rcp.approx.ftz.f64) + two Newton steps; about 1 ulp for normal arguments.
// 0 and +-inf give NaN instead of +-inf / 0: the engine only divides by variances and scales that are positive and
// finite at every point the solvers can accept, and a non-finite log density is a rejected point either way.
__device__ __forceinline__ double bdrt_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
// Sums of N = 2^n per-lane values over the warp in one transposed butterfly: the first n exchange steps halve the number
// of values a lane carries (it keeps the sums its lane bits select and sends the others), the remaining 5 - n steps are
// plain butterflies.  N + 4 - n... shuffles instead of 5 N.  On return v[0] of lane L is the total of value number
// L >> (5 - n) (identical bits in every lane of that group; the summation order depends on nothing but N).
template <int N>
__device__ __forceinline__ double warp_sum_multi(double (&v)[N], int lane) {
  static_assert(N == 2 || N == 4 || N == 8 || N == 16, "N must be 2, 4, 8 or 16");
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool up = lane & off;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const double send = up ? v[i] : v[i + n / 2], keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// Is coordinate i of the unconstrained vector a lower=0 parameter (theta = exp(u))?
__device__ __forceinline__ bool bdrt_is_exp(const BdrtModel& m, int i) {
  for (int dd = 0; dd < m.ND; ++dd)
    if (!m.d[dd].pos && i >= m.d[dd].off_x && i < m.d[dd].off_x + m.d[dd].K) return false;
  return true;
}

// Warp-mode table of one distribution: T_p[(col - row) + nfp - 1 + TPAD] = A_p[row][col] * ascale, zero elsewhere
// (first row of A_p for col - row >= 0, first column below).  Filled by `nthr` threads (a CTA or one warp).
__device__ inline void engine_fill_table(const BdrtModel& m, const BdrtDist& D, double* sT, const double* gA, int tid,
                                         int nthr) {
  for (int i = tid; i < 2 * D.lt2; i += nthr) {
    const int p = i >= D.lt2, d = i - p * D.lt2 - (m.nfp - 1) - TPAD;
    double v = 0.0;
    if (d >= 0 && d < D.K) v = gA[(long long)p * m.Nf * D.K + d];
    else if (d < 0 && -d < m.Nf) v = gA[((long long)p * m.Nf - d) * D.K];
    sT[i] = v * D.ascale;
  }
}

// Cooperative load of the resident operands.  Called by all threads once (or once per spectrum when the grid is
// per-spectrum and the kernel is not in warp mode); ends with a CTA barrier.
__device__ inline void engine_load(const BdrtModel& m, double* sm, long long spec) {
  const int tid = threadIdx.x;
  for (int dd = 0; dd < m.ND; ++dd) {
    const BdrtDist& D = m.d[dd];
    double* sA = sm + D.oA;
    const double* gA = D.A + spec * D.A_stride;
    if (m.wmode) {
      if (!m.pslot) engine_fill_table(m, D, sA, gA, tid, NTHREADS);
    } else if (m.toepA) {
      // table of part p: T_p[(col - row) + nfp - 1] = A_p[row][col]; first row for col - row >= 0, first column below
      for (int i = tid; i < 2 * D.lt; i += NTHREADS) {
        const int p = i >= D.lt, d = i - p * D.lt - (m.nfp - 1);
        double v = 0.0;
        if (d >= 0 && d < D.K) v = gA[(long long)p * m.Nf * D.K + d];
        else if (d < 0 && -d < m.Nf) v = gA[((long long)p * m.Nf - d) * D.K];
        sA[i] = v * D.ascale;
      }
    } else if (!m.gdense) {
      for (int i = tid; i < m.n2p * D.lda + 8; i += NTHREADS) {
        const int rp = i / D.lda, c = i - rp * D.lda;
        const int p = rp >= m.nfp, r = rp - p * m.nfp;  // padded row -> (part, row)
        sA[i] = (rp < m.n2p && r < m.Nf && c < D.K) ? gA[((long long)p * m.Nf + r) * D.K + c] * D.ascale : 0.0;
      }
    }
    // Toeplitz taps: row K/2 of the banded copies
    for (int i = tid; i < 3 * LBW; i += NTHREADS) {
      const int j = i / LBW, d = i - j * LBW;
      sm[D.oTap + i] = D.Lb[((long long)j * D.K + D.K / 2) * LBW + d];
    }
  }
  for (int i = tid; i < m.ND * NSLOT * m.ldxv; i += NTHREADS) sm[m.oXV + i] = 0.0;
  for (int i = tid; i < NSLOT * m.st; i += NTHREADS) sm[m.oSt + i] = 0.0;
  const double* f = m.freq + spec * m.f_stride;
  for (int i = tid; i < m.Nf; i += NTHREADS) sm[m.oOm + i] = 2.0 * M_PI * f[i];
  if (tid == 0) *reinterpret_cast<long long*>(sm + m.oCur) = spec;
  cta_sync();
}

// Warp mode with per-spectrum grids: the calling warp loads the tables and omega of spectrum `spec` into its own slot.
__device__ inline void engine_load_slot(const BdrtModel& m, double* sm, long long spec) {
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  double* sSt = sm + m.oSt + slot * m.st;
  __syncwarp();
  for (int dd = 0; dd < m.ND; ++dd)
    engine_fill_table(m, m.d[dd], sSt + m.d[dd].oTs, m.d[dd].A + spec * m.d[dd].A_stride, lane, 32);
  const double* f = m.freq + spec * m.f_stride;
  for (int i = lane; i < m.Nf; i += 32) sSt[m.oOmS + i] = 2.0 * M_PI * f[i];
  __syncwarp();
}

// log p(u) and d/du for the slot of the calling warp; all NWARP warps of the CTA must call it together.
//   active : this slot has a point to evaluate (inactive slots only help with the matrix products)
//   u, grad: the slot's D-vectors (generic pointers: shared or global)
//   Zs     : the slot's stacked scaled spectrum [N2] (global)
//   nact/snap: optional CTA-wide "slots still working" counter; *snap receives its value at a point where no warp can
//           be modifying it (between the first and last barrier), so every warp of the CTA reads the same value.
// Returns lp (non-finite lp or gradient entries must be checked by the caller).
// Template parameters: TOEP resident-operand layout; MK model kind (0 Series family, 1 Parallel, 2 Series-Parallel,
// 3 Series-2Parallel: fixes the number of distributions and which of them are parallel at compile time); FAST
// register-tiled per-slot phases.
template <int TOEP, int MK, int FAST>
__device__ __forceinline__ double engine_eval(const BdrtModel& m, double* sm, bool active, const double* u, double* grad,
                                     const double* Zs, int jacobian, const volatile int* nact = nullptr,
                                     int* snap = nullptr) {
  constexpr int ND = MK == 0 ? 1 : MK;
  auto is_par = [](int dd) { return MK == 1 || dd >= 1; };  // distribution dd contributes Z_p = 1 / (A x)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp;
  const int Nf = m.Nf, bw = m.bw, nfp = m.nfp;
  const bool outl = (ND == 1) && (m.flags & F_OUT);
  double* sSt = sm + m.oSt + slot * m.st;
  double* sTh = sSt + ND * m.sd;  // Rinf_raw, induc_raw, sigma_res_raw, alpha_prop/re/im_raw, d_0(3) [, d_1(3)]
  double* sSo = sTh + 16;         // sigma_out_raw [Nf], sigma_out_scale [Nf]
  const double* sOm = (TOEP == 2 && m.pslot) ? sSt + m.oOmS : sm + m.oOm;
  const double jac = jacobian ? 1.0 : 0.0;
  // global-dense operands: the spectrum whose grids the CTA has loaded (engine_load)
  const long long gsp = (TOEP == 0 && m.gdense) ? *reinterpret_cast<const long long*>(sm + m.oCur) : 0;
  if (TOEP == 2) {
    if (m.wsync && snap) {
      // synchronised warp mode: every warp of the CTA calls in every round (finished ones with active = false) and the
      // barrier itself counts the slots that still work -- the same number in every warp
      *snap = __syncthreads_count(active ? 1 : 0);
    }
    if (!active) return 0.0;  // warp mode: nobody else needs this warp
    __syncwarp();
  }
#ifdef BDRT_PHASE_CLOCKS
  long long pc_ = clock64();
#define PCLK(i)                                                                         \
  do {                                                                                  \
    const long long c_ = clock64();                                                     \
    if (lane == 0 && m.dbg_clk) atomicAdd(&m.dbg_clk[i], (unsigned long long)(c_ - pc_)); \
    pc_ = c_;                                                                           \
  } while (0)
#else
#define PCLK(i)
#endif
  // warp-mode table of distribution dd (CTA-shared, or the slot's own with per-spectrum grids)
  auto tabp = [&](int dd) { return m.pslot ? sSt + m.d[dd].oTs : sm + m.d[dd].oA; };
  // per-distribution views of the slot's rows
  auto rowX = [&](int dd) { return sm + m.oXV + (dd * NSLOT + slot) * m.ldxv + m.xoff; };  // x_k at [k], zero margins
  auto rowZ = [&](int dd) { return sSt + dd * m.sd; };  // on top of the phase-1 scratch

  double lp = 0.0;
  double xsum = 0.0;  // sum(xs) + sum(xp_raw)  (Series-Parallel :56)
  double gpr[ND][4];  // fast path: the prior part of d lp / d x of the lane's 4 coefficients, kept until phase 5
  // ---------------------------------------------------------------- phase 1: transforms, priors, stencils (per slot)
  if (active) {
    double ujac = 0.0;
    // 1a. scalars: theta = exp(u)
    if (lane < 6 + 3 * ND) {
      const int i = lane < 2 ? lane : (lane < 6 ? m.off_err + lane - 2 : m.d[0].off_d + lane - 6);
      const double ui = u[i];
      ujac += ui;
      sTh[lane] = bdrt_exp(ui);
    }
    if (outl) {
      for (int i = lane; i < 2 * Nf; i += 32) {
        const double ui = u[m.off_so + i];
        ujac += ui;
        sSo[i] = exp(ui);
      }
    }
    if (FAST) {
      // Register-tiled per-slot phase (Toeplitz L with |d| <= FBW, K <= 128), written branch-free: every lane runs the
      // same straight-line code on clamped indices and masks the results by selects / predicated stores, so that the
      // compiler interleaves the independent exp / reciprocal / FMA chains of the lane's four coefficients (at 4 warps
      // per scheduler the kernel is latency bound: instruction-level parallelism inside the warp is what fills the
      // pipes).  Everything that touches u / grad uses the interleaved ownership k = lane + 32 j (conflict-free 8-byte
      // accesses); the stencils use the tiled ownership kq .. kq+3 with 16-byte shared-memory windows, and the taps
      // come straight from the parameter bank.  Warp sums are deferred and batched (warp_sum_multi).
      const int kq = 4 * lane;
      constexpr int NRED = ND == 1 ? 4 : (ND == 2 ? 8 : 16);
      double red[NRED];
#pragma unroll
      for (int i = 0; i < NRED; ++i) red[i] = 0.0;
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const BdrtDist& Dd = m.d[dd];
        const int K = Dd.K;
        double* sX = rowX(dd);
        double* sW = sSt + dd * m.sd + m.wm;
        double* sUps = sSt + dd * m.sd + 3 * m.ws;
        double* sIu = sUps + m.kup;
        // 1a. transforms; the separable hyper-prior terms are summed here
        {
          double ux[4], uu[4], xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j, kc = k < K ? k : K - 1;
            ux[j] = u[Dd.off_x + kc];
            uu[j] = u[Dd.off_ups + kc];
            xv[j] = ux[j];
          }
          if (Dd.pos) {
#pragma unroll
            for (int j = 0; j < 4; ++j) xv[j] = bdrt_exp(ux[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j;
            const bool v = k < K;
            const double ups = 0.15 * bdrt_exp(uu[j]);
            const double iu = bdrt_rcp(ups);
            sX[k] = v ? xv[j] : 0.0;  // zeros past K (phase 3 of the previous call wrote V here)
            if (v) {
              sUps[k] = ups;
              sIu[k] = iu;
            }
            ujac += v ? uu[j] + (Dd.pos ? ux[j] : 0.0) : 0.0;
            if (ND > 1) xsum += v ? xv[j] : 0.0;
            // - log ups ; ups_raw ~ inv_gamma(alpha, beta)
            lp += v ? -(LOG_015 + uu[j]) - (m.ups_alpha + 1.0) * uu[j] - m.ups_beta * 0.15 * iu : 0.0;
          }
        }
        for (int k = 128 + lane; k < (TOEP == 2 ? m.xz : 128 + FBW + 2); k += 32) sX[k] = 0.0;
        __syncwarp();
        // 1b. a_j = L_j x, q^2, dups, d lp / d ups, W_j = d_j a_j / ups^2.  Lanes past K load the windows of the last
        // tile (so that no lane reads beyond the slot's own rows -- the next slot's scratch belongs to another warp),
        // compute on them and store nothing; every sum is masked by the true index.
        const int kql = kq < Dd.kpad4 ? kq : Dd.kpad4 - 4;
        const double d0 = sTh[6 + 3 * dd], d1 = sTh[7 + 3 * dd], d2 = sTh[8 + 3 * dd];
        double sa0 = 0, sa1 = 0, sa2 = 0;
        double gu4[4];
        {
          double xw[16];  // x[kq - 6 .. kq + 9]
#pragma unroll
          for (int i = 0; i < 8; ++i) ld2(sX + kql - FBW + 2 * i, xw[2 * i], xw[2 * i + 1]);
          double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
#pragma unroll
          for (int t = 0; t < 2 * FBW + 1; ++t) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              a0[j] = fma(Dd.tapc[0][t], xw[j + t], a0[j]);
              a1[j] = fma(Dd.tapc[1][t], xw[j + t], a1[j]);
              a2[j] = fma(Dd.tapc[2][t], xw[j + t], a2[j]);
            }
          }
          double upw[8], iuw[8];  // ups / (1/ups) [kq - 2 .. kq + 5]; own values at index j + 2
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ld2(sUps + kql - 2 + 2 * i, upw[2 * i], upw[2 * i + 1]);
            ld2(sIu + kql - 2 + 2 * i, iuw[2 * i], iuw[2 * i + 1]);
          }
          double w0[4], w1[4], w2[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = kq + j;
            const bool v = k < K;
            const double upk = upw[j + 2], iuk = iuw[j + 2], iu2 = iuk * iuk;
            const double q0 = a0[j] * a0[j] * iu2, q1 = a1[j] * a1[j] * iu2, q2 = a2[j] * a2[j] * iu2;
            const double qq = d0 * q0 + d1 * q1 + d2 * q2;  // q^2 / ups^2
            lp += v ? -0.5 * qq : 0.0;  // q ~ normal(0, ups)
            sa0 += v ? q0 : 0.0;
            sa1 += v ? q1 : 0.0;
            sa2 += v ? q2 : 0.0;
            w0[j] = v ? d0 * a0[j] * iu2 : 0.0;
            w1[j] = v ? d1 * a1[j] * iu2 : 0.0;
            w2[j] = v ? d2 * a2[j] * iu2 : 0.0;
            // dups_i = 0.5 - 0.25 (ups_i + ups_{i+2}) / ups_{i+1}  (Series_modelcode.txt:51-53); window index j + 2 + e
            double gu = qq * iuk - iuk;
            {
              const double e = 0.5 - 0.25 * (upk + upw[j + 4]) * iuw[j + 3];
              const bool c = k + 2 < K;
              gu += c ? e * 0.25 * iuw[j + 3] : 0.0;
              lp += c ? -0.5 * e * e : 0.0;
            }
            {
              const double sum = upw[j + 1] + upw[j + 3];
              const double e = 0.5 - 0.25 * sum * iuk;
              gu -= (k >= 1 && k + 1 < K) ? e * 0.25 * sum * iu2 : 0.0;
            }
            {
              const double e = 0.5 - 0.25 * (upw[j] + upk) * iuw[j + 1];
              gu += (k >= 2 && v) ? e * 0.25 * iuw[j + 1] : 0.0;
            }
            gu4[j] = gu * upk - (m.ups_alpha + 1.0) + m.ups_beta * 0.15 * iuk + jac;
          }
          // margins of the three scratch vectors (the previous evaluation's Z / G row was here); lane 0's ups window
          // starts two entries inside the last margin, so the loads above must be complete in every lane first
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            if (lane < m.wm) sW[q * m.ws - m.wm + lane] = 0.0;
            for (int k = Dd.kpad4 + lane; k < m.ws - m.wm; k += 32) sW[q * m.ws + k] = 0.0;
          }
          if (kq < K) {
            st2(sW + kq, w0[0], w0[1]);  // zeros past K keep the right margin zero
            st2(sW + kq + 2, w0[2], w0[3]);
            st2(sW + m.ws + kq, w1[0], w1[1]);
            st2(sW + m.ws + kq + 2, w1[2], w1[3]);
            st2(sW + 2 * m.ws + kq, w2[0], w2[1]);
            st2(sW + 2 * m.ws + kq + 2, w2[2], w2[3]);
          }
        }
        __syncwarp();  // every lane has its ups / 1/ups windows: the 1/ups row now stages d lp / d u_ups
        if (kq < K) {
          st2(sIu + kq, gu4[0], gu4[1]);
          st2(sIu + kq + 2, gu4[2], gu4[3]);
        }
        // 1c. prior part of d lp / d x:  - sum_j L_j^T W_j  (kept in registers until phase 5); one accumulator set per
        // derivative order: twelve independent chains
        {
          double acc[3][4];
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            double ww[16];  // W_q[kq - 6 .. kq + 9]
#pragma unroll
            for (int i = 0; i < 8; ++i) ld2(sW + q * m.ws + kql - FBW + 2 * i, ww[2 * i], ww[2 * i + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[q][j] = 0.0;
#pragma unroll
            for (int t = 0; t < 2 * FBW + 1; ++t) {  // row n = k + d, column k -> tap_q[-d]
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[q][j] = fma(Dd.tapc[q][2 * FBW - t], ww[j + t], acc[q][j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) gpr[dd][j] = -(acc[0][j] + acc[1][j] + acc[2][j]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // staged d lp / d u_ups -> grad, interleaved ownership
          const int k = lane + 32 * j;
          if (k < K) grad[Dd.off_ups + k] = sIu[k];
        }
        red[3 * dd] = sa0;
        red[3 * dd + 1] = sa1;
        red[3 * dd + 2] = sa2;
      }
      if (ND > 1) red[3 * ND] = xsum;
      // one batched reduction: value i ends up in the lanes with lane >> LSH == i; its first lane writes the gradient
      constexpr int LSH = ND == 1 ? 3 : (ND == 2 ? 2 : 1);
      const double tot = warp_sum_multi<NRED>(red, lane);
      if (ND > 1) xsum = __shfl_sync(0xffffffffu, tot, (3 * ND) << LSH);
      {
        const int i = lane >> LSH;  // i = 3 dd + j
        if (i < 3 * ND && (lane & ((1 << LSH) - 1)) == 0) {
          const int dd = i / 3, j = i - 3 * dd;
          const double dj = sTh[6 + i], idj = bdrt_rcp(dj);
          // d_j ~ inv_gamma(5, 5): -6 log d - 5/d
          lp += -6.0 * u[m.d[dd].off_d + j] - 5.0 * idj;
          grad[m.d[dd].off_d + j] = -0.5 * tot * dj - 6.0 + 5.0 * idj + jac;
        }
      }
    } else {
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const BdrtDist& Dd = m.d[dd];
      const int K = Dd.K;
      double* sX = rowX(dd);
      double* sUps = sSt + dd * m.sd + 3 * m.ws;
      double* sIu = sUps + m.kup;
      for (int k = lane; k < K; k += 32) {
        const double ux = u[Dd.off_x + k], uu = u[Dd.off_ups + k];
        const double xv = Dd.pos ? exp(ux) : ux;
        const double ups = 0.15 * exp(uu);
        sX[k] = xv;
        sUps[k] = ups;
        sIu[k] = __drcp_rn(ups);
        ujac += uu + (Dd.pos ? ux : 0.0);
        if (ND > 1) xsum += xv;
      }
      const int kend = TOEP == 2 ? m.xz : ((Dd.kpad4 > K + bw) ? Dd.kpad4 : K + bw);
      for (int k = K + lane; k < kend; k += 32) sX[k] = 0.0;  // right margin (phase 3 of the previous call wrote V here)
    }
    __syncwarp();
    // 1b. a_j = L_j x (banded), q^2, hyper-priors, d lp / d ups, W_j = d_j a_j / ups^2
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const BdrtDist& Dd = m.d[dd];
      const int K = Dd.K;
      const double* sX = rowX(dd);
      double* sW = sSt + dd * m.sd + m.wm;  // W_j[k] at sW[j * ws + k], zero margins
      const double* sUps = sSt + dd * m.sd + 3 * m.ws;
      const double* sIu = sUps + m.kup;
      const double* sTap = sm + Dd.oTap + MAXBW;  // tap_j[d] at sTap[j * LBW + d]
      const double d0 = sTh[6 + 3 * dd], d1 = sTh[7 + 3 * dd], d2 = sTh[8 + 3 * dd];
      double sa0 = 0, sa1 = 0, sa2 = 0;
      // margins of the three scratch vectors (the previous evaluation's Z / G row was here)
      for (int q = 0; q < 3; ++q) {
        for (int k = lane; k < m.wm; k += 32) sW[q * m.ws - m.wm + k] = 0.0;
        for (int k = K + lane; k < m.ws - m.wm; k += 32) sW[q * m.ws + k] = 0.0;
      }
      for (int kb = 0; kb < K; kb += 128) {
        double a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
        if (Dd.toepL) {
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
              const double* l0 = Dd.Lb + (long long)k * LBW + MAXBW;
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
            const double uk = u[Dd.off_ups + k];
            const double q2 = d0 * a0[j] * a0[j] + d1 * a1[j] * a1[j] + d2 * a2[j] * a2[j];
            // q ~ normal(0, ups): -1/2 q^2/ups^2 - log ups ;  ups_raw ~ inv_gamma(alpha, beta)
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
            grad[Dd.off_ups + k] = gu * ups - (m.ups_alpha + 1.0) + m.ups_beta * 0.15 * iu + jac;
          }
        }
      }
      __syncwarp();
      // 1c. prior part of d lp / d x:  - sum_j L_j^T W_j
      for (int kb = 0; kb < K; kb += 128) {
        double acc[4] = {0, 0, 0, 0};
        if (Dd.toepL) {
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
                const double* l0 = Dd.Lb + (long long)(k + d) * LBW + MAXBW - d;
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
          if (k < K) grad[Dd.off_x + k] = -acc[j];
        }
      }
      sa0 = warp_sum(sa0);
      sa1 = warp_sum(sa1);
      sa2 = warp_sum(sa2);
      if (lane == 0) {
        // d_j ~ inv_gamma(5, 5): -6 log d - 5/d
        lp += -6.0 * (u[Dd.off_d] + u[Dd.off_d + 1] + u[Dd.off_d + 2]) - 5.0 / d0 - 5.0 / d1 - 5.0 / d2;
        grad[Dd.off_d] = -0.5 * sa0 * d0 - 6.0 + 5.0 / d0 + jac;
        grad[Dd.off_d + 1] = -0.5 * sa1 * d1 - 6.0 + 5.0 / d1 + jac;
        grad[Dd.off_d + 2] = -0.5 * sa2 * d2 - 6.0 + 5.0 / d2 + jac;
      }
    }
    }
    if (lane == 0) {  // half-normal priors on the six scalar raw parameters
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < 6; ++i) ss = fma(sTh[i], sTh[i], ss);
      lp += -0.5 * ss;
    }
    if (ND > 1) {
      if (!FAST) xsum = warp_sum(xsum);  // (the register-tiled path has it already: batched reduction above)
      // x_sum = x_sum_raw * x_sum_invscale ~ std_normal(); real<lower=0> x_sum_raw is validity-checked by Stan (:56-57)
      if (lane == 0) lp += (xsum < 0.0) ? -INFINITY : -0.5 * (xsum * m.x_sum_invscale) * (xsum * m.x_sum_invscale);
    }
    if (jacobian) lp += ujac;
  } else {
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      double* sX = rowX(dd);
      const int kend = (m.d[dd].kpad4 > m.d[dd].K + bw) ? m.d[dd].kpad4 : m.d[dd].K + bw;
      for (int k = lane; k < kend; k += 32) sX[k] = 0.0;
    }
  }
  phase_sync<TOEP>();
  PCLK(1);
  if (snap) *snap = *nact;

  // ---------------------------------------------------------------- phase 2: Zhat_d = A_d X_d on the FP64 tensor cores
  const int g = lane >> 2, t = lane & 3;
  if (TOEP == 2) {
    // Warp-private Hankel product: M = (part p, row group i) pairs, N = n (row inside the group), K = kk.
    //   A[(p, i)][kk] = T_p[kk - 8 - 8 i + nfp - 1 + TPAD],   B[kk][n] = x[n + kk - 8],   C[(p, i)][n] = Zhat_p[n + 8 i]
    // Rows g of a tile alternate the parts (p = g & 1) so that a half-warp's A fragment spans 16 distinct banks.
    const int NR8 = nfp >> 3;
    const int p = g & 1;
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const BdrtDist& Dd = m.d[dd];
      const double* sT = tabp(dd) + p * Dd.lt2 + (nfp - 1 + TPAD - 8) + t;
      const double* bp = rowX(dd) + g + t - 8;
      double* sZ = rowZ(dd) + p * nfp + 2 * t;
      const int KKf = (Dd.K + 8 + 3) & ~3;
      for (int q0 = 0; q0 < NR8; q0 += 12) {  // up to three M tiles (12 row groups of both parts) share the B fragments
        const int i0 = q0 + (g >> 1), i1 = i0 + 4, i2 = i0 + 8;
        const bool t1 = q0 + 4 < NR8, t2 = q0 + 8 < NR8;  // warp-uniform
        const double* a0 = sT - 8 * (i0 < NR8 ? i0 : NR8 - 1);
        const double* a1 = sT - 8 * (i1 < NR8 ? i1 : NR8 - 1);
        const double* a2 = sT - 8 * (i2 < NR8 ? i2 : NR8 - 1);
        double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0};  // two accumulation chains per tile
        int kk = 0;
#pragma unroll 2
        for (; kk + 4 < KKf; kk += 8) {
          const double b0 = bp[kk], b1 = bp[kk + 4];
          dmma(c0[0], c0[1], a0[kk], b0);
          dmma(c0[2], c0[3], a0[kk + 4], b1);
          if (t1) {
            dmma(c1[0], c1[1], a1[kk], b0);
            dmma(c1[2], c1[3], a1[kk + 4], b1);
          }
          if (t2) {
            dmma(c2[0], c2[1], a2[kk], b0);
            dmma(c2[2], c2[3], a2[kk + 4], b1);
          }
        }
        if (kk < KKf) {
          const double b0 = bp[kk];
          dmma(c0[0], c0[1], a0[kk], b0);
          if (t1) dmma(c1[0], c1[1], a1[kk], b0);
          if (t2) dmma(c2[0], c2[1], a2[kk], b0);
        }
        if (i0 < NR8) st2(sZ + 8 * i0, c0[0] + c0[2], c0[1] + c0[3]);
        if (t1 && i1 < NR8) st2(sZ + 8 * i1, c1[0] + c1[2], c1[1] + c1[3]);
        if (t2 && i2 < NR8) st2(sZ + 8 * i2, c2[0] + c2[2], c2[1] + c2[3]);
      }
    }
  } else {
    const int nmt = m.n2p >> 3;
    for (int ti = warp; ti < ND * nmt; ti += 2 * NWARP) {
      const int ti2 = ti + NWARP;
      const bool two = ti2 < ND * nmt;
      const int tj = two ? ti2 : ti;
      const int da = (ND > 1) ? ti / nmt : 0, db = (ND > 1) ? tj / nmt : 0;
      const int mt = ti - da * nmt, mt2 = tj - db * nmt;
      const BdrtDist &Da = m.d[da], &Db = m.d[db];
      const double* b0p = sm + m.oXV + (da * NSLOT + g) * m.ldxv + m.xoff + t;
      const double* b1p = sm + m.oXV + (db * NSLOT + g) * m.ldxv + m.xoff + t;
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
      // (a lambda, so that the shared-memory and the global-dense call sites each keep their own address space)
      auto prod = [&](const double* a0p, const double* a1p) {
        if (ND == 1 || (da == db && Da.kpad4 == Db.kpad4)) {
#pragma unroll 5
          for (int kk = 0; kk < Da.kpad4; kk += 4) {
            const double b = b0p[kk];
            dmma(c00, c01, a0p[kk], b);
            dmma(c10, c11, a1p[kk], b);
          }
        } else {
#pragma unroll 5
          for (int kk = 0; kk < Da.kpad4; kk += 4) dmma(c00, c01, a0p[kk], b0p[kk]);
#pragma unroll 5
          for (int kk = 0; kk < Db.kpad4; kk += 4) dmma(c10, c11, a1p[kk], b1p[kk]);
        }
      };
      if (TOEP) {  // A_p[row][col] = T_p[col - row + nfp - 1]
        const int r0 = mt * 8 + g, r1 = mt2 * 8 + g;
        const int p0 = r0 >= nfp, p1 = r1 >= nfp;
        prod(sm + Da.oA + p0 * Da.lt + (nfp - 1) - (r0 - p0 * nfp) + t,
             sm + Db.oA + p1 * Db.lt + (nfp - 1) - (r1 - p1 * nfp) + t);
      } else if (m.gdense) {
        prod(Da.Ag + gsp * Da.Ag_stride + (mt * 8 + g) * Da.lda + t, Db.Ag + gsp * Db.Ag_stride + (mt2 * 8 + g) * Db.lda + t);
      } else {
        prod(sm + Da.oA + (mt * 8 + g) * Da.lda + t, sm + Db.oA + (mt2 * 8 + g) * Db.lda + t);
      }
      double* z0 = sm + m.oSt + da * m.sd;  // column (slot) 2 t, 2 t + 1 -> that slot's Z / G row
      z0[(2 * t) * m.st + mt * 8 + g] = c00;
      z0[(2 * t + 1) * m.st + mt * 8 + g] = c01;
      if (two) {
        double* z1 = sm + m.oSt + db * m.sd;
        z1[(2 * t) * m.st + mt2 * 8 + g] = c10;
        z1[(2 * t + 1) * m.st + mt2 * 8 + g] = c11;
      }
    }
  }
  phase_sync<TOEP>();
  PCLK(2);

  // ---------------------------------------------------------------- phase 3: error model, residual weights (per slot)
  if (active) {
    const double rinf_raw = sTh[0], ind_raw = sTh[1], sr_raw = sTh[2], ap_raw = sTh[3], are_raw = sTh[4],
                 aim_raw = sTh[5];
    const double Rinf = 100.0 * rinf_raw, induc = ind_raw * m.induc_scale;
    const double sr = 0.05 * sr_raw, ap = 0.05 * ap_raw, are = 0.05 * are_raw, aim = 0.05 * aim_raw;
    const double base = m.sigma_min2 + sr * sr;
    const double ap2 = ap * ap, are2 = are * are, aim2 = aim * aim;
    double Sv = 0, Swv = 0, Sg = 0, Sgz = 0, SGre = 0, SGim = 0;
    if (FAST) {
      // Branch-free: three frequencies per lane and pass on clamped indices, results masked by selects / predicated
      // stores, so the three reciprocal chains are interleaved; the log-determinant term takes ONE logarithm per lane
      // and pass, of the product of the (up to six) variances -- each is >= sigma_min^2, far from under/overflow.
      for (int nb = 0; nb < Nf; nb += 96) {
        double prodS = 1.0;
#pragma unroll
        for (int jn = 0; jn < 3; ++jn) {
          const int n = nb + lane + 32 * jn;
          const bool vld = n < Nf;
          const int nc = vld ? n : Nf - 1;
          const double om = sOm[nc];
          double zre = Rinf, zim = induc * om;
          double Yr[ND], Yi[ND], iM[ND];
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) {
            const double* sZd = rowZ(dd);
            if (is_par(dd)) {
              // Z_p = 1 / (Y' + i Y'')  (Parallel_modelcode.txt:46-50, Series-Parallel :63-66)
              Yr[dd] = sZd[nc];
              Yi[dd] = sZd[nfp + nc];
              iM[dd] = bdrt_rcp(Yr[dd] * Yr[dd] + Yi[dd] * Yi[dd]);
              zre += Yr[dd] * iM[dd];
              zim -= Yi[dd] * iM[dd];
            } else {
              zre += sZd[nc];
              zim += sZd[nfp + nc];
            }
          }
          double common = are2 * zre * zre + aim2 * zim * zim;
          double so_raw = 0, so_scale = 0, so = 0;
          if (outl) {
            so_raw = sSo[nc];
            so_scale = sSo[Nf + nc];
            so = 0.05 * so_raw * so_scale;  // Series_outliers_modelcode.txt:45
            common += so * so;
            // sigma_out_raw ~ exponential(lambda); sigma_out_scale ~ inv_gamma(alpha, beta)
            const double t_ = -m.so_lambda * so_raw - (m.so_alpha + 1.0) * u[m.off_so + Nf + nc] -
                              m.so_beta * bdrt_rcp(so_scale);
            lp += vld ? t_ : 0.0;
          }
          const double s_re = base + ap2 * zre * zre + common, s_im = base + ap2 * zim * zim + common;
          const double i_re = bdrt_rcp(s_re), i_im = bdrt_rcp(s_im);
          const double r_re = Zs[nc] - zre, r_im = Zs[Nf + nc] - zim;
          const double e_re = r_re * r_re * i_re, e_im = r_im * r_im * i_im;
          lp += vld ? -0.5 * (e_re + e_im) : 0.0;
          prodS *= vld ? s_re * s_im : 1.0;
          const double g_re = 0.5 * (e_re - 1.0) * i_re, g_im = 0.5 * (e_im - 1.0) * i_im;
          const double G = g_re + g_im;
          const double v_re = r_re * i_re + 2.0 * zre * (ap2 * g_re + are2 * G);
          const double v_im = r_im * i_im + 2.0 * zim * (ap2 * g_im + aim2 * G);
#pragma unroll
          for (int dd = 0; dd < ND; ++dd) {
            double* sVd = rowX(dd);
            double o_re = v_re, o_im = v_im;
            if (is_par(dd)) {  // d lp / d Y
              const double i2 = iM[dd] * iM[dd];
              const double c1 = (Yi[dd] * Yi[dd] - Yr[dd] * Yr[dd]) * i2, c2 = 2.0 * Yr[dd] * Yi[dd] * i2;
              o_re = v_re * c1 + v_im * c2;
              o_im = -v_re * c2 + v_im * c1;
            }
            if (vld) {
              sVd[n] = o_re;
              sVd[m.vim + n] = o_im;
            }
          }
          Sv += vld ? v_re : 0.0;
          Swv += vld ? om * v_im : 0.0;
          Sg += vld ? G : 0.0;
          Sgz += vld ? g_re * zre * zre + g_im * zim * zim : 0.0;
          SGre += vld ? G * zre * zre : 0.0;
          SGim += vld ? G * zim * zim : 0.0;
          if (outl && vld) {
            const double dso = 2.0 * so * G * so;  // (d lp/d sigma_out) * sigma_out ; sigma_out = .05 raw scale
            grad[m.off_so + n] = dso - m.so_lambda * so_raw + jac;
            grad[m.off_so + Nf + n] = dso - (m.so_alpha + 1.0) + m.so_beta * bdrt_rcp(so_scale) + jac;
          }
        }
        lp -= 0.5 * log(prodS);
      }
    } else
    // three independent frequencies per lane and pass (instruction-level parallelism: the body is a long chain of
    // reciprocals / a logarithm)
    for (int nb = 0; nb < Nf; nb += 96) {
#pragma unroll
     for (int jn = 0; jn < 3; ++jn) {
      const int n = nb + lane + 32 * jn;
      if (n >= Nf) continue;
      const double om = sOm[n];
      double zre = Rinf, zim = induc * om;
      double Yr[ND], Yi[ND], iM[ND];
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        const double* sZd = rowZ(dd);
        if (is_par(dd)) {
          // Z_p = 1 / (Y' + i Y'')  (Parallel_modelcode.txt:46-50, Series-Parallel :63-66)
          Yr[dd] = sZd[n];
          Yi[dd] = sZd[nfp + n];
          iM[dd] = __drcp_rn(Yr[dd] * Yr[dd] + Yi[dd] * Yi[dd]);
          zre += Yr[dd] * iM[dd];
          zim -= Yi[dd] * iM[dd];
        } else {
          zre += sZd[n];
          zim += sZd[nfp + n];
        }
      }
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
      const double i_re = __drcp_rn(s_re), i_im = __drcp_rn(s_im);
      const double r_re = Zs[n] - zre, r_im = Zs[Nf + n] - zim;
      lp += -0.5 * (r_re * r_re * i_re + r_im * r_im * i_im) - 0.5 * log(s_re * s_im);
      const double g_re = 0.5 * r_re * r_re * i_re * i_re - 0.5 * i_re;
      const double g_im = 0.5 * r_im * r_im * i_im * i_im - 0.5 * i_im;
      const double G = g_re + g_im;
      const double v_re = r_re * i_re + 2.0 * zre * (ap2 * g_re + are2 * G);
      const double v_im = r_im * i_im + 2.0 * zim * (ap2 * g_im + aim2 * G);
#pragma unroll
      for (int dd = 0; dd < ND; ++dd) {
        double* sVd = rowX(dd);
        if (is_par(dd)) {  // d lp / d Y
          const double i2 = iM[dd] * iM[dd];
          const double c1 = (Yi[dd] * Yi[dd] - Yr[dd] * Yr[dd]) * i2, c2 = 2.0 * Yr[dd] * Yi[dd] * i2;
          sVd[n] = v_re * c1 + v_im * c2;
          sVd[m.vim + n] = -v_re * c2 + v_im * c1;
        } else {
          sVd[n] = v_re;
          sVd[m.vim + n] = v_im;
        }
      }
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
    }
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      double* sVd = rowX(dd);
      // padding rows of both parts (warp mode: also the zero gap behind either part that the sliding window reads;
      // phase 1 wrote x there)
      for (int n = Nf + lane; n < m.vim; n += 32) {
        sVd[n] = 0.0;
        sVd[m.vim + n] = 0.0;
      }
    }
    if (FAST) {
      // batched reduction of the six scalar sums and of lp (complete here: phases 4 and 5 add nothing to it); value i
      // ends up in lanes 4 i .. 4 i + 3, lane 4 i writes its gradient entry
      double red[8] = {Sv, Swv, Sg, Sgz, SGre, SGim, lp, 0.0};
      const double tot = warp_sum_multi<8>(red, lane);
      lp = __shfl_sync(0xffffffffu, tot, 24);
      if ((lane & 3) == 0 && lane < 24) {
        const int i = lane >> 2;
        const double sc = i == 0 ? 100.0 : (i == 1 ? m.induc_scale : 0.1 * (i == 2 ? sr : (i == 3 ? ap : (i == 4 ? are : aim))));
        const double raw = sTh[i];
        grad[i < 2 ? i : m.off_err + i - 2] = (sc * tot - raw) * raw + jac;
      }
    } else {
    Sv = warp_sum(Sv);
    Swv = warp_sum(Swv);
    Sg = warp_sum(Sg);
    Sgz = warp_sum(Sgz);
    SGre = warp_sum(SGre);
    SGim = warp_sum(SGim);
    if (lane == 0) {
      grad[0] = (100.0 * Sv - rinf_raw) * rinf_raw + jac;
      grad[1] = (m.induc_scale * Swv - ind_raw) * ind_raw + jac;
      grad[m.off_err] = (0.1 * sr * Sg - sr_raw) * sr_raw + jac;
      grad[m.off_err + 1] = (0.1 * ap * Sgz - ap_raw) * ap_raw + jac;
      grad[m.off_err + 2] = (0.1 * are * SGre - are_raw) * are_raw + jac;
      grad[m.off_err + 3] = (0.1 * aim * SGim - aim_raw) * aim_raw + jac;
    }
    }
  } else {
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      double* sVd = rowX(dd);
      for (int n = lane; n < m.n2p; n += 32) sVd[n] = 0.0;
    }
  }
  phase_sync<TOEP>();
  PCLK(3);

  // ---------------------------------------------------------------- phase 4: G_d = A_d^T V_d on the FP64 tensor cores
  if (TOEP == 2) {
    // Warp-private Hankel product: M = column group i, N = n (column inside the group), K = (part p, kk).
    //   A[i][(p, kk)] = T_p[8 i + 7 - kk + nfp - 1 + TPAD],   B[(p, kk)][n] = V_p[n - 7 + kk],   C[i][n] = G[n + 8 i]
    const int KKt = (Nf + 7 + 3) & ~3;
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const BdrtDist& Dd = m.d[dd];
      const int K8 = Dd.kpad8 >> 3, lt2 = Dd.lt2;
      const double* sT = tabp(dd) + (nfp + 6 + TPAD) - t;
      const double* bp = rowX(dd) + g - 7 + t;
      const double* bq = bp + m.vim;
      double* sG = rowZ(dd) + 2 * t;
      for (int q0 = 0; q0 < K8; q0 += 16) {  // two M tiles share the B fragments; one chain per (tile, part)
        const int i0 = q0 + g, i1 = i0 + 8;
        const bool t1 = q0 + 8 < K8;  // warp-uniform
        const double* a0 = sT + 8 * (i0 < K8 ? i0 : K8 - 1);
        const double* a1 = sT + 8 * (i1 < K8 ? i1 : K8 - 1);
        double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
#pragma unroll 2
        for (int kk = 0; kk < KKt; kk += 4) {
          const double b0 = bp[kk], b1 = bq[kk];
          dmma(c0[0], c0[1], a0[-kk], b0);
          dmma(c0[2], c0[3], a0[lt2 - kk], b1);
          if (t1) {
            dmma(c1[0], c1[1], a1[-kk], b0);
            dmma(c1[2], c1[3], a1[lt2 - kk], b1);
          }
        }
        if (i0 < K8) st2(sG + 8 * i0, c0[0] + c0[2], c0[1] + c0[3]);
        if (t1 && i1 < K8) st2(sG + 8 * i1, c1[0] + c1[2], c1[1] + c1[3]);
      }
    }
  } else {
    int nmt_d[MAXD], nmt_tot = 0;
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      nmt_d[dd] = m.d[dd].kpad8 >> 3;
      nmt_tot += nmt_d[dd];
    }
    for (int ti = warp; ti < nmt_tot; ti += 2 * NWARP) {
      const int ti2 = ti + NWARP;
      const bool two = ti2 < nmt_tot;
      const int tj = two ? ti2 : ti;
      int da = 0, db = 0, mt = ti, mt2 = tj;  // tile index -> (distribution, tile of that distribution)
#pragma unroll
      for (int dd = 0; dd + 1 < ND; ++dd) {
        if (da == dd && mt >= nmt_d[dd]) { mt -= nmt_d[dd]; da = dd + 1; }
        if (db == dd && mt2 >= nmt_d[dd]) { mt2 -= nmt_d[dd]; db = dd + 1; }
      }
      const BdrtDist &Da = m.d[da], &Db = m.d[db];
      const int col0 = mt * 8 + g, col1 = mt2 * 8 + g;
      const double* b0p = sm + m.oXV + (da * NSLOT + g) * m.ldxv + m.xoff + t;
      const double* b1p = sm + m.oXV + (db * NSLOT + g) * m.ldxv + m.xoff + t;
      double c00 = 0, c01 = 0, c10 = 0, c11 = 0;
      const int step0 = TOEP ? -4 : 4 * Da.lda, step1 = TOEP ? -4 : 4 * Db.lda;
#pragma unroll
      for (int p = 0; p < 2; ++p) {  // A^T[col][row] = A_p[row][col], rows of part p
        const double* bq0 = b0p + p * nfp;
        const double* bq1 = b1p + p * nfp;
        auto prod = [&](const double* a0p, const double* a1p) {
          if (ND == 1) {
#pragma unroll 3
            for (int i0 = 0, ao = 0; i0 < nfp; i0 += 4, ao += step0) {
              const double b = bq0[i0];
              dmma(c00, c01, a0p[ao], b);
              dmma(c10, c11, a1p[ao], b);
            }
          } else {
#pragma unroll 3
            for (int i0 = 0, ao0 = 0, ao1 = 0; i0 < nfp; i0 += 4, ao0 += step0, ao1 += step1) {
              dmma(c00, c01, a0p[ao0], bq0[i0]);
              dmma(c10, c11, a1p[ao1], bq1[i0]);
            }
          }
        };
        if (TOEP)
          prod(sm + Da.oA + p * Da.lt + (nfp - 1) - t + col0, sm + Db.oA + p * Db.lt + (nfp - 1) - t + col1);
        else if (m.gdense)
          prod(Da.Ag + gsp * Da.Ag_stride + (p * nfp + t) * Da.lda + col0,
               Db.Ag + gsp * Db.Ag_stride + (p * nfp + t) * Db.lda + col1);
        else
          prod(sm + Da.oA + (p * nfp + t) * Da.lda + col0, sm + Db.oA + (p * nfp + t) * Db.lda + col1);
      }
      double* z0 = sm + m.oSt + da * m.sd;  // column (slot) 2 t, 2 t + 1 -> that slot's Z / G row
      z0[(2 * t) * m.st + mt * 8 + g] = c00;
      z0[(2 * t + 1) * m.st + mt * 8 + g] = c01;
      if (two) {
        double* z1 = sm + m.oSt + db * m.sd;
        z1[(2 * t) * m.st + mt2 * 8 + g] = c10;
        z1[(2 * t + 1) * m.st + mt2 * 8 + g] = c11;
      }
    }
  }
  phase_sync<TOEP>();
  PCLK(4);

  // ---------------------------------------------------------------- phase 5: assemble d lp / d u_x (per slot)
  if (active) {
    const double gsum = (ND > 1) ? xsum * m.x_sum_invscale * m.x_sum_invscale : 0.0;
#pragma unroll
    for (int dd = 0; dd < ND; ++dd) {
      const BdrtDist& Dd = m.d[dd];
      const double* sG = rowZ(dd);
      if (FAST) {
        // tiled ownership: G + prior part back into the row (16-byte accesses); then interleaved ownership -> grad
        const int kq = 4 * lane;
        double* sGw = rowZ(dd);
        if (kq < Dd.K) {
          double g4[4];
          ld2(sGw + kq, g4[0], g4[1]);
          ld2(sGw + kq + 2, g4[2], g4[3]);
#pragma unroll
          for (int j = 0; j < 4; ++j) g4[j] += gpr[dd][j] - gsum;
          st2(sGw + kq, g4[0], g4[1]);
          st2(sGw + kq + 2, g4[2], g4[3]);
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = lane + 32 * j;
          if (k < Dd.K) {
            double gx = sGw[k];
            if (Dd.pos) gx = gx * exp(u[Dd.off_x + k]) + jac;
            grad[Dd.off_x + k] = gx;
          }
        }
        continue;
      }
      for (int k = lane; k < Dd.K; k += 32) {
        double gx = sG[k] + grad[Dd.off_x + k] - gsum;
        if (Dd.pos) gx = gx * exp(u[Dd.off_x + k]) + jac;
        grad[Dd.off_x + k] = gx;
      }
    }
    if (!FAST) lp = warp_sum(lp);  // (register-tiled path: reduced with the sums of phase 3)
    __syncwarp();
  }
  PCLK(5);
#ifdef BDRT_PHASE_CLOCKS
  if (lane == 0 && m.dbg_clk) atomicAdd(&m.dbg_clk[0], 1ull);
#endif
  return lp;
}
#endif  // __CUDACC__

// host side (model.cu)
// allow_wmode: 0 the calling kernel only has the cooperative instantiations (engine_eval<0 / 1, ..>: Newton);
// 1 it has all three (the log_prob test hook: warp mode whenever the operands are Toeplitz, BDRT_COOP=1 in the
// environment selects the cooperative Toeplitz products instead, so that the tests cover them);
// 2 it has warp mode and the dense layout only (the NUTS driver: BDRT_COOP is ignored);
// 3 it has all three and prefers the cooperative products when the grid is shared (the L-BFGS driver: see the note in
//   bdrt_model_prepare; per-spectrum grids use warp mode, eight spectra with their own tables per CTA)
int bdrt_model_prepare(bdrt_ctx* ctx, const bdrt_series_data* data, BdrtModel* m, size_t extra_ws_bytes,
                       void** extra_ws, int allow_wmode = 0);

// Shared-memory plan of a persistent solver kernel: the engine region plus as many of the solver's per-slot work
// vectors (Dpad doubles each, NSLOT slots) as fit.  With Toeplitz-resident operands two CTAs share an SM.
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

// launch KERNEL<TOEP, MK, FAST>: Toeplitz- or dense-resident operands, model kind, register-tiled or generic per-slot
// phases (FAST only exists with TOEP)
#define BDRT_LAUNCH_ONE(ctx, KERNEL, T, N, F, grid, smem, ...)                                                       \
  do {                                                                                                               \
    BDRT_CUDA(ctx, cudaFuncSetAttribute(KERNEL<T, N, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem))); \
    KERNEL<T, N, F><<<grid, NTHREADS, smem, (ctx)->stream>>>(__VA_ARGS__);                                           \
  } while (0)
#define BDRT_LAUNCH_ND(ctx, m, KERNEL, T, F, grid, smem, ...)                                \
  do {                                                                                       \
    if ((m).ND == 1 && !(m).d[0].par) BDRT_LAUNCH_ONE(ctx, KERNEL, T, 0, F, grid, smem, __VA_ARGS__); \
    else if ((m).ND == 1) BDRT_LAUNCH_ONE(ctx, KERNEL, T, 1, F, grid, smem, __VA_ARGS__);    \
    else if ((m).ND == 2) BDRT_LAUNCH_ONE(ctx, KERNEL, T, 2, F, grid, smem, __VA_ARGS__);    \
    else BDRT_LAUNCH_ONE(ctx, KERNEL, T, 3, F, grid, smem, __VA_ARGS__);                     \
  } while (0)
// BDRT_LAUNCH: every layout (model_prepare(.., allow_wmode = 1)); BDRT_LAUNCH_COOP: cooperative products only
// (allow_wmode = 0); BDRT_LAUNCH_SOLVER: warp mode or dense (allow_wmode = 2)
#define BDRT_LAUNCH_WARP_(ctx, m, KERNEL, grid, smem, ...)                                    \
    if ((m).wmode && (m).fast) BDRT_LAUNCH_ND(ctx, m, KERNEL, 2, 1, grid, smem, __VA_ARGS__); \
    else if ((m).wmode) BDRT_LAUNCH_ND(ctx, m, KERNEL, 2, 0, grid, smem, __VA_ARGS__);
#define BDRT_LAUNCH_COOP_(ctx, m, KERNEL, grid, smem, ...)                                         \
    if ((m).toepA && (m).fast) BDRT_LAUNCH_ND(ctx, m, KERNEL, 1, 1, grid, smem, __VA_ARGS__);      \
    else if ((m).toepA) BDRT_LAUNCH_ND(ctx, m, KERNEL, 1, 0, grid, smem, __VA_ARGS__);
#define BDRT_LAUNCH_END_(ctx, m, KERNEL, grid, smem, ...)                                     \
    else BDRT_LAUNCH_ND(ctx, m, KERNEL, 0, 0, grid, smem, __VA_ARGS__);                       \
    (ctx)->launches++;                                                                        \
    BDRT_CUDA(ctx, cudaGetLastError());
#define BDRT_LAUNCH(ctx, m, KERNEL, grid, smem, ...)                      \
  do {                                                                    \
    BDRT_LAUNCH_WARP_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)            \
    else BDRT_LAUNCH_COOP_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)       \
    BDRT_LAUNCH_END_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)             \
  } while (0)
#define BDRT_LAUNCH_COOP(ctx, m, KERNEL, grid, smem, ...)                 \
  do {                                                                    \
    BDRT_LAUNCH_COOP_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)            \
    BDRT_LAUNCH_END_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)             \
  } while (0)
#define BDRT_LAUNCH_SOLVER(ctx, m, KERNEL, grid, smem, ...)               \
  do {                                                                    \
    BDRT_LAUNCH_WARP_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)            \
    BDRT_LAUNCH_END_(ctx, m, KERNEL, grid, smem, __VA_ARGS__)             \
  } while (0)
