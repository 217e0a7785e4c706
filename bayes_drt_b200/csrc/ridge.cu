// Batched hyper-parametric ridge fit: bound-constrained QP + hyper-lambda fixed-point loop.
//
// Replaces cvxopt.solvers.qp inside Inverter._convex_opt (bayes_drt/inversion.py:1043-1067) and the hyper-lambda
// loop of Inverter.ridge_fit (inversion.py:489-753; lambda updates :947-954 and :973-983).
//
// The QP  min 1/2 c'Pc + q'c  s.t. c >= lb  (P = WA'WA + sum_o frac_o Lam_o^1/2 Pen_o Lam_o^1/2, strictly convex, simple
// bounds) is solved EXACTLY by block principal pivoting: every pivot step is one dense Cholesky solve on the free set.
// Mapping: one CTA (8 warps) per spectrum, persistent, grid-stride; the working matrix (n = K + 2 padded to a multiple
// of 8, row stride % 16 in {4, 12}: conflict-free DMMA fragments) lives in shared memory, two CTAs per SM; bound
// variables stay in the system as identity rows / columns so the factorisation never changes size.
//
// Linear algebra (round 2): right-looking blocked Cholesky with 8 x 8 blocks --
//   diagonal block: one warp, rows in registers (lane r holds row r), eight fully unrolled steps with shuffles, then the
//                   inverse of the 8 x 8 factor by forward substitution (kept: the triangular solves reuse it);
//   panel:          L21 = A21 L11^-T as 8 x 8 x 8 products on the FP64 tensor cores (mma.sync.m8n8k4.f64);
//   trailing:       A22 -= L21 L21^T tile by tile on the FP64 tensor cores, lower tiles dealt round-robin to the 8 warps;
//   solves:         per block an 8 x 8 mat-vec with the stored inverse and a row-parallel update of the remaining rows.
// The Gram matrix WA'WA is built on the tensor cores as well, streaming WA from global memory (any Nf).
#include "common.cuh"

#ifndef RIDGE_RT
#define RIDGE_RT 256
#endif
#define RT RIDGE_RT  // threads per CTA
#define RW (RT / 32)

struct RidgeArgs {
  bdrt_ridge_opts o;
  const double *WA_re, *WA_im, *WZ_re, *WZ_im, *Pen, *Lmat;
  long long wa_stride;  // 0: shared WA
  int B, Nf, K, n, np, ld;
  int p_smem;  // the system matrix P lives in shared memory next to the working matrix (else in the global scratch)
  double* coef;
  double* lam;
  int* iters;
  int* converged;
  int* pivots;      // [B] Cholesky factorisations spent on the spectrum (may be NULL)
  double* scratch;  // per CTA: G0 [n*n] | Pg [n*n]
  unsigned long long* dbg;  // profiling builds (-DBDRT_PHASE_CLOCKS): [16] clock sums of thread 0 of every CTA
};

#ifdef BDRT_PHASE_CLOCKS
#define RCLK(i)                                                                 \
  do {                                                                          \
    const long long c_ = clock64();                                             \
    if (threadIdx.x == 0 && rdbg) atomicAdd(&rdbg[i], (unsigned long long)(c_ - rc_)); \
    rc_ = c_;                                                                   \
  } while (0)
#else
#define RCLK(i)
#endif

namespace {

__device__ __forceinline__ void rdmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < RW; ++i) s += red[i];
  return s;
}
__device__ double block_max(double v, double* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = red[0];
  for (int i = 1; i < RW; ++i) s = fmax(s, red[i]);
  return s;
}

// sums of two per-lane values in one butterfly: value 0 ends up in lanes 0..15, value 1 in lanes 16..31
__device__ __forceinline__ double warp_sum_multi2(double (&v)[2], int lane) {
  const bool up = lane & 16;
  const double send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
  double t = keep + __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;
}

// 1 / sqrt(d): hardware seed + two Newton steps (about 1 ulp for normal d > 0)
__device__ __forceinline__ double rsqrt_nr(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-d * y, y, 1.0);
  return fma(0.5 * y, e, y);
}

// Cholesky factor (lower) of the 8 x 8 block at D (row stride ld) and its inverse Li [8][8] (row-major, lower), by one
// warp.  Lane r < 8 holds row r in registers; a non-positive pivot is replaced by a tiny positive number (flagged).
__device__ void chol8_inv(double* D, int ld, double* Li, int lane, int* flag) {
  const int r = lane & 7;
  double row[8], isv[8];  // isv: 1 / L[j][j] (the same in every lane)
#pragma unroll
  for (int c = 0; c < 8; ++c) row[c] = D[r * ld + c];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    double d = __shfl_sync(0xffffffffu, row[j], j);  // pivot
    if (!(d > 0.0)) { d = 1e-300; if (lane == 0) *flag = 1; }
    const double is = rsqrt_nr(d);
    isv[j] = is;
    row[j] = (r == j) ? d * is : row[j] * is;  // column j of L (rows above the diagonal hold garbage, never used)
#pragma unroll
    for (int c = j + 1; c < 8; ++c) {
      const double lcj = __shfl_sync(0xffffffffu, row[j], c);  // L[c][j]
      row[c] = fma(-row[j], lcj, row[c]);
    }
  }
  __syncwarp();  // (lanes 8 .. 31 read the same rows as lanes 0 .. 7: their loads come before the write-back)
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) D[r * ld + c] = (c <= r) ? row[c] : 0.0;
  }
  __syncwarp();
  // inverse: lane c < 8 computes column c of X = L^-1 by forward substitution (L read as shared-memory broadcasts)
  if (lane < 8) {
    const int c = lane;
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int p2 = 0; p2 < i; ++p2) acc = fma(-D[i * ld + p2], x[p2], acc);
      x[i] = (i >= c) ? acc * isv[i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Li[i * 8 + c] = x[i];
  }
  __syncwarp();
}

// In-place blocked Cholesky (lower) of the np x np matrix M (row stride ld, np % 8 == 0) by the whole CTA; Linv receives
// the inverses of the diagonal blocks ([np / 8][64]).  Look-ahead: in the trailing update of block column kb warp 0
// takes the next diagonal tile first and factors it at once, while the other warps work through the remaining tiles, so
// the serial 8 x 8 factorisations hide behind the tensor-core updates.
__device__ void chol_blocked(double* M, int np, int ld, double* Linv, int* flag) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int NB = np >> 3;
  if (warp == 0) chol8_inv(M, ld, Linv, lane, flag);
  __syncthreads();
  for (int kb = 0; kb < NB; ++kb) {
    const int k0 = kb << 3;
    const double* Li = Linv + kb * 64;
    // panel: X = A21 Li^T   (rows i0 .. i0 + 7 of the panel per tile; one tile per warp and pass)
    for (int ib = kb + 1 + warp; ib < NB; ib += RW) {
      double* A = M + (ib << 3) * ld + k0;
      double c0 = 0.0, c1 = 0.0;
      // C[g][n] = sum_k A[g][k] Li[n][k]:  a = A[g][kk + t],  b = B[k = kk + t][n = g] = Li[g][kk + t]
      const double a0 = A[g * ld + t], a1 = A[g * ld + 4 + t];
      const double b0 = Li[g * 8 + t], b1 = Li[g * 8 + 4 + t];
      rdmma(c0, c1, a0, b0);
      rdmma(c0, c1, a1, b1);
      __syncwarp();
      *reinterpret_cast<double2*>(A + g * ld + 2 * t) = make_double2(c0, c1);
    }
    __syncthreads();
    // trailing update: A22[ti][tj] -= X[ti] X[tj]^T for the lower tiles tj <= ti
    auto update_tile = [&](int ti, int tj) {
      const double* Xi = M + (ti << 3) * ld + k0;
      const double* Xj = M + (tj << 3) * ld + k0;
      double* Cp = M + ((ti << 3) + g) * ld + (tj << 3) + 2 * t;
      double2 c = *reinterpret_cast<double2*>(Cp);
      rdmma(c.x, c.y, -Xi[g * ld + t], Xj[g * ld + t]);
      rdmma(c.x, c.y, -Xi[g * ld + 4 + t], Xj[g * ld + 4 + t]);
      *reinterpret_cast<double2*>(Cp) = c;
    };
    if (kb + 1 < NB) {
      if (warp == 0) {  // the next diagonal tile, then its factorisation (look-ahead)
        update_tile(kb + 1, kb + 1);
        __syncwarp();
        chol8_inv(M + (k0 + 8) * ld + k0 + 8, ld, Linv + (kb + 1) * 64, lane, flag);
      } else {  // the other lower tiles (linear index 1 .. mt (mt + 1) / 2 - 1), dealt round-robin to warps 1 .. RW - 1
        const int mt = NB - kb - 1, ntile = mt * (mt + 1) / 2;
        for (int idx = warp; idx < ntile; idx += RW - 1) {
          int i = (int)((sqrtf(8.0f * idx + 1.0f) - 1.0f) * 0.5f);  // idx = i (i + 1) / 2 + j, j <= i
          if (i * (i + 1) / 2 > idx) --i;
          if ((i + 1) * (i + 2) / 2 <= idx) ++i;
          update_tile(kb + 1 + i, kb + 1 + idx - i * (i + 1) / 2);
        }
      }
    }
    __syncthreads();
  }
}

// Solve L L^T x = r in place (r -> x) with the factor in M and the inverse diagonal blocks in Linv; whole CTA.
__device__ void chol_solve_blocked(const double* M, int np, int ld, const double* Linv, double* r) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NB = np >> 3;
  const int row = lane >> 2, part = lane & 3;  // 8 x 8 mat-vec by one warp: lane (row, part) takes two columns
  for (int kb = 0; kb < NB; ++kb) {  // forward: L z = r
    const int k0 = kb << 3;
    if (warp == 0) {
      const double* Li = Linv + kb * 64;
      double v = Li[row * 8 + 2 * part] * r[k0 + 2 * part] + Li[row * 8 + 2 * part + 1] * r[k0 + 2 * part + 1];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      __syncwarp();
      if (part == 0) r[k0 + row] = v;
    }
    __syncthreads();
    for (int i = k0 + 8 + tid; i < np; i += RT) {
      const double* Lr = M + i * ld + k0;
      double acc = r[i];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc = fma(-Lr[c], r[k0 + c], acc);
      r[i] = acc;
    }
    __syncthreads();
  }
  for (int kb = NB - 1; kb >= 0; --kb) {  // backward: L^T x = z
    const int k0 = kb << 3;
    if (warp == 0) {
      const double* Li = Linv + kb * 64;  // x_blk = Li^T z_blk
      double v = Li[(2 * part) * 8 + row] * r[k0 + 2 * part] + Li[(2 * part + 1) * 8 + row] * r[k0 + 2 * part + 1];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      __syncwarp();
      if (part == 0) r[k0 + row] = v;
    }
    __syncthreads();
    for (int i = tid; i < k0; i += RT) {
      double acc = r[i];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc = fma(-M[(k0 + c) * ld + i], r[k0 + c], acc);
      r[i] = acc;
    }
    __syncthreads();
  }
}

// Shared-memory carve-up of one solver CTA
struct RidgeSmem {
  double *M, *P, *Linv, *q, *x, *prev, *rhs, *y, *lb, *lam, *sl, *red;
  int *F, *ictl;
};
__device__ __forceinline__ RidgeSmem ridge_carve(double* sm, int np, int ld, int p_smem) {
  RidgeSmem s;
  s.M = sm;
  s.P = s.M + np * ld;
  s.Linv = s.P + (p_smem ? np * ld : 0);
  s.q = s.Linv + (np >> 3) * 64;
  s.x = s.q + np;
  s.prev = s.x + np;
  s.rhs = s.prev + np;
  s.y = s.rhs + np;
  s.lb = s.y + np;
  s.lam = s.lb + np;      // 3 np
  s.sl = s.lam + 3 * np;  // 3 np: sqrt(lam)
  s.red = s.sl + 3 * np;  // 32
  s.F = (int*)(s.red + 32);
  s.ictl = s.F + np;  // [0] flag, [2] tries, [3] ninf, [5] pivots
  return s;
}
static size_t ridge_smem_bytes(int np, int ld, int p_smem) {
  return ((size_t)np * ld * (p_smem ? 2 : 1) + (size_t)(np >> 3) * 64 + 12 * (size_t)np + 32) * sizeof(double) +
         ((size_t)np + 8) * sizeof(int);
}

// Block principal pivoting on  min 1/2 x'Px + q'x, x >= lb  with P = Pg (n x n, row stride ldp, symmetric; shared or
// global memory), warm-started from the free set F.  On return rhs holds the solution.  Returns the KKT residual when
// want_res.
__device__ double bpp_solve(const RidgeSmem& s, const double* Pg, int ldp, int n, int np, int ld, double qinf,
                            bool want_res, int* n_pivot, unsigned long long* rdbg = nullptr) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double res = 0.0;
#ifdef BDRT_PHASE_CLOCKS
  long long rc_ = clock64();
#endif
  if (tid == 0) { s.ictl[2] = 3; s.ictl[3] = n + 1; }
  __syncthreads();
  double boost = 1.0;
  for (int pit = 0; pit < 500; ++pit) {
    // Near-singular programs can cycle on sign tests decided by rounding noise: every 100 pivot steps both tolerances are
    // relaxed a hundredfold and the exchange rule starts afresh (as oracle/ridge.py: qp_bound, measured there on the
    // reference's own saved cvxopt runs).
    if (pit > 0 && pit % 100 == 0) {
      boost *= 100.0;
      if (tid == 0) { s.ictl[2] = 3; s.ictl[3] = n + 1; }
      __syncthreads();
    }
    // working matrix: P on the free set, identity elsewhere (and on the padding)
    // (loads of four column chunks of two rows are issued before the first store: P may live in L2)
    for (int i0 = 2 * warp; i0 < np; i0 += 2 * RW) {
      for (int j0 = lane; j0 < np; j0 += 128) {
        double v[2][4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = i0 + r, j = j0 + 32 * c;
            v[r][c] = (i < n && j < n && s.F[i] && s.F[j]) ? Pg[(long long)i * ldp + j] : (i == j ? 1.0 : 0.0);
          }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int i = i0 + r, j = j0 + 32 * c;
            if (i < np && j < np) s.M[i * ld + j] = v[r][c];
          }
      }
    }
    for (int i = tid; i < np; i += RT) {
      double r = 0.0;
      if (i < n) {
        if (s.F[i]) {
          r = -s.q[i];
          for (int j = 0; j < n; ++j)
            if (!s.F[j] && s.lb[j] != 0.0) r = fma(-Pg[(long long)j * ldp + i], s.lb[j], r);  // P symmetric: coalesced
        } else {
          r = s.lb[i];
        }
      }
      s.rhs[i] = r;
    }
    if (tid == 0) s.ictl[0] = 0;
    __syncthreads();
    RCLK(1);  // assembly of the working matrix and right-hand side
    chol_blocked(s.M, np, ld, s.Linv, &s.ictl[0]);
    RCLK(2);  // factorisation
    chol_solve_blocked(s.M, np, ld, s.Linv, s.rhs);
    RCLK(3);  // triangular solves
    ++*n_pivot;
#ifdef BDRT_PHASE_CLOCKS
    if (tid == 0 && rdbg) atomicAdd(&rdbg[0], 1ull);
#endif
    // y = P x + q on the bound set (on every row when the KKT residual is wanted), violations
    double xm = 0.0;
    for (int i = tid; i < n; i += RT) xm = fmax(xm, fabs(s.rhs[i]));
    const double tol_x = boost * 1e-14 * fmax(block_max(xm, s.red), 1e-300), tol_y = boost * 1e-12 * qinf;
    int myv = 0, mymax = -1;
    double myres = 0.0;
    // (one warp per row, lanes over the columns: the rows of the bound set are spread over all warps)
    for (int i0 = 2 * warp; i0 < n; i0 += 2 * RW) {
      const int i1 = i0 + 1;
      const bool n0 = !s.F[i0] || want_res, n1 = i1 < n && (!s.F[i1] || want_res);
      double y0 = 0.0, y1 = 0.0;
      for (int j = lane; j < n; j += 32) {
        const double a0 = n0 ? Pg[(long long)i0 * ldp + j] : 0.0, a1 = n1 ? Pg[(long long)i1 * ldp + j] : 0.0;
        const double xj = s.rhs[j];
        y0 = fma(a0, xj, y0);
        y1 = fma(a1, xj, y1);
      }
      double two[2] = {y0, y1};
      const double tot = warp_sum_multi2(two, lane);
      if (lane == 0) s.y[i0] = n0 ? tot + s.q[i0] : 0.0;
      if (lane == 16 && i1 < n) s.y[i1] = n1 ? tot + s.q[i1] : 0.0;
    }
    __syncthreads();
    for (int i = tid; i < n; i += RT) {
      const double yi = s.y[i];
      const int v = s.F[i] ? (s.rhs[i] < s.lb[i] - tol_x) : (yi < -tol_y);
      // KKT residual: |gradient| on the free set, negative part of the multiplier on the bound set, bound violation
      myres = fmax(myres, s.F[i] ? fabs(yi) : fmax(0.0, -yi));
      myres = fmax(myres, fmax(0.0, s.lb[i] - s.rhs[i]));
      if (v) { ++myv; mymax = i; }
      s.F[i] = s.F[i] | (v << 1);  // bit 1: violation flag, consumed by the exchange step below
    }
    const double nv = block_sum((double)myv, s.red);
    const double vmax = block_max((double)mymax, s.red);
    if (want_res) res = block_max(myres, s.red);
    __syncthreads();
    if (nv == 0.0) {
      for (int i = tid; i < n; i += RT) s.F[i] &= 1;
      __syncthreads();
      break;
    }
    int mode;  // 0: exchange all, 1: exchange only the largest violating index
    {
      const int inv = (int)nv;
      if (inv < s.ictl[3]) mode = 0;
      else if (s.ictl[2] >= 1) mode = 0;
      else mode = 1;
      __syncthreads();
      if (tid == 0) {
        if (inv < s.ictl[3]) { s.ictl[3] = inv; s.ictl[2] = 3; }
        else if (s.ictl[2] >= 1) s.ictl[2] -= 1;
      }
    }
    for (int i = tid; i < n; i += RT) {
      const int v = (s.F[i] >> 1) & 1, f = s.F[i] & 1;
      s.F[i] = (v && (mode == 0 || i == (int)vmax)) ? (f ^ 1) : f;
    }
    __syncthreads();
    RCLK(4);  // multipliers, violations, exchange
  }
  RCLK(4);
  return res;
}

// G (n x n, global, lower AND upper filled) = W_re' W_re + W_im' W_im with W [Nf][n] row-major in global memory:
// 8 x 8 tiles of the lower triangle dealt round-robin to the warps, rows streamed four at a time through the
// tensor cores (rows past Nf and columns past n contribute zeros)
__device__ void gram_dmma(const double* Wr, const double* Wi, int Nf, int n, int np, double* G) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int NB = np >> 3;
  int cnt = 0;
  for (int ti = 0; ti < NB; ++ti) {
    for (int tj = 0; tj <= ti; ++tj, ++cnt) {
      if ((cnt & (RW - 1)) != warp) continue;
      const int ci = (ti << 3) + g, cj = (tj << 3) + g;
      const bool vi = ci < n, vj = cj < n;
      double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
      for (int part = 0; part < 2; ++part) {
        const double* W = part ? Wi : Wr;
#pragma unroll 4
        for (int r0 = 0; r0 < Nf; r0 += 4) {
          const int r = r0 + t;
          const bool vr = r < Nf;
          // C[i][j] = sum_r W[r][i] W[r][j]:  a = A[m = g][k = t] = W[r0 + t][ci],  b = B[k = t][n = g] = W[r0 + t][cj]
          const double a = (vr && vi) ? W[(long long)r * n + ci] : 0.0;
          const double b = (vr && vj) ? W[(long long)r * n + cj] : 0.0;
          if (part) rdmma(d0, d1, a, b);
          else rdmma(c0, c1, a, b);
        }
      }
      c0 += d0;
      c1 += d1;
      // thread holds C[row g][cols 2t, 2t+1] of the tile
      const int i = (ti << 3) + g, j = (tj << 3) + 2 * t;
      if (i < n) {
        if (j < n) { G[(long long)i * n + j] = c0; G[(long long)j * n + i] = c0; }
        if (j + 1 < n) { G[(long long)i * n + j + 1] = c1; G[(long long)(j + 1) * n + i] = c1; }
      }
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(RT, RT > 256 ? 1 : 2) ridge_kernel(RidgeArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.n, np = a.np, ld = a.ld, K = a.K, Nf = a.Nf;
  const RidgeSmem s = ridge_carve(sm, np, ld, a.p_smem);
  double* G0 = a.scratch + (long long)blockIdx.x * 2 * n * n;
  double* Pg = a.p_smem ? s.P : G0 + (long long)n * n;
  const int ldp = a.p_smem ? ld : n;
  const bool integral = a.o.penalty == 1;
  bool have_G0 = false;
#ifdef BDRT_PHASE_CLOCKS
  unsigned long long* rdbg = a.dbg;
  long long rc_ = clock64();
#endif

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const double* WAr = a.WA_re + (long long)b * a.wa_stride;
    const double* WAi = a.WA_im + (long long)b * a.wa_stride;
    const double* WZr = a.WZ_re + (long long)b * Nf;
    const double* WZi = a.WZ_im + (long long)b * Nf;
    // ---- G0 = WA_re' WA_re + WA_im' WA_im  (inversion.py:1045)
    if (a.wa_stride != 0 || !have_G0) {
      __syncthreads();
      gram_dmma(WAr, WAi, Nf, n, np, G0);
      have_G0 = true;
    }
    // ---- q = -WA_re' WZ_re - WA_im' WZ_im + L1_vec  (inversion.py:1046, :450-452)
    for (int i = tid; i < np; i += RT) {
      double sacc = 0.0;
      if (i < n)
        for (int r = 0; r < Nf; ++r) sacc += WAr[(long long)r * n + i] * WZr[r] + WAi[(long long)r * n + i] * WZi[r];
      const double l1 = (i < 2) ? 0.0 : sqrt(M_PI) / a.o.epsilon * a.o.L1_penalty;
      s.q[i] = i < n ? -sacc + l1 : 0.0;
      s.x[i] = i < n ? 1e-6 : 0.0;  // inversion.py:495
      s.lb[i] = (a.o.nonneg || i < 2 || i >= n) ? 0.0 : -10.0;  // inversion.py:1054-1064
      s.F[i] = 0;
      for (int o = 0; o < 3; ++o) s.lam[o * np + i] = a.o.lambda_0;
    }
    if (tid == 0) s.ictl[5] = 0;
    __syncthreads();
    double qm = 0.0;
    for (int i = tid; i < n; i += RT) qm = fmax(qm, fabs(s.q[i]));
    const double qinf = fmax(block_max(qm, s.red), 1e-300);
    RCLK(7);  // Gram matrix, right-hand side

    int it = 0, conv = 0, n_hyper = 0, n_pivot = 0;
    while (it < a.o.max_iter) {
      for (int i = tid; i < np; i += RT) s.prev[i] = s.x[i];
      __syncthreads();
      // ---- lambda update from the previous coefficients
      for (int o = 0; o < 3; ++o) {
        if (!(a.o.reg_ord[o] > 0.0)) continue;
        double* lamo = s.lam + o * np;
        if (!integral) {
          // lam_k = 1 / ((L_o c)_k^2/(beta-1) + 1/lambda_0), lam[0:2] = 1   (inversion.py:947-954)
          const double* Lo = a.Lmat + (long long)o * K * n;
          for (int k0 = 2 * warp; k0 < K; k0 += 2 * RW) {
            const int k1 = k0 + 1 < K ? k0 + 1 : k0;
            double a0 = 0.0, a1 = 0.0;
            for (int j = lane; j < n; j += 32) {
              const double pj = s.prev[j];
              a0 = fma(Lo[(long long)k0 * n + j], pj, a0);
              a1 = fma(Lo[(long long)k1 * n + j], pj, a1);
            }
            double two[2] = {a0, a1};
            const double acc = warp_sum_multi2(two, lane);
            const int k = lane < 16 ? k0 : k0 + 1;
            if ((lane & 15) == 0 && k < K) {
              if (a.o.hl_fbeta > 0.0) s.y[2 + k] = acc * acc;  // second pass below needs max_k (L c)_k^2
              else lamo[2 + k] = 1.0 / (acc * acc / (a.o.hl_beta - 1.0) + 1.0 / a.o.lambda_0);
            }
          }
          if (a.o.hl_fbeta > 0.0) {
            // lam_k = lambda_0 / ((L_o c)_k^2 / (max_k (L_o c)_k^2 * hl_fbeta) + 1)   (inversion.py:956-964)
            __syncthreads();
            double mx = 0.0;
            for (int k = tid; k < K; k += RT) mx = fmax(mx, s.y[2 + k]);
            mx = block_max(mx, s.red);
            for (int k = tid; k < K; k += RT) lamo[2 + k] = a.o.lambda_0 / (s.y[2 + k] / (mx * a.o.hl_fbeta) + 1.0);
          }
          if (tid < 2) lamo[tid] = 1.0;
        } else {
          // closed form of the integral penalty (inversion.py:973-983), coefficient factors 100 / 10 / 1 (:680-687)
          const double* Mo = a.Pen + (long long)o * n * n;
          const double factor = (o == 0) ? 100.0 : (o == 1 ? 10.0 : 1.0);
          for (int j = tid; j < n; j += RT) s.rhs[j] = factor * s.prev[j] * sqrt(lamo[j]);  // X Lambda^1/2 (previous)
          __syncthreads();
          for (int j = tid; j < n; j += RT) {
            const double cj = factor * s.prev[j];
            double C = 0.0;
            for (int i = 0; i < n; ++i)
              if (i != j) C = fma(s.rhs[i], Mo[(long long)i * n + j], C);
            C *= cj;
            const double aa = a.o.hl_beta / 2.0, bb = 0.5 * (2.0 * aa - 2.0) / a.o.lambda_0;
            const double d = cj * cj * Mo[(long long)j * n + j] + 2.0 * bb;
            const double sg = (C > 0.0) ? 1.0 : (C < 0.0 ? -1.0 : 0.0);
            double lv = (C * C - sg * C * sqrt(4.0 * d * (2.0 * aa - 2.0) + C * C) + 2.0 * d * (2.0 * aa - 2.0)) /
                        (2.0 * d * d);
            if (lv <= 0.0) lv = 1e-15;  // inversion.py:689
            s.y[j] = lv;
          }
          __syncthreads();
          for (int j = tid; j < n; j += RT) lamo[j] = s.y[j];
        }
        __syncthreads();
      }
      RCLK(5);  // lambda update
      // ---- P = G0 + sum_o frac_o Lam_o^1/2 Pen_o Lam_o^1/2   (inversion.py:695-700)
      for (int i = tid; i < 3 * np; i += RT) s.sl[i] = sqrt(s.lam[i]);
      __syncthreads();
      for (int i = warp; i < n; i += RW) {
        for (int j0 = lane; j0 < n; j0 += 128) {
          double acc[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = j0 + 32 * c;
            acc[c] = j < n ? G0[(long long)i * n + j] : 0.0;
          }
          for (int o = 0; o < 3; ++o) {
            if (!(a.o.reg_ord[o] > 0.0)) continue;
            const double si = a.o.reg_ord[o] * s.sl[o * np + i];
            double pv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j = j0 + 32 * c;
              pv[c] = j < n ? a.Pen[(long long)o * n * n + (long long)i * n + j] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int j = j0 + 32 * c;
              if (j < n) acc[c] = fma(si * pv[c], s.sl[o * np + j], acc[c]);
            }
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j = j0 + 32 * c;
            if (j < n) Pg[(long long)i * ldp + j] = acc[c];
          }
        }
      }
      __syncthreads();
      // ---- QP by block principal pivoting, warm-started from the previous free set
      RCLK(6);  // penalty assembly
#ifdef BDRT_PHASE_CLOCKS
      bpp_solve(s, Pg, ldp, n, np, ld, qinf, false, &n_pivot, rdbg);
      rc_ = clock64();
#else
      bpp_solve(s, Pg, ldp, n, np, ld, qinf, false, &n_pivot);
#endif
      for (int i = tid; i < n; i += RT) s.x[i] = s.rhs[i];
      __syncthreads();
      ++n_hyper;
      // ---- stop test  mean(|(c - prev)/prev|) < xtol  (inversion.py:730-736).  stop_rule 0: numpy semantics, 0/0 = NaN
      // compares false (what the reference's code does with an exact QP solver: a coefficient that sits on its bound in two
      // consecutive iterations blocks the test for good); stop_rule 1: such a coefficient counts as unchanged
      double dsum = 0.0;
      for (int i = tid; i < n; i += RT) {
        double d = fabs((s.x[i] - s.prev[i]) / s.prev[i]);
        if (a.o.stop_rule == 1 && s.x[i] == s.prev[i]) d = 0.0;
        if (i == 1 && !a.o.fit_inductance) d = 0.0;
        dsum += d;
      }
      dsum = block_sum(dsum, s.red);
      RCLK(8);  // stop test
      if (dsum / n < a.o.xtol) { conv = 1; break; }
      ++it;
    }
    for (int i = tid; i < n; i += RT) {
      a.coef[(long long)b * n + i] = s.x[i];
      for (int o = 0; o < 3; ++o) a.lam[((long long)b * 3 + o) * n + i] = s.lam[o * np + i];
    }
    if (tid == 0) {
      if (a.iters) a.iters[b] = n_hyper;
      if (a.converged) a.converged[b] = conv;
      if (a.pivots) a.pivots[b] = n_pivot;
    }
    __syncthreads();
  }
}

// stand-alone batched QP (test hook for the parity of the solver itself)
__global__ void __launch_bounds__(RT, RT > 256 ? 1 : 2) qp_kernel(const double* P, const double* qv, const double* lbv, int B, int n,
                                                   int np, int ld, int p_smem, double* xo, double* kkt, int* iters) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const RidgeSmem s = ridge_carve(sm, np, ld, p_smem);
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const double* Pg = P + (long long)b * n * n;
    int ldp = n;
    if (p_smem) {
      for (int i = tid >> 5; i < n; i += RW)
        for (int j = tid & 31; j < n; j += 32) s.P[i * ld + j] = Pg[(long long)i * n + j];
      Pg = s.P;
      ldp = ld;
      __syncthreads();
    }
    double qm = 0.0;
    for (int i = tid; i < np; i += RT) {
      s.q[i] = i < n ? qv[(long long)b * n + i] : 0.0;
      s.lb[i] = i < n ? lbv[i] : 0.0;
      s.F[i] = 0;
      if (i < n) qm = fmax(qm, fabs(s.q[i]));
    }
    const double qinf = fmax(block_max(qm, s.red), 1e-300);
    int n_pivot = 0;
    const double res = bpp_solve(s, Pg, ldp, n, np, ld, qinf, true, &n_pivot);
    for (int i = tid; i < n; i += RT) xo[(long long)b * n + i] = s.rhs[i];
    if (tid == 0) {
      if (kkt) kkt[b] = res;
      if (iters) iters[b] = n_pivot;
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
  o->stop_rule = 1;   // a coefficient that stays on its bound counts as unchanged in the stop test
}

static inline int ridge_ld(int np) {  // row stride % 16 in {4, 12}: conflict-free DMMA fragments
  int ld = np;
  while ((ld & 15) != 4 && (ld & 15) != 12) ++ld;
  return ld;
}

extern "C" int bdrt_ridge_fit(bdrt_ctx* ctx, const bdrt_ridge_opts* opts, const double* WA_re, const double* WA_im,
                              int per_spectrum_W, const double* WZ_re, const double* WZ_im, const double* Pen,
                              const double* Lmat, int B, int Nf, int K, double* coef, double* lam, int* iters,
                              int* converged, int* n_factor) {
  if (!ctx) return BDRT_E_NULL;
  if (!opts || !WA_re || !WA_im || !WZ_re || !WZ_im || !Pen || !coef || !lam)
    BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_ridge_fit: null pointer");
  if (opts->penalty != 0 && opts->penalty != 1) BDRT_FAIL(ctx, BDRT_E_MODEL, "penalty must be 0 (discrete) or 1 (integral)");
  if (opts->stop_rule != 0 && opts->stop_rule != 1) BDRT_FAIL(ctx, BDRT_E_MODEL, "stop_rule must be 0 or 1");
  if (opts->penalty == 0 && !Lmat) BDRT_FAIL(ctx, BDRT_E_NULL, "bdrt_ridge_fit: Lmat is required for the discrete penalty");
  if (opts->penalty == 0 && !(opts->hl_beta > 1.0))
    BDRT_FAIL(ctx, BDRT_E_SIZE, "hl_beta must be greater than 1 for penalty 'discrete'");  // inversion.py:286-288
  if (opts->penalty == 1 && !(opts->hl_beta > 2.0))
    BDRT_FAIL(ctx, BDRT_E_SIZE, "hl_beta must be greater than 2 for penalty 'integral'");  // inversion.py:289-291
  if (B < 0 || Nf < 1 || K < 1 || opts->max_iter < 1) BDRT_FAIL(ctx, BDRT_E_SIZE, "bad sizes");
  if (B == 0) return BDRT_OK;
  const int n = K + 2, np = (n + 7) & ~7, ld = ridge_ld(np);
  // CTAs of 8 warps: two per SM with the system matrix in the (L2-resident) global scratch -- their serial sections
  // (8 x 8 diagonal factorisations, triangular solves) overlap; CTAs of 16 warps: one per SM with both matrices on chip
  const int p_smem = RT > 256 && ridge_smem_bytes(np, ld, 1) <= (size_t)ctx->smem_optin;
  const size_t smem = ridge_smem_bytes(np, ld, p_smem);
  if (smem > (size_t)ctx->smem_optin) BDRT_FAIL(ctx, BDRT_E_SMEM, "K too large for the shared-memory QP solver");
  int per_sm = (int)(((size_t)ctx->smem_per_sm - 2048) / (smem + 1024));
  if (per_sm > (RT > 256 ? 1 : 2)) per_sm = RT > 256 ? 1 : 2;
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > B) grid = B;
  int rc = bdrt_ws_reserve(ctx, (size_t)grid * 2 * n * n * sizeof(double));
  if (rc) return rc;
  RidgeArgs a;
  a.o = *opts;
  a.WA_re = WA_re; a.WA_im = WA_im; a.WZ_re = WZ_re; a.WZ_im = WZ_im; a.Pen = Pen; a.Lmat = Lmat;
  a.wa_stride = per_spectrum_W ? (long long)Nf * n : 0;
  a.B = B; a.Nf = Nf; a.K = K; a.n = n; a.np = np; a.ld = ld; a.p_smem = p_smem;
  a.coef = coef; a.lam = lam; a.iters = iters; a.converged = converged; a.pivots = n_factor;
  a.scratch = (double*)ctx->ws;
  a.dbg = nullptr;
#ifdef BDRT_PHASE_CLOCKS
  if (!ctx->dbg_clk) {
    BDRT_CUDA(ctx, cudaMalloc(&ctx->dbg_clk, 16 * sizeof(unsigned long long)));
    BDRT_CUDA(ctx, cudaMemset(ctx->dbg_clk, 0, 16 * sizeof(unsigned long long)));
  }
  a.dbg = ctx->dbg_clk;
#endif
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
  const int np = (n + 7) & ~7, ld = ridge_ld(np);
  const int p_smem = RT > 256 && ridge_smem_bytes(np, ld, 1) <= (size_t)ctx->smem_optin;
  const size_t smem = ridge_smem_bytes(np, ld, p_smem);
  if (smem > (size_t)ctx->smem_optin) BDRT_FAIL(ctx, BDRT_E_SMEM, "n too large for the shared-memory QP solver");
  int per_sm = (int)(((size_t)ctx->smem_per_sm - 2048) / (smem + 1024));
  if (per_sm > (RT > 256 ? 1 : 2)) per_sm = RT > 256 ? 1 : 2;
  if (per_sm < 1) per_sm = 1;
  int grid = ctx->sm_count * per_sm;
  if (grid > B) grid = B;
  BDRT_CUDA(ctx, cudaFuncSetAttribute(qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  qp_kernel<<<grid, RT, smem, ctx->stream>>>(P, q, lb, B, n, np, ld, p_smem, x, kkt, iters);
  ctx->launches++;
  BDRT_CUDA(ctx, cudaGetLastError());
  return BDRT_OK;
}
