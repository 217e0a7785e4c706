// Kernel template of lbfgs.cu, shared by the per-layout translation units lbfgs_t*.cu (explicit
// instantiations, compiled in parallel) and by the host code in lbfgs.cu.
#pragma once
// Batched on-device L-BFGS MAP driver.
//
// Replaces StanModel.optimizing(dat, iter=max_iter, seed, init) (bayes_drt/inversion.py:1216): Stan 2.19.1's
// L-BFGS [Stan-upstream] on the unconstrained vector, objective f = -log_prob(jacobian=false): two-loop recursion
// with history 5, bracketing + cubic-zoom strong-Wolfe line search, Stan's convergence tests in Stan's order.
//
// Mapping: persistent CTAs (one per SM) pull spectra from an atomic queue; each of the 8 warps of a CTA runs the
// whole optimiser for one spectrum with warp-uniform control flow (vectors distributed over lanes, dot products by
// shuffle), and all 8 meet in engine_eval() for every objective/gradient evaluation (see engine.cuh).  A warp that
// finishes refills its slot from the queue, so slots never idle while work remains.
#include <math.h>

#include "engine.cuh"

#define MAXHIST 16

namespace {

struct Ctl {  // shared control block
  int n_active;
};

__device__ __forceinline__ double vdot(const double* a, const double* b, int D, int lane) {
  double s = 0.0;
  for (int i = lane; i < D; i += 32) s = fma(a[i], b[i], s);
  return warp_sum(s);
}

struct LS {  // per-warp optimiser state living in registers (warp-uniform)
  double f1, dfp1, alpha;
  int neval;
};

// cubic_interp of Stan's bfgs_linesearch.hpp [Stan-upstream]; see oracle/lbfgs.py:cubic_interp
__device__ double cubic_interp(double df0, double x1, double f1, double df1, double lo, double hi) {
  const double c3 = (-12.0 * f1 + 6.0 * x1 * (df0 + df1)) / (x1 * x1 * x1);
  const double c2 = -(4.0 * df0 + 2.0 * df1) / x1 + 6.0 * f1 / (x1 * x1);
  const double c1 = df0;
  const double t_s = sqrt(c2 * c2 - 2.0 * c1 * c3);
  const double s1 = -(c2 + t_s) / c3, s2 = -(c2 - t_s) / c3;
  auto val = [&](double s) { return s * (s * (s * c3 / 3.0 + c2) / 2.0 + c1); };
  double minF = val(lo), minX = lo;
  double tmp = val(hi);
  if (tmp < minF) { minF = tmp; minX = hi; }
  if (lo < s1 && s1 < hi) { tmp = val(s1); if (tmp < minF) { minF = tmp; minX = s1; } }
  if (lo < s2 && s2 < hi) { tmp = val(s2); if (tmp < minF) { minF = tmp; minX = s2; } }
  return minX;
}

// L-BFGS two-loop recursion  pk = -H gk  over the history entries head-nh .. head-1 (mod H), newest first.
// NC > 0: the search direction lives in registers (lane owns elements lane + 32 c, c < NC) and both history vectors
// of an entry are fetched together, so each entry costs one L2 round trip instead of two and no shared-memory
// traffic; the summation order is that of vdot(), so the result is bitwise that of the generic path (NC == 0).
template <int NC>
__device__ __forceinline__ double two_loop(const double* S, const double* Y, const double* rho, double* al,
                                         const double* gk, double* pk, int D, int Dpad, int nh, int head, int H,
                                         double gamma, int lane) {
  if (NC == 0) {
    for (int i = lane; i < D; i += 32) pk[i] = -gk[i];
    __syncwarp();
    for (int j = 0; j < nh; ++j) {  // newest -> oldest
      const int h = head - 1 - j + (head - 1 - j < 0 ? H : 0);  // j < nh <= H
      const double* Sh = S + (long long)h * Dpad;
      const double* Yh = Y + (long long)h * Dpad;
      const double a = rho[h] * vdot(Sh, pk, D, lane);
      al[h] = a;
      for (int i = lane; i < D; i += 32) pk[i] = fma(-a, Yh[i], pk[i]);
      __syncwarp();
    }
    for (int i = lane; i < D; i += 32) pk[i] *= gamma;
    __syncwarp();
    for (int j = nh - 1; j >= 0; --j) {  // oldest -> newest
      const int h = head - 1 - j + (head - 1 - j < 0 ? H : 0);  // j < nh <= H
      const double* Sh = S + (long long)h * Dpad;
      const double* Yh = Y + (long long)h * Dpad;
      const double beta = rho[h] * vdot(Yh, pk, D, lane);
      const double c = al[h] - beta;
      for (int i = lane; i < D; i += 32) pk[i] = fma(c, Sh[i], pk[i]);
      __syncwarp();
    }
    return vdot(gk, pk, D, lane);
  }
  constexpr int NCC = NC > 0 ? NC : 1;
  double pr[NCC];
#pragma unroll
  for (int c = 0; c < NCC; ++c) {
    const int i = lane + 32 * c;
    pr[c] = -gk[i];
  }
  // history entry -> registers (both vectors of the entry in one go: one L2 round trip)
  auto fetch = [&](int h, double (&sv)[NCC], double (&yv)[NCC]) {
    const double* Sh = S + (long long)h * Dpad;
    const double* Yh = Y + (long long)h * Dpad;
#pragma unroll
    for (int c = 0; c < NCC; ++c) {
      const int i = lane + 32 * c;
      sv[c] = Sh[i];
      yv[c] = Yh[i];
    }
  };
  auto hidx = [&](int j) { return head - 1 - j + (head - 1 - j < 0 ? H : 0); };  // j < nh <= H
  // (prefetching the next entry, or just the next entry's first vector, into another register buffer was measured
  // slower: it spills under the 128-register cap of the two-CTAs-per-SM kernels -- 52.9 M resp. 61 M vs 72 M
  // gradients/s)
  double sa[NCC], ya[NCC];
  auto first = [&](int h, const double (&sv)[NCC], const double (&yv)[NCC]) {  // newest -> oldest
    double d = 0.0;
#pragma unroll
    for (int c = 0; c < NCC; ++c) d = fma(sv[c], pr[c], d);
    const double a = rho[h] * warp_sum(d);
    al[h] = a;
#pragma unroll
    for (int c = 0; c < NCC; ++c) pr[c] = fma(-a, yv[c], pr[c]);
  };
  auto second = [&](int h, const double (&sv)[NCC], const double (&yv)[NCC]) {  // oldest -> newest
    double d = 0.0;
#pragma unroll
    for (int c = 0; c < NCC; ++c) d = fma(yv[c], pr[c], d);
    const double cc = al[h] - rho[h] * warp_sum(d);
#pragma unroll
    for (int c = 0; c < NCC; ++c) pr[c] = fma(cc, sv[c], pr[c]);
  };
  {
    for (int j = 0; j < nh; ++j) {
      fetch(hidx(j), sa, ya);
      first(hidx(j), sa, ya);
    }
#pragma unroll
    for (int c = 0; c < NCC; ++c) pr[c] *= gamma;
    for (int j = nh - 1; j >= 0; --j) {
      fetch(hidx(j), sa, ya);
      second(hidx(j), sa, ya);
    }
  }
  double gp = 0.0;  // gk . pk in vdot's summation order
#pragma unroll
  for (int c = 0; c < NCC; ++c) {
    const int i = lane + 32 * c;
    pk[i] = pr[c];
    gp = fma(gk[i], pr[c], gp);
  }
  __syncwarp();
  return warp_sum(gp);
}

}  // namespace

template <int TOEP, int MK, int FAST>
__global__ void __launch_bounds__(NTHREADS, TOEP ? 2 : 1)
lbfgs_kernel(BdrtModel m, bdrt_lbfgs_opts o, double* __restrict__ U, double* lp_out, int* iters_out, int* neval_out,
             int* status_out, int* queue, double* hist, double* gvec, int nvec_smem, int Dpad) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dtrue = m.D;
  const int D = Dpad;  // every work vector is zero-padded to a multiple of 224 coordinates: no sweep has a bounds check
  const int H = o.history;
  volatile int* n_active = (volatile int*)(sm + m.oUser);
  double* suser = sm + m.oUser + 2;
  if (threadIdx.x == 0) *n_active = NWARP;
  engine_load(m, sm, 0);

  // the slot's five work vectors: the first nvec_smem live in shared memory, the rest in global scratch
  double* vec[5];
  {
    double* gbase = gvec + ((long long)blockIdx.x * NSLOT + warp) * 5 * Dpad;
    for (int i = 0; i < 5; ++i)
      vec[i] = (i < nvec_smem) ? (suser + ((long long)warp * nvec_smem + i) * Dpad) : (gbase + (long long)i * Dpad);
  }
  // xn / gn are what the engine reads and writes (scattered accesses): they always live in the first (shared-memory)
  // vectors; gk, which is touched once per iteration, is the one that overflows to global scratch
  double* const xn = vec[0];
  double* const gn = vec[1];
  double* const pk = vec[2];
  double* const xk = vec[3];
  double* const gk = vec[4];
  double* S = hist + ((long long)blockIdx.x * NSLOT + warp) * 2 * H * Dpad;  // compact: the L2 footprint is what is used
  double* Y = S + (long long)H * Dpad;
  int snap;

  const double EPS = 2.220446049250313e-16;
  int neval = 0;
  const double* Zs = m.Z;

  // f(xn) -> f, gn = grad f ; returns false when not finite
  // also returns the slope gn . pk along the current direction (same summation order as vdot): one pass instead of two
  auto feval = [&](double& f, double& slope) -> bool {
    const double lp = engine_eval<TOEP, MK, FAST>(m, sm, true, xn, gn, Zs, 0, n_active, TOEP == 2 ? &snap : nullptr);
    ++neval;
    int fin = isfinite(lp);
    double d = 0.0;
    for (int i = lane; i < D; i += 32) {
      const double v = -gn[i];
      gn[i] = v;
      fin &= isfinite(v);
      d = fma(v, pk[i], d);
    }
    slope = warp_sum(d);
    __syncwarp();
    f = -lp;
    return warp_and(fin);
  };
  auto step_to = [&](double alpha) {
    for (int i = lane; i < D; i += 32) xn[i] = fma(alpha, pk[i], xk[i]);
    __syncwarp();
  };

  auto run_spectrum = [&](int b) {
    Zs = m.Z + (long long)b * m.N2;
    double* ub = U + (long long)b * Dtrue;
    for (int i = lane; i < D; i += 32) {  // zero padding of every vector; pk is read (unused) by the first evaluation
      xn[i] = i < Dtrue ? ub[i] : 0.0;
      gn[i] = 0.0;
      pk[i] = 0.0;
      xk[i] = 0.0;
      gk[i] = 0.0;
    }
    __syncwarp();
    neval = 0;
    double fk = nan("");
    int code = BDRT_TERM_RUNNING, it = 0;
    double slope_ev = 0.0;
    if (!feval(fk, slope_ev)) {
      code = BDRT_TERM_BADINIT;
    } else {
      for (int i = lane; i < D; i += 32) {
        const double g = gn[i];
        xk[i] = xn[i];
        gk[i] = g;
        pk[i] = -g;
      }
      __syncwarp();
      int nh = 0, head = 0;  // history: entries head-nh .. head-1 (mod H), newest = head-1
      double rho[MAXHIST], al[MAXHIST];
      double gamma = 1.0, alpha = o.init_alpha;
      double gp_prev = 0.0;  // g.p of the accepted point = the next line search's initial slope
      double fk_1 = 0.0, dfp_old = 0.0, dfp_new = 0.0;

      while (code == BDRT_TERM_RUNNING) {
        ++it;
        bool reset = (it == 1);
        bool ls_ok = false;
        double f1 = 0.0, newDFp = 0.0;
        while (true) {
          if (reset) {
            for (int i = lane; i < D; i += 32) pk[i] = -gk[i];
            __syncwarp();
          }
          if (it > 1 && !reset)
            alpha = fmin(1.0, 1.01 * cubic_interp(dfp_old, alpha, fk - fk_1, dfp_new, 1e-12, 1.0));
          else
            alpha = o.init_alpha;
          // ---------------- WolfeLineSearch (c1 = 1e-4, c2 = 0.9, minAlpha = 1e-12, <= 20 its, <= 10 restarts)
          // same vectors, same summation order as the convergence test of the previous iteration: reuse its value
          const double dfp = (it > 1 && !reset) ? gp_prev : vdot(gk, pk, D, lane);
          const double c1dfp = 1e-4 * dfp, c2dfp = 0.9 * dfp;
          double alpha0 = 1e-12, prevF = fk, prevDFp = dfp;
          int nits = 0, restarts = 0, ret = -1;  // ret: 0 ok, 1 fail
          // One loop for both stages of Stan's line search (bracketing: WolfeLineSearch, then WolfLSZoom with
          // min_range 1e-16), so that the objective is evaluated at a single call site (the engine is inlined there).
          bool zoom = false, retry = false;
          double alo = 0, aloF = 0, aloDFp = 0, ahi = 0, ahiF = 0, ahiDFp = 0;
          int itn = 0;
          while (ret < 0) {
            if (!zoom) {
              if (nits >= 20) { ret = 1; break; }
            } else if (!retry) {
              ++itn;
              if (fabs(alo - ahi) < 1e-16) { ret = 1; break; }
              if (itn % 5 == 0) {
                alpha = 0.5 * (alo + ahi);
              } else {
                const double d1 = aloDFp + ahiDFp - 3.0 * (aloF - ahiF) / (alo - ahi);
                double d2 = sqrt(d1 * d1 - aloDFp * ahiDFp);
                if (ahi < alo) d2 = -d2;
                alpha = ahi - (ahi - alo) * (ahiDFp + d2 - d1) / (ahiDFp - aloDFp + 2.0 * d2);
                const double lo = fmin(alo, ahi), hi = fmax(alo, ahi), rng = fabs(alo - ahi);
                if (!isfinite(alpha) || alpha < lo + 0.01 * rng || alpha > hi - 0.01 * rng) alpha = 0.5 * (alo + ahi);
              }
            }
            step_to(alpha);
            const bool okev = feval(f1, slope_ev);
            if (!zoom) {
              if (!okev) {
                if (restarts >= 10) { ret = 1; break; }
                alpha = 0.5 * (alpha0 + alpha);
                ++restarts;
                continue;
              }
              restarts = 0;
              newDFp = slope_ev;
              if (f1 > fk + alpha * c1dfp || (f1 >= prevF && nits > 0)) {
                alo = alpha0; aloF = prevF; aloDFp = prevDFp; ahi = alpha; ahiF = f1; ahiDFp = newDFp;
                zoom = true;
                continue;
              }
              if (fabs(newDFp) <= -c2dfp) { ret = 0; break; }
              if (newDFp >= 0) {
                alo = alpha; aloF = f1; aloDFp = newDFp; ahi = alpha0; ahiF = prevF; ahiDFp = prevDFp;
                zoom = true;
                continue;
              }
              alpha0 = alpha; prevF = f1; prevDFp = newDFp;
              alpha *= 10.0;
              ++nits;
            } else {
              if (!okev) {
                alpha = 0.5 * (alpha + fmin(alo, ahi));
                if (fabs(fmin(alo, ahi) - alpha) < 1e-16) { ret = 1; break; }
                retry = true;
                continue;
              }
              retry = false;
              newDFp = slope_ev;
              if (f1 > (fk + alpha * c1dfp) || f1 >= aloF) {
                ahi = alpha; ahiF = f1; ahiDFp = newDFp;
              } else {
                if (fabs(newDFp) <= -c2dfp) { ret = 0; break; }
                if (newDFp * (ahi - alo) >= 0) { ahi = alo; ahiF = aloF; ahiDFp = aloDFp; }
                alo = alpha; aloF = f1; aloDFp = newDFp;
              }
            }
          }
          if (ret != 0) {
            if (reset) break;  // failed even from steepest descent
            reset = true;
            continue;
          }
          ls_ok = true;
          dfp_old = dfp;
          break;
        }
        if (!ls_ok) { code = BDRT_TERM_LSFAIL; break; }

        // accepted: xn, gn, f1, newDFp.  s = xn - xk, y = gn - gk
        double* Sn = S + (long long)head * Dpad;
        double* Yn = Y + (long long)head * Dpad;
        double skyk = 0, yy = 0, gg = 0, ss = 0;
        for (int i = lane; i < D; i += 32) {  // and the accepted point becomes the current one
          const double xv = xn[i], gv = gn[i];
          const double s = xv - xk[i], y = gv - gk[i];
          Sn[i] = s;
          Yn[i] = y;
          xk[i] = xv;
          gk[i] = gv;
          skyk = fma(s, y, skyk);
          yy = fma(y, y, yy);
          ss = fma(s, s, ss);
          gg = fma(gv, gv, gg);
        }
        {  // one batched butterfly: value i ends up in lanes 8 i .. 8 i + 7
          double four[4] = {skyk, yy, ss, gg};
          const double tot = warp_sum_multi<4>(four, lane);
          skyk = __shfl_sync(0xffffffffu, tot, 0);
          yy = __shfl_sync(0xffffffffu, tot, 8);
          ss = __shfl_sync(0xffffffffu, tot, 16);
          gg = __shfl_sync(0xffffffffu, tot, 24);
        }
        const double gradNorm = sqrt(gg), stepNorm = sqrt(ss);
        dfp_new = newDFp;
        if (reset) {
          const double B0 = yy / skyk;
          nh = 0;
          // keep the entry we just wrote as the only one
          dfp_old /= B0;
          dfp_new /= B0;
          alpha *= B0;
          if (head != 0) {
            for (int i = lane; i < D; i += 32) { S[i] = Sn[i]; Y[i] = Yn[i]; }
            head = 0;
          }
        }
        gamma = skyk / yy;
        rho[head] = 1.0 / skyk;
        head = head + 1 == H ? 0 : head + 1;
        if (nh < H) ++nh;
        fk_1 = fk;
        fk = f1;
        __syncwarp();
        __threadfence_block();
        // ---------------- two-loop recursion -> pk
        double gp;
        if (D == 32 * 7)
          gp = two_loop<7>(S, Y, rho, al, gn, pk, D, Dpad, nh, head, H, gamma, lane);
        else if (D == 32 * 14)
          gp = two_loop<14>(S, Y, rho, al, gn, pk, D, Dpad, nh, head, H, gamma, lane);
        else
          gp = two_loop<0>(S, Y, rho, al, gn, pk, D, Dpad, nh, head, H, gamma, lane);
        // ---------------- convergence tests, Stan's order
        const double df = fabs(fk_1 - fk);
        gp_prev = gp;
        if (df < o.tol_obj)
          code = BDRT_TERM_ABSF;
        else if (df < o.tol_rel_obj * fmax(fabs(fk_1), fmax(fabs(fk), 1.0)) * EPS)
          code = BDRT_TERM_RELF;
        else if (gradNorm < o.tol_grad)
          code = BDRT_TERM_ABSGRAD;
        else if (-gp / fmax(fabs(fk), 1.0) < o.tol_rel_grad * EPS)
          code = BDRT_TERM_RELGRAD;
        else if (stepNorm < o.tol_param)
          code = BDRT_TERM_ABSX;
        else if (it >= o.max_iter)
          code = BDRT_TERM_MAXIT;
      }
      for (int i = lane; i < Dtrue; i += 32) ub[i] = xk[i];
    }
    if (lane == 0) {
      if (lp_out) lp_out[b] = -fk;
      if (iters_out) iters_out[b] = it;
      if (neval_out) neval_out[b] = neval;
      if (status_out) status_out[b] = code;
    }
  };
  // Work distribution.  Shared grid: every warp pulls spectra from the queue until it is empty.  Per-spectrum grids: the
  // slots of a CTA share the resident operands, so the CTA takes one spectrum at a time, slot 0 optimises it and the
  // other warps only serve the cooperative products.  Either way a warp that is out of work keeps serving engine_eval()
  // until every slot of the CTA is done (single call site: the engine is inlined there).
  if constexpr (TOEP == 2) {
    // warp mode: the slots are independent -- every warp pulls spectra until the queue is empty (with per-spectrum
    // grids it first loads that spectrum's tables into its own slot)
    while (true) {
      int b = 0;
      if (lane == 0) b = atomicAdd(queue, 1);
      b = __shfl_sync(0xffffffffu, b, 0);
      if (b >= m.B) break;
      if (m.pslot) engine_load_slot(m, sm, b);
      run_spectrum(b);
    }
    if (m.wsync) {  // synchronised warp mode: keep answering the barrier until every slot of the CTA is done
      do {
        engine_eval<TOEP, MK, FAST>(m, sm, false, nullptr, nullptr, nullptr, 0, n_active, &snap);
      } while (snap != 0);
    }
  } else {
  const bool per_spec = m.d[0].A_stride != 0;
  __shared__ int s_spec;
  while (true) {
    int b_cta = -1;
    if (per_spec) {
      cta_sync();
      if (threadIdx.x == 0) {
        s_spec = atomicAdd(queue, 1);
        *n_active = 1;
      }
      cta_sync();
      b_cta = s_spec;
      if (b_cta >= m.B) break;
      engine_load(m, sm, b_cta);
    }
    if (!per_spec || warp == 0) {
      while (true) {
        int b = b_cta;
        if (!per_spec) {
          if (lane == 0) b = atomicAdd(queue, 1);
          b = __shfl_sync(0xffffffffu, b, 0);
        }
        if (b < 0 || b >= m.B) break;
        run_spectrum(b);
        b_cta = -1;
      }
      if (lane == 0) atomicSub((int*)n_active, 1);
    }
    do {
      engine_eval<TOEP, MK, FAST>(m, sm, false, nullptr, nullptr, nullptr, 0, n_active, &snap);
    } while (snap != 0);
    if (!per_spec) break;
  }
  }
}

