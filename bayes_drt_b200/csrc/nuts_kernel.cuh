// Kernel template of nuts.cu, shared by the per-layout translation units nuts_t*.cu (explicit
// instantiations, compiled in parallel) and by the host code in nuts.cu.
#pragma once
// Batched on-device adaptive NUTS sampler.
//
// Replaces StanModel.sampling(dat, warmup, iter, chains, seed, init, control={'adapt_delta':0.9,'adapt_t0':10})
// (bayes_drt/inversion.py:1218-1221): Stan 2.19.1's multinomial NUTS with diagonal metric, dual-averaging step size
// and windowed variance adaptation [Stan-upstream; restated in oracle/nuts.py, SURVEY appendix C].
//
// Mapping: one (spectrum, chain) per warp / column slot, 8 per persistent CTA, work pulled from an atomic queue; every
// leapfrog's gradient is a cooperative engine_eval() (engine.cuh).  The tree is built iteratively (no recursion): the
// no-U-turn checks of all sub-trees ending at a leaf use O(depth) checkpoints of (p, running rho) kept in L2-resident
// scratch; proposals inside a sub-tree are drawn by reservoir sampling, which has the same distribution as Stan's
// pairwise multinomial merges.  Random numbers: Philox4x32-10 keyed by (seed, global spectrum index, chain), counter =
// (iteration, draw index, stream), so results do not depend on how the batch is sharded or scheduled.
#include <math.h>

#include "engine.cuh"

#define NV 15  // per-slot work vectors
enum { V_ZQ, V_ZP, V_ZG, V_RSUB, V_MINV, V_RHO, V_OQ, V_OP, V_OG, V_SQ, V_SG, V_PQ, V_PG, V_WMEAN, V_WM2 };
#define MAXDEPTH 12
#define NCHN 7  // columns per lane and sweep chunk
// Sweep over a work vector: every vector is zero-padded to a multiple of 32 * NCHN coordinates (Dpad), so the hot sweeps
// have no bounds checks and unroll into NCHN independent element operations per lane
#define NUTS_SWEEP(i)                                             \
  for (int i0_ = lane; i0_ < Dpad; i0_ += 32 * NCHN)             \
    _Pragma("unroll") for (int c_ = 0, i = i0_; c_ < NCHN; ++c_, i += 32)

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ double u01(unsigned a, unsigned b) {  // (0, 1)
  return (((double)a * 4294967296.0 + (double)b) + 0.5) * (1.0 / 18446744073709551616.0);
}
struct Rng {
  uint2 key;
  unsigned w;
  __device__ double uniform(unsigned iter, unsigned idx, unsigned stream) const {
    const uint4 r = philox4x32_10(make_uint4(idx, iter, stream, w), key);
    return u01(r.x, r.y);
  }
  __device__ double normal(unsigned iter, unsigned idx, unsigned stream) const {
    const uint4 r = philox4x32_10(make_uint4(idx, iter, stream, w), key);
    const double u1 = u01(r.x, r.y), u2 = u01(r.z, r.w);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  }
};
__device__ __forceinline__ double logaddexp(double a, double b) {
  const double mx = fmax(a, b), mn = fmin(a, b);
  if (mx == -INFINITY) return -INFINITY;
  return mx + log1p(exp(mn - mx));
}

}  // namespace

template <int TOEP, int MK, int FAST>
__global__ void __launch_bounds__(NTHREADS, TOEP ? 2 : 1)
nuts_kernel(BdrtModel m, bdrt_nuts_opts o, const double* __restrict__ U0, double* draws, double* stepsize_out,
            long long* nleap_out, int* ndiv_out, int* nmax_out, double* accept_out, int* queue, double* gvec,
            double* ckpt, int nvec_smem, int Dpad) {
  extern __shared__ __align__(16) double sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = m.D;
  volatile int* n_active = (volatile int*)(sm + m.oUser);
  double* suser = sm + m.oUser + 2;
  if (threadIdx.x == 0) *n_active = NWARP;
  engine_load(m, sm, 0);

  double* v[NV];
  {
    double* gbase = gvec + ((long long)blockIdx.x * NSLOT + warp) * NV * Dpad;
    for (int i = 0; i < NV; ++i)
      v[i] = (i < nvec_smem) ? (suser + ((long long)warp * nvec_smem + i) * Dpad) : (gbase + (long long)i * Dpad);
  }
  double* CP = ckpt + ((long long)blockIdx.x * NSLOT + warp) * 2 * MAXDEPTH * Dpad;  // checkpoint momenta
  double* CR = CP + (long long)MAXDEPTH * Dpad;                                      // checkpoint running rho
  const long long n_work = (long long)m.B * o.chains;
  const int n_iter = o.warmup + o.samples;
  const double* Zs = m.Z;
  long long n_grad = 0;

  // leapfrog of the active end (V_ZQ, V_ZP, V_ZG); returns lp at the new point and the kinetic energy of the new
  // momentum (accumulated in the second half-step: one pass over the inverse metric less per leaf)
  double kin_new = 0.0;
  auto leapfrog = [&](double eps) -> double {
    double *q = v[V_ZQ], *p = v[V_ZP], *g = v[V_ZG];
    const double* mi = v[V_MINV];
    NUTS_SWEEP(i) {
      const double pi = fma(0.5 * eps, g[i], p[i]);
      p[i] = pi;
      q[i] = fma(eps * mi[i], pi, q[i]);
    }
    __syncwarp();
    int rsnap;
    const double lp = engine_eval<TOEP, MK, FAST>(m, sm, true, q, g, Zs, 1, n_active, TOEP == 2 ? &rsnap : nullptr);
    ++n_grad;
    double ks = 0.0;
    NUTS_SWEEP(i) {
      const double pi = fma(0.5 * eps, g[i], p[i]);
      p[i] = pi;
      ks = fma(mi[i] * pi, pi, ks);
    }
    kin_new = 0.5 * warp_sum(ks);
    __syncwarp();
    return lp;
  };
  auto kinetic = [&](const double* p) -> double {
    const double* mi = v[V_MINV];
    double s = 0.0;
    NUTS_SWEEP(i) s = fma(mi[i] * p[i], p[i], s);
    return 0.5 * warp_sum(s);
  };
  auto vcopy = [&](double* dst, const double* src) {
    NUTS_SWEEP(i) dst[i] = src[i];
  };

  // one adaptive chain: work item wi = spectrum * chains + chain
  auto run_chain = [&](long long wi) {
    const int b = (int)(wi / o.chains);
    Zs = m.Z + (long long)b * m.N2;
    const long long sid = o.spectrum_ids ? o.spectrum_ids[b] : o.spectrum_offset + b;
    const long long wglob = sid * (long long)o.chains + (wi % o.chains);
    Rng rng;
    rng.key = make_uint2((unsigned)(o.seed & 0xffffffffull), (unsigned)(o.seed >> 32) ^ (unsigned)(wglob >> 32));
    rng.w = (unsigned)wglob;
    n_grad = 0;

    // ---- initial point
    for (int i = lane; i < Dpad; i += 32) {  // zero padding of every work vector (the metric's is 1)
      for (int k = 0; k < NV; ++k) v[k][i] = 0.0;
      v[V_SQ][i] = i < D ? U0[wi * D + i] : 0.0;
      v[V_MINV][i] = 1.0;
    }
    __syncwarp();
    int rsnap0;
    double s_lp = engine_eval<TOEP, MK, FAST>(m, sm, true, v[V_SQ], v[V_SG], Zs, 1, n_active, TOEP == 2 ? &rsnap0 : nullptr);
    ++n_grad;
    bool bad = !isfinite(s_lp);

    double eps = 1.0;
    // dual averaging state
    double da_mu = 0, da_sbar = 0, da_xbar = 0;
    int da_count = 0;
    // windowed adaptation state (Stan windowed_adaptation)
    int w_init = 75, w_term = 50, w_base = 25;
    bool w_enabled = o.warmup >= 20;
    if (w_enabled && w_init + w_base + w_term > o.warmup) {
      w_init = (int)(0.15 * o.warmup);
      w_term = (int)(0.1 * o.warmup);
      w_base = o.warmup - (w_init + w_term);
    }
    int w_size = w_base, w_counter = 0, w_next = w_init + w_base - 1, w_n = 0;
    unsigned hdraw = 0;  // draw counter of the step-size heuristic

    // step-size heuristic (Stan base_hmc::init_stepsize): doubles / halves eps until the one-step acceptance crosses 0.8
    auto init_stepsize = [&]() {
      int direction = 0;
      for (int guard = 0; guard < 200; ++guard) {
        vcopy(v[V_ZQ], v[V_SQ]);
        vcopy(v[V_ZG], v[V_SG]);
        for (int i = lane; i < D; i += 32) v[V_ZP][i] = rng.normal(0xffffffffu, hdraw * 4096u + i, 2) * rsqrt(v[V_MINV][i]);
        // (the padding of V_ZP stays 0)
        ++hdraw;
        __syncwarp();
        const double H0 = -s_lp + kinetic(v[V_ZP]);
        const double lp1 = leapfrog(eps);
        double h = -lp1 + kin_new;
        if (isnan(h)) h = INFINITY;
        const double dH = H0 - h;
        const double thr = log(0.8);
        if (direction == 0) {
          direction = (dH > thr) ? 1 : -1;
          continue;  // Stan re-draws the momentum and re-tests at the same eps before changing it
        }
        if (direction == 1 && !(dH > thr)) break;
        if (direction == -1 && !(dH < thr)) break;
        eps = (direction == 1) ? 2.0 * eps : 0.5 * eps;
        if (eps > 1e7 || eps == 0.0) { bad = true; break; }
      }
    };

    double acc_sum = 0.0;
    int n_div = 0, n_maxd = 0;
    for (int i = lane; i < D; i += 32) { v[V_WMEAN][i] = 0.0; v[V_WM2][i] = 0.0; }
    bool need_stepsize = !bad;  // the heuristic runs before the first iteration and after every metric update

    for (int it = 0; it < n_iter && !bad; ++it) {
      if (need_stepsize) {  // single call site (the engine is inlined into it)
        init_stepsize();
        da_mu = log(10.0 * eps);
        need_stepsize = false;
        if (bad) break;
      }
      // ---------------------------------------------------------------- one NUTS transition from (SQ, SG, s_lp)
      unsigned udraw = 0;
      NUTS_SWEEP(i) {
        const double p = i < D ? rng.normal(it, i, 0) * rsqrt(v[V_MINV][i]) : 0.0;
        v[V_ZP][i] = p;
        v[V_OP][i] = p;
        v[V_RHO][i] = p;
        const double q = v[V_SQ][i], g = v[V_SG][i];
        v[V_ZQ][i] = q;
        v[V_OQ][i] = q;
        v[V_ZG][i] = g;
        v[V_OG][i] = g;
      }
      __syncwarp();
      double z_lp = s_lp, o_lp = s_lp;
      const double H0 = -s_lp + kinetic(v[V_ZP]);
      double lsw = 0.0, sum_metro = 0.0;
      int n_leap = 0, depth = 0, active_dir = 1;
      bool divergent = false;
      while (depth < o.max_treedepth) {
        const int dir = (rng.uniform(it, udraw++, 1) > 0.5) ? 1 : -1;
        if (dir != active_dir) {  // bring the other end of the trajectory into the working vectors
          NUTS_SWEEP(i) {
            double t_;
            t_ = v[V_ZQ][i]; v[V_ZQ][i] = v[V_OQ][i]; v[V_OQ][i] = t_;
            t_ = v[V_ZP][i]; v[V_ZP][i] = v[V_OP][i]; v[V_OP][i] = t_;
            t_ = v[V_ZG][i]; v[V_ZG][i] = v[V_OG][i]; v[V_OG][i] = t_;
          }
          const double t_ = z_lp; z_lp = o_lp; o_lp = t_;
          active_dir = dir;
          __syncwarp();
        }
        NUTS_SWEEP(i) v[V_RSUB][i] = 0.0;
        __syncwarp();
        double lsw_sub = -INFINITY, p_lp = 0.0;
        bool valid = true;
        const int n_leaves = 1 << depth;
        for (int leaf = 0; leaf < n_leaves; ++leaf) {
          z_lp = leapfrog(dir * eps);
          ++n_leap;
          double h = -z_lp + kin_new;
          if (isnan(h)) h = INFINITY;
          if (h - H0 > 1000.0) divergent = true;
          const double w = H0 - h;
          const double lsw_new = logaddexp(lsw_sub, w);
          sum_metro += (w > 0.0) ? 1.0 : exp(w);
          // uniform-over-weights (multinomial) proposal inside the sub-tree by reservoir sampling
          const bool take = (leaf == 0) || (rng.uniform(it, udraw++, 1) < exp(w - lsw_new));
          lsw_sub = lsw_new;
          NUTS_SWEEP(i) v[V_RSUB][i] += v[V_ZP][i];
          if (take) {  // warp-uniform, and rare deep in a sub-tree (probability ~ 1 / leaf)
            NUTS_SWEEP(i) {
              v[V_PQ][i] = v[V_ZQ][i];
              v[V_PG][i] = v[V_ZG][i];
            }
            p_lp = z_lp;
          }
          __syncwarp();
          if (divergent) { valid = false; break; }
          if ((leaf & 1) == 0) {
            const int idx = __popc(leaf >> 1);
            double *cp = CP + (long long)idx * Dpad, *cr = CR + (long long)idx * Dpad;
            NUTS_SWEEP(i) { cp[i] = v[V_ZP][i]; cr[i] = v[V_RSUB][i]; }
          } else {
            const int nsub = __ffs(~leaf) - 1;  // trailing ones: sub-trees of size 2, 4, .., 2^nsub end at this leaf
            const int idx_max = __popc(leaf >> 1);
            for (int idx = idx_max; idx > idx_max - nsub; --idx) {
              const double *cp = CP + (long long)idx * Dpad, *cr = CR + (long long)idx * Dpad;
              double d_s = 0.0, d_e = 0.0;
              NUTS_SWEEP(i) {
                const double rho_s = v[V_RSUB][i] - cr[i] + cp[i];
                const double mi = v[V_MINV][i];
                d_s = fma(mi * cp[i], rho_s, d_s);
                d_e = fma(mi * v[V_ZP][i], rho_s, d_e);
              }
              {  // both dot products in one butterfly: value 0 in lanes 0..15, value 1 in lanes 16..31
                double two[2] = {d_s, d_e};
                const double tot = warp_sum_multi<2>(two, lane);
                d_s = __shfl_sync(0xffffffffu, tot, 0);
                d_e = __shfl_sync(0xffffffffu, tot, 16);
              }
              if (!(d_s > 0.0 && d_e > 0.0)) { valid = false; break; }
            }
            if (!valid) break;
          }
        }
        if (!valid) break;
        ++depth;
        // biased progressive sampling between the old tree and the new sub-tree
        bool take_sub = lsw_sub > lsw;
        if (!take_sub) take_sub = rng.uniform(it, udraw++, 1) < exp(lsw_sub - lsw);
        if (take_sub) {
          vcopy(v[V_SQ], v[V_PQ]);
          vcopy(v[V_SG], v[V_PG]);
          s_lp = p_lp;
        }
        lsw = logaddexp(lsw, lsw_sub);
        double d_z = 0.0, d_o = 0.0;
        NUTS_SWEEP(i) {
          const double r = v[V_RHO][i] + v[V_RSUB][i];
          v[V_RHO][i] = r;
          const double mi = v[V_MINV][i];
          d_z = fma(mi * v[V_ZP][i], r, d_z);
          d_o = fma(mi * v[V_OP][i], r, d_o);
        }
        {
          double two[2] = {d_z, d_o};
          const double tot = warp_sum_multi<2>(two, lane);
          d_z = __shfl_sync(0xffffffffu, tot, 0);
          d_o = __shfl_sync(0xffffffffu, tot, 16);
        }
        __syncwarp();
        if (!(d_z > 0.0 && d_o > 0.0)) break;
      }
      const double accept = sum_metro / (double)(n_leap > 0 ? n_leap : 1);

      if (it < o.warmup) {
        // ---------------------------------------------------------------- adaptation (Stan adapt_diag_e_nuts)
        ++da_count;
        const double a = fmin(1.0, accept);
        const double eta = 1.0 / (da_count + o.adapt_t0);
        da_sbar = (1.0 - eta) * da_sbar + eta * (o.adapt_delta - a);
        const double x = da_mu - da_sbar * sqrt((double)da_count) / o.adapt_gamma;
        const double x_eta = pow((double)da_count, -o.adapt_kappa);
        da_xbar = (1.0 - x_eta) * da_xbar + x_eta * x;
        eps = exp(x);
        if (w_enabled) {
          const bool in_window = (w_counter >= w_init) && (w_counter < o.warmup - w_term) && (w_counter != o.warmup);
          if (in_window) {  // Welford
            ++w_n;
            for (int i = lane; i < D; i += 32) {
              const double q = v[V_SQ][i];
              const double d = q - v[V_WMEAN][i];
              const double mean = v[V_WMEAN][i] + d / w_n;
              v[V_WMEAN][i] = mean;
              v[V_WM2][i] += (q - mean) * d;
            }
          }
          const bool end_window = (w_counter == w_next) && (w_counter != o.warmup);
          if (end_window) {
            const int last = o.warmup - w_term - 1;
            if (w_next != last) {  // compute_next_window
              w_size *= 2;
              w_next = w_counter + w_size;
              if (w_next != last && w_next + 2 * w_size >= o.warmup - w_term) w_next = last;
            }
            const double n = (double)w_n;
            for (int i = lane; i < D; i += 32) {
              const double var = v[V_WM2][i] / (n - 1.0);
              v[V_MINV][i] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
              v[V_WMEAN][i] = 0.0;
              v[V_WM2][i] = 0.0;
            }
            w_n = 0;
            __syncwarp();
            need_stepsize = true;  // re-initialised from the current step size at the top of the next iteration
            da_count = 0;
            da_sbar = 0.0;
            da_xbar = 0.0;
          }
          ++w_counter;
        }
        if (it == o.warmup - 1) eps = exp(da_xbar);
      } else {
        if (draws) {
          double* dst = draws + (wi * o.samples + (it - o.warmup)) * D;
          for (int i = lane; i < D; i += 32) dst[i] = v[V_SQ][i];
        }
        acc_sum += accept;
        n_div += divergent ? 1 : 0;
        n_maxd += (depth >= o.max_treedepth) ? 1 : 0;
      }
    }
    if (bad && draws) {
      const double qnan = nan("");
      for (long long i = lane; i < (long long)o.samples * D; i += 32) draws[wi * o.samples * D + i] = qnan;
    }
    if (lane == 0) {
      if (stepsize_out) stepsize_out[wi] = bad ? nan("") : eps;
      if (nleap_out) nleap_out[wi] = n_grad;
      if (ndiv_out) ndiv_out[wi] = n_div;
      if (nmax_out) nmax_out[wi] = n_maxd;
      if (accept_out) accept_out[wi] = o.samples > 0 ? acc_sum / o.samples : 0.0;
    }
  };
  // Work distribution.  Shared grid: every warp pulls (spectrum, chain) items from the queue until it is empty.
  // Per-spectrum grids: the 8 slots of a CTA share the resident operands, so a CTA takes one spectrum at a time and
  // its slots run that spectrum's chains (chain = warp, warp + 8, ...).  A warp that is out of work keeps serving
  // engine_eval() until every slot of the CTA is done (single call site: the engine is inlined there).
  if constexpr (TOEP == 2) {
    // warp mode: the slots are independent -- every warp pulls (spectrum, chain) items until the queue is empty (with
    // per-spectrum grids it first loads that spectrum's tables into its own slot)
    long long loaded = -1;
    while (true) {
      long long wi = 0;
      if (lane == 0) wi = atomicAdd(queue, 1);
      wi = __shfl_sync(0xffffffffu, wi, 0);
      if (wi >= n_work) break;
      if (m.pslot && wi / o.chains != loaded) {
        loaded = wi / o.chains;
        engine_load_slot(m, sm, loaded);
      }
      run_chain(wi);
    }
    if (m.wsync) {  // keep answering the barrier until every slot of the CTA is done
      int snap;
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
        *n_active = o.chains < NWARP ? o.chains : NWARP;
      }
      cta_sync();
      b_cta = s_spec;
      if (b_cta >= m.B) break;
      engine_load(m, sm, b_cta);
    }
    if (!per_spec || warp < o.chains) {
      int c = warp;
      while (true) {
        long long wi;
        if (per_spec) {
          wi = c < o.chains ? (long long)b_cta * o.chains + c : n_work;
          c += NWARP;
        } else {
          wi = 0;
          if (lane == 0) wi = atomicAdd(queue, 1);
          wi = __shfl_sync(0xffffffffu, wi, 0);
        }
        if (wi >= n_work) break;
        run_chain(wi);
      }
      if (lane == 0) atomicSub((int*)n_active, 1);
    }
    int snap;
    do {
      engine_eval<TOEP, MK, FAST>(m, sm, false, nullptr, nullptr, nullptr, 0, n_active, &snap);
    } while (snap != 0);
    if (!per_spec) break;
  }
  }
}

