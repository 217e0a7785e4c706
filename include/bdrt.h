/*
 * bdrt.h -- C ABI of the B200-native bayes-drt inversion hot path (libbdrt.so).
 *
 * The reference (jdhuang-csm/bayes-drt) is pure Python and has no FFI of its own: its seams are Python call
 * sites into third-party natives.  Each entry point below names the reference call site
 * (file:line, relative to the reference root) that it replaces; INTEGRATION.md shows the ctypes stub a maintainer
 * would add at that call site.
 *
 * Conventions
 *   - every function returns an int status: 0 ok, <0 argument error (BDRT_E_*), >0 a cudaError_t value;
 *     bdrt_last_error(ctx) gives a human-readable message for the last failure on that context;
 *   - no exceptions cross the boundary, no global state besides the opaque bdrt_ctx;
 *   - the CALLER owns every input/output buffer; all pointers are DEVICE pointers (cudaMalloc / torch
 *     tensor.data_ptr()), FP64, row-major, leading batch dimension, unless the name ends in _host;
 *   - work is enqueued on the context's stream and is asynchronous with respect to the host; the library only
 *     allocates its own scratch workspace (grown lazily, freed by bdrt_ctx_destroy);
 *   - one context may be used by one host thread at a time; distinct contexts are independent.
 */
#ifndef BDRT_H
#define BDRT_H

#ifdef __cplusplus
extern "C" {
#endif

#define BDRT_VERSION 100

/* ---- status codes ------------------------------------------------------------------------------------------ */
#define BDRT_OK 0
#define BDRT_E_NULL (-1)        /* null pointer argument */
#define BDRT_E_SIZE (-2)        /* a size / shape argument is out of range */
#define BDRT_E_MODEL (-3)       /* unknown / unsupported model or kernel id */
#define BDRT_E_SMEM (-4)        /* problem does not fit the per-CTA shared-memory budget */
#define BDRT_E_UNSUPPORTED (-5) /* option recognised but not implemented (never a silent fallback) */

/* ---- enumerations ------------------------------------------------------------------------------------------ */
/* kernels of construct_A (bayes_drt/matrices.py:27-117) */
enum { BDRT_KERNEL_DRT = 0, BDRT_KERNEL_DDT = 1 };
enum { BDRT_DIST_SERIES = 0, BDRT_DIST_PARALLEL = 1 };
enum { BDRT_SYM_PLANAR = 0, BDRT_SYM_SPHERICAL = 1 };
enum { BDRT_BC_TRANSMISSIVE = 0, BDRT_BC_BLOCKING = 1 };

/* Stan programs the reference selects in Inverter._get_stan_model (bayes_drt/inversion.py:1576-1610):
 * model id = base | flags. */
enum {
  BDRT_MODEL_SERIES = 0,          /* stan_model_files/Series_modelcode.txt */
  BDRT_MODEL_SERIES_PARALLEL = 1, /* stan_model_files/Series-Parallel_modelcode.txt (xp_raw is always lower=0) */
  BDRT_MODEL_SERIES_2PARALLEL = 3, /* stan_model_files/Series-2Parallel_modelcode.txt: one series + two parallel */
  BDRT_MODEL_PARALLEL = 2,        /* stan_model_files/Parallel_modelcode.txt: one parallel distribution, Z_hat = 1/(A x)
                                     + offsets, vector<lower=0> x (parameter layout of Series_pos) */
  BDRT_MODEL_POS = 16,            /* *_pos_modelcode.txt: vector<lower=0> x */
  BDRT_MODEL_OUTLIERS = 32        /* *_outliers_modelcode.txt */
};

/* per-spectrum termination codes of bdrt_map_lbfgs (Stan's L-BFGS return codes [Stan-upstream]) */
enum {
  BDRT_TERM_RUNNING = 0,
  BDRT_TERM_ABSX = 10,
  BDRT_TERM_ABSF = 20,
  BDRT_TERM_RELF = 21,
  BDRT_TERM_ABSGRAD = 30,
  BDRT_TERM_RELGRAD = 31,
  BDRT_TERM_MAXIT = 40,
  BDRT_TERM_LSFAIL = -1,
  BDRT_TERM_BADINIT = -2 /* log-density or gradient not finite at the initial point */
};

typedef struct bdrt_ctx bdrt_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
int bdrt_version(void);
/* stream: a cudaStream_t cast to void* (NULL = legacy default stream) */
int bdrt_ctx_create(int device, void* stream, bdrt_ctx** out);
int bdrt_ctx_destroy(bdrt_ctx* ctx);
const char* bdrt_last_error(const bdrt_ctx* ctx);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
long long bdrt_launch_count(const bdrt_ctx* ctx);

/* ---- kernel / penalty matrices ------------------------------------------------------------------------------
 * Replace bayes_drt/matrices.py construct_A (:120-265, default integrate_method='trapz': np.trapz over
 * np.linspace(-20,20,1000)), construct_L (:268-325) and construct_M (:366-411), called from
 * Inverter._prep_matrices (inversion.py:2250-2271, :2296-2307).
 *
 * freq : [n_grids, Nf] measurement frequencies in Hz (any order; row n of A belongs to freq[n])
 * tau  : [n_grids, K] if tau_per_grid else [K] basis time constants
 * A_re, A_im : [n_grids, Nf, K] outputs (either may be NULL)
 * k_ct is ignored unless ct != 0.  symmetry / bc are ignored for the DRT kernel. */
int bdrt_build_A(bdrt_ctx* ctx, const double* freq, int n_grids, int Nf, const double* tau, int K,
                 int tau_per_grid, double epsilon, int kernel, int dist_type, int symmetry, int bc, int ct,
                 double k_ct, double* A_re, double* A_im);

/* L[n,m] = d^order/dy^order exp(-(eps y)^2) at y = ln(1/(2 pi f_n tau_m)); construct_L(frequencies, tau, ...,
 * order) with integer order 0..3.  freq: [n_grids, N], tau: [n_grids, K] (or [K]); L: [n_grids, N, K]. */
int bdrt_build_L(bdrt_ctx* ctx, const double* freq, int n_grids, int N, const double* tau, int K,
                 int tau_per_grid, double epsilon, int order, double* L);

/* construct_M(frequencies, order, epsilon): closed-form integral-penalty matrix, order 0..2.
 * toeplitz != 0 reproduces the reference's symmetric-Toeplitz shortcut M = toeplitz(first column)
 * (matrices.py:396-405; the caller evaluates utils.is_loguniform).  freq: [n_grids, K]; M: [n_grids, K, K]. */
int bdrt_build_M(bdrt_ctx* ctx, const double* freq, int n_grids, int K, double epsilon, int order, int toeplitz,
                 double* M);

/* ---- hierarchical-Bayes model data (what Inverter._prep_stan_data packs, inversion.py:1684-1880) -------------
 * One bdrt_series_data describes a batch of B spectra fitted with one Stan program of the 'Series' family
 * (single DRT).  Matrices are those of _prep_stan_data: A = [A_re; A_im] stacked (2Nf x K), Z = [Z'; Z''] of the
 * *scaled* spectrum, L0/L1/L2 already multiplied by the per-mode constants. */
/* Structure of the matrices of a problem, found once by bdrt_series_analyze (which waits for the device) and handed
 * back through bdrt_series_data.info: with it every solver call below only enqueues work on the context's stream; without
 * it (info == NULL) each call repeats the analysis and synchronises the stream twice while doing so. */
#define BDRT_INFO_TAPS 13 /* Toeplitz taps kept per penalty matrix: |n - m| <= 6 (the reference's default epsilon) */
typedef struct {
  int valid;    /* set by bdrt_series_analyze */
  int toepA;    /* every kernel matrix is Toeplitz in each part (log-uniform grids of equal spacing, matrices.py:145-242) */
  int bw[3];    /* per distribution: half bandwidth of L0/L1/L2 (entries below 1e-18 of the largest dropped) */
  int toepL[3]; /* per distribution: L0/L1/L2 are Toeplitz */
  double taps[3][3][BDRT_INFO_TAPS]; /* per distribution and derivative order: the taps L[K/2][K/2 - 6 .. K/2 + 6] */
} bdrt_series_info;

typedef struct {
  int model;  /* BDRT_MODEL_SERIES [| BDRT_MODEL_POS] [| BDRT_MODEL_OUTLIERS] */
  int Nf;     /* measured frequencies per spectrum */
  int K;      /* basis functions */
  int B;      /* spectra in the batch */
  int per_spectrum_grid; /* 0: A / freq shared by the batch; 1: A is [B,2Nf,K], freq is [B,Nf] */
  const double* A;    /* [2Nf,K] or [B,2Nf,K] */
  const double* Z;    /* [B,2Nf] */
  const double* freq; /* [Nf] or [B,Nf], Hz, descending (as the reference sorts, inversion.py:2138) */
  const double* L;    /* [3,K,K]: scaled L0, L1, L2 (shared by the batch) */
  double sigma_min, ups_alpha, ups_beta, induc_scale;
  double sigma_out_lambda, sigma_out_alpha, sigma_out_beta; /* outlier models only */
  /* BDRT_MODEL_SERIES_PARALLEL only (Series-Parallel[_pos]_modelcode.txt; Stan data of inversion.py:1886-1959):
   * K / A / L above describe the series distribution (Ks, As, L0s..L2s); the parallel distribution is */
  int Kp;             /* basis functions of the parallel distribution */
  const double* Ap;   /* [2Nf,Kp] or [B,2Nf,Kp] stacked kernel matrix of the parallel distribution */
  const double* Lp;   /* [3,Kp,Kp] scaled L0p, L1p, L2p */
  double x_sum_invscale, xp_scale;
  /* BDRT_MODEL_SERIES_2PARALLEL only: the second parallel distribution (the reference orders the parallel
   * distributions by sorted name, inversion.py:1963) */
  int Kp2;
  const double* Ap2;  /* [2Nf,Kp2] or [B,2Nf,Kp2] */
  const double* Lp2;  /* [3,Kp2,Kp2] */
  double xp2_scale;
  const bdrt_series_info* info; /* HOST pointer, result of bdrt_series_analyze for these matrices; NULL: analyse per call */
} bdrt_series_data;

/* Looks at A / L (/ Ap, Lp, Ap2, Lp2) of `data` on the device and fills *info (host memory).  Synchronises the context's
 * stream.  The result stays valid as long as the matrices keep their contents and shapes. */
int bdrt_series_analyze(bdrt_ctx* ctx, const bdrt_series_data* data, bdrt_series_info* info);

/* number of unconstrained parameters D of the model: Series 2K+9 (+2Nf with outliers);
 * Series-Parallel 2(Ks+Kp)+12; Series-2Parallel 2(Ks+Kp1+Kp2)+15 */
int bdrt_num_params(const bdrt_series_data* data);

/* Test hook == Stan's log_prob + grad_log_prob (what .optimizing / .sampling evaluate internally):
 * u [n_cols, D] unconstrained points; spec [n_cols] int32 spectrum index of every column (NULL: column c uses
 * spectrum c % B); jacobian 0 (optimizing) / 1 (sampling); outputs lp [n_cols], grad [n_cols, D]. */
int bdrt_logpost_grad(bdrt_ctx* ctx, const bdrt_series_data* data, const double* u, const int* spec, int n_cols,
                      int jacobian, double* lp, double* grad);

/* ---- MAP: replaces StanModel.optimizing(dat, iter=max_iter, seed, init) (inversion.py:1216) ------------------ */
typedef struct {
  int max_iter;      /* Stan 'iter' (reference passes 50000) */
  int history;       /* L-BFGS history size (Stan default 5; <= 16) */
  double init_alpha; /* 1e-3 */
  double tol_obj, tol_rel_obj, tol_grad, tol_rel_grad, tol_param; /* 1e-12, 1e4, 1e-8, 1e7, 1e-8 */
} bdrt_lbfgs_opts;

void bdrt_lbfgs_default_opts(bdrt_lbfgs_opts* o);

/* u [B, D]: in = initial unconstrained points (Stan's init), out = last iterate.
 * lp [B] (log-density without Jacobian at the result), iters [B], n_eval [B] (gradient evaluations),
 * status [B] (BDRT_TERM_*).  Any of lp / iters / n_eval / status may be NULL. */
int bdrt_map_lbfgs(bdrt_ctx* ctx, const bdrt_series_data* data, const bdrt_lbfgs_opts* opts, double* u, double* lp,
                   int* iters, int* n_eval, int* status);

/* Damped-Newton polish of MAP estimates (not in the reference: brings both sides of the parity test to the unique
 * optimum, SURVEY.md section 7 hard part 1).  u [B,D] in/out; gnorm [B] = max|grad| over the free coordinates at the
 * result.  Stops on max|grad| < gtol, on max_iter, or at the rounding floor of the gradient (six iterations that did
 * not halve the best max|grad| below 1e-6: the floor is 1e-10 .. 1e-8 depending on the spectrum).  lower=0 parameters
 * whose optimum is theta = 0 are returned at u = -40 (theta = 4e-18). */
typedef struct {
  int max_iter;   /* Newton iterations (default 200) */
  double gtol;    /* stop when max|grad| < gtol (default 1e-9) */
  double fd_step; /* relative forward-difference step of the Hessian (default 1e-6) */
} bdrt_newton_opts;
void bdrt_newton_default_opts(bdrt_newton_opts* o);
int bdrt_map_newton(bdrt_ctx* ctx, const bdrt_series_data* data, const bdrt_newton_opts* opts, double* u,
                    double* lp, double* gnorm, int* iters, int* n_eval);

/* ---- HMC: replaces StanModel.sampling(dat, warmup, iter, chains, seed, init, control) (inversion.py:1218-1221) */
typedef struct {
  int chains, warmup, samples;
  int max_treedepth;  /* 10 */
  double adapt_delta; /* 0.9 */
  double adapt_t0;    /* 10 */
  double adapt_gamma; /* 0.05 */
  double adapt_kappa; /* 0.75 */
  unsigned long long seed;
  long long spectrum_offset; /* global index of spectrum 0 of this batch (Philox key; makes results independent of
                                how the batch is sharded over GPUs) */
  const long long* spectrum_ids; /* optional device array [B] of global spectrum indices (overrides
                                    spectrum_offset + b); NULL for a contiguous batch */
} bdrt_nuts_opts;
void bdrt_nuts_default_opts(bdrt_nuts_opts* o);

/* u0 [B, chains, D] initial points.  Outputs (any may be NULL):
 *   draws    [B, chains, samples, D]  post-warm-up draws, unconstrained space
 *   stepsize [B, chains]  adapted step size;  n_leapfrog [B, chains] total gradient evaluations
 *   n_divergent [B, chains];  n_maxdepth [B, chains] (post-warm-up iterations that saturated max_treedepth)
 *   accept [B, chains] mean post-warm-up accept_stat */
int bdrt_nuts(bdrt_ctx* ctx, const bdrt_series_data* data, const bdrt_nuts_opts* opts, const double* u0,
              double* draws, double* stepsize, long long* n_leapfrog, int* n_divergent, int* n_maxdepth,
              double* accept);

/* Unconstrained draws -> constrained / transformed parameters the reference reads back
 * (_extract_parameter, inversion.py:2494-2519): out [n, P] with P = bdrt_num_outputs():
 * [x(K) | Rinf | induc | sigma_res | alpha_prop | alpha_re | alpha_im | sigma_tot(2Nf) | sigma_out(Nf if outliers)]
 * (Series-Parallel: x(K) is replaced by [xs(Ks) | xp(Kp)], xp = xp_raw * xp_scale);  spec as in bdrt_logpost_grad. */
int bdrt_num_outputs(const bdrt_series_data* data);
int bdrt_constrain(bdrt_ctx* ctx, const bdrt_series_data* data, const double* u, const int* spec, int n,
                   double* out);

/* Posterior summaries of (constrained) draws, on device: replaces np.mean(axis=0) in Inverter._extract_parameter
 * (inversion.py:2514-2519) and np.percentile(axis=0) (linear interpolation) in coef_percentile / predict_Z /
 * predict_sigma (inversion.py:2560, :2702, :3096).
 * draws [G, S, P]: G groups (spectra), S merged draws, P parameters; probs_host [nq] HOST array of probabilities in
 * [0, 1] (percentile / 100, nq <= 64); outputs mean [G, P] and quant [nq, G, P] (either may be NULL). */
int bdrt_summarize(bdrt_ctx* ctx, const double* draws, int G, int S, int P, const double* probs_host, int nq,
                   double* mean, double* quant);

/* Sampler diagnostics, on device: split R-hat and rank-normalised split-chain bulk ESS (Vehtari et al. 2021) of every
 * column -- the numbers pystan prints after StanModel.sampling (inversion.py:1218-1221; SURVEY.md section 8b lists
 * them among the outputs of the sampler) and the ESS behind the benchmark's ESS/s metric.
 * draws [G, chains * n, P] (draw index = chain * n + i, i.e. the layout of bdrt_constrain applied to bdrt_nuts draws);
 * outputs rhat [G, P], ess_bulk [G, P] (either may be NULL).  n >= 4. */
int bdrt_diagnostics(bdrt_ctx* ctx, const double* draws, int G, int chains, int n, int P, double* rhat,
                     double* ess_bulk);

/* ---- hyper-parametric ridge: replaces cvxopt.solvers.qp inside Inverter._convex_opt (inversion.py:1043-1067)
 *      and the hyper-lambda loop of Inverter.ridge_fit (inversion.py:489-753) ---------------------------------- */
/* Batched bound-constrained strictly convex QP   min 1/2 x'Px + q'x  s.t. x >= lb.
 * P [B,n,n] (symmetric; only read), q [B,n], lb [n] (shared), x [B,n] out; kkt [B] max KKT residual, iters [B]. */
int bdrt_qp_bound(bdrt_ctx* ctx, const double* P, const double* q, const double* lb, int B, int n, double* x,
                  double* kkt, int* iters);

typedef struct {
  int penalty;      /* 0 'discrete' (L'L), 1 'integral' (M) */
  int nonneg;       /* 1: c >= 0 ; 0: c[0:2] >= 0, c[2:] >= -10 (inversion.py:1054-1064) */
  int max_iter;     /* 20 */
  double xtol;      /* 1e-3 */
  double hl_beta;   /* 2.5 */
  double lambda_0;  /* 1e-2 */
  double reg_ord[3];/* fraction of each derivative order (reg_ord=2 -> 0, 0, 1) */
  double L1_penalty;/* 0 */
  double epsilon;   /* basis epsilon (only enters the L1 penalty vector) */
  int fit_inductance;
  double hl_fbeta;  /* > 0: lambda_k = lambda_0 / ((L c)_k^2 / (max_k (L c)_k^2 hl_fbeta) + 1) instead of the hl_beta rule
                     * (discrete penalty, _hyper_lambda_fbeta inversion.py:956-964; preset 'Ciucci' uses 0.1); 0: off */
  int stop_rule;    /* stop test mean(|(c - c_prev) / c_prev|) < xtol (inversion.py:730-736) when a coefficient is 0 in
                     * two consecutive iterations (0/0): 0 = numpy semantics, NaN never passes -- what the reference's own
                     * code does once the QP solver returns exact zeros on the bound (cvxopt's interior-point iterates
                     * never do, SURVEY.md section 7 hard part 1b); 1 (default) = such a coefficient counts as unchanged */
} bdrt_ridge_opts;
void bdrt_ridge_default_opts(bdrt_ridge_opts* o);

/* Hyper-lambda ridge fit of B spectra sharing one grid.
 * WA_re, WA_im [B?,Nf,K+2] weighted augmented matrices ([1|0|A_re], [0|2 pi f 1e-4|A_im], inversion.py:401-421)
 *   -- per_spectrum_W = 0: one [Nf,K+2] pair shared by the batch (unity weights), 1: [B,Nf,K+2];
 * WZ_re, WZ_im [B,Nf] weighted scaled targets;
 * Pen [3,K+2,K+2] zero-padded base penalty matrices per derivative order (L_o'L_o or M_o, inversion.py:424-447);
 * Lmat [3,K,K+2] zero-padded L_o (discrete penalty only, for the lambda update inversion.py:947-954; may be NULL
 *   for the integral penalty);
 * outputs: coef [B,K+2] (scaled units, before the rescaling of inversion.py:875-894), lam [B,3,K+2],
 *   iters [B] hyper-iterations, converged [B], n_factor [B] Cholesky factorisations (pivot steps of the QP solver)
 *   spent on the spectrum; iters / converged / n_factor may be NULL.  Any Nf. */
int bdrt_ridge_fit(bdrt_ctx* ctx, const bdrt_ridge_opts* opts, const double* WA_re, const double* WA_im,
                   int per_spectrum_W, const double* WZ_re, const double* WZ_im, const double* Pen,
                   const double* Lmat, int B, int Nf, int K, double* coef, double* lam, int* iters, int* converged,
                   int* n_factor);

/* ---- micro-benchmarks used for the roofline denominators (bench.py) ------------------------------------------ */
/* Achieved FP64 FMA and FP64 DMMA (mma.sync.m8n8k4) throughput in TFLOP/s, measured with CUDA events. */
int bdrt_peak_fp64(bdrt_ctx* ctx, double* dfma_tflops_host, double* dmma_tflops_host);

#ifdef __cplusplus
}
#endif
#endif /* BDRT_H */
