"""Batched, GPU-resident drop-in for the reference's ``Inverter`` on the inversion hot path.

Mirrors ``bayes_drt.inversion.Inverter`` (inversion.py:28-4110): same constructor, ``fit`` / ``ridge_fit`` signatures,
result attributes (``distribution_fits[name]['coef']``, ``R_inf``, ``inductance``, ``error_fit``, ``fit_type``) and
post-fit queries (``coef_percentile``, ``predict_distribution``, ``predict_Z``, ``predict_Rp``, ``predict_sigma``).

Differences that come with batching:
  * ``Z`` may be ``[Nf]`` (one spectrum, reference shapes come back as numpy arrays) or ``[B, Nf]`` (a batch sharing
    one frequency grid; results are CUDA tensors with a leading batch dimension);
  * everything numeric runs in the CUDA library (no CPU fallback); options the CUDA path does not implement raise
    ``NotImplementedError`` instead of silently taking another route.
"""
import warnings

import numpy as np
import torch

from . import capi
from . import matrices as mat
from ._lib import context

# Inverter._prep_stan_data constants (inversion.py:1725-1737, :1873-1880)
_MODE = {
    'sample': dict(ups_alpha=1.0, ups_beta=0.1, l=(1.0, 1.0, 0.75), sigma_out_alpha=5.0),
    'optimize': dict(ups_alpha=0.05, ups_beta=0.1, l=(1.5 * 0.24, 1.5 * 0.16, 1.5 * 0.08), sigma_out_alpha=2.0),
}


def _hash_uniform(seed, start, n, D, device, lo=-2.0, hi=2.0):
    """U(lo, hi) initial points keyed by the *global* spectrum index (splitmix64), so a spectrum gets the same init
    whatever the sharding (Stan: init='random' draws U(-2, 2) per unconstrained coordinate [Stan-upstream])."""
    idx = (torch.arange(start, start + n, dtype=torch.int64, device=device)[:, None] * D
           + torch.arange(D, dtype=torch.int64, device=device)[None, :])
    z = idx + (int(seed) * 0x9E3779B97F4A7C15 + 0x632BE59BD9B4E019) % (1 << 63)
    for sh, mul in ((30, 0xBF58476D1CE4E5B9), (27, 0x94D049BB133111EB)):
        z = (z ^ ((z >> sh) & ((1 << (64 - sh)) - 1))) * (mul - (1 << 64) if mul >= (1 << 63) else mul)
    z = z ^ ((z >> 31) & ((1 << 33) - 1))
    u = ((z >> 11) & ((1 << 53) - 1)).to(torch.float64) * (1.0 / (1 << 53))
    return lo + (hi - lo) * u


class Inverter:
    def __init__(self, basis_freq=None, basis='gaussian', epsilon=None, fit_inductance=True,
                 distributions={'DRT': {'kernel': 'DRT'}}, device=None):
        self.device = context(device).device
        self._recalc_mat = True
        self.distribution_matrices = {}
        self.basis_freq = basis_freq
        self.basis = basis
        self.epsilon = epsilon
        self.fit_inductance = fit_inductance
        self.set_distributions({k: dict(v) for k, v in distributions.items()})
        self.f_train = None
        self.Z_train = None
        self._Z_scale = 1.0
        self._init_params = {}
        self.distribution_fits = {}
        self.error_fit = {}
        self.fit_type = None
        self._single = False
        self._mat_key = None

    # ------------------------------------------------------------------------------------------------------------
    # configuration (inversion.py:66-129, :4069-4110)
    # ------------------------------------------------------------------------------------------------------------
    def set_distributions(self, distributions):
        for name, info in distributions.items():
            if info['kernel'] == 'DRT':
                if info.get('dist_type', 'series') != 'series':
                    warnings.warn(f"dist_type for DRT kernel must be series. Overwriting supplied dist_type for "
                                  f"distribution '{name}' with 'series'")
                info['dist_type'] = 'series'
                invalid = [k for k in ('symmetry', 'bc', 'ct', 'k_ct') if k in info]
                if invalid:
                    warnings.warn(f"The following keys are invalid for distribution '{name}': {invalid}. "
                                  f"These keys will be ignored")
            elif info['kernel'] == 'DDT':
                if info.get('dist_type', 'parallel') not in ['series', 'parallel']:
                    raise ValueError(f"Invalid dist_type '{info.get('dist_type', 'NA')}' for distribution '{name}'")
                elif info.get('symmetry', 'planar') not in ['planar', 'spherical']:
                    raise ValueError(f"Invalid symmetry '{info.get('symmetry', 'NA')}' for distribution '{name}'")
                elif info.get('bc', 'transmissive') not in ['transmissive', 'blocking']:
                    raise ValueError(f"Invalid bc '{info.get('bc', 'NA')}' for distribution '{name}'")
                elif info.get('ct', True) not in [True, False]:
                    raise ValueError(f"Invalid ct {info['ct']} for distribution '{name}'")
                if info.get('ct', False) and 'k_ct' not in info:
                    raise ValueError(f"k_ct must be supplied for distribution '{name}' if ct==True")
                defaults = {'dist_type': 'parallel', 'symmetry': 'planar', 'bc': 'blocking', 'ct': False}
                defaults.update(info)
                distributions[name] = defaults
            else:
                raise ValueError(f"Invalid kernel {info['kernel']}. Options are DRT and DDT")
            self.distribution_matrices.setdefault(name, {})
        self._distributions = distributions
        self._recalc_mat = True

    distributions = property(lambda self: self._distributions, set_distributions)

    def set_basis(self, basis):
        if basis != 'gaussian':
            raise ValueError(f'Invalid basis {basis}. Options are gaussian')  # inversion.py:38-39, matrices.py:22-23
        self._basis = basis
        self._recalc_mat = True

    basis = property(lambda self: self._basis, set_basis)

    # ------------------------------------------------------------------------------------------------------------
    # preprocessing (Inverter._prep_matrices, inversion.py:2127-2336; _scale_Z :2411-2443)
    # ------------------------------------------------------------------------------------------------------------
    def _to_batch(self, frequencies, Z):
        f = torch.as_tensor(np.asarray(frequencies) if not torch.is_tensor(frequencies) else frequencies,
                            dtype=torch.float64)
        Zt = torch.as_tensor(np.asarray(Z) if not torch.is_tensor(Z) else Z)
        if not Zt.is_complex():
            raise ValueError('Z must be complex')
        Zt = Zt.to(torch.complex128)
        self._single = Zt.dim() == 1
        if self._single:
            Zt = Zt[None, :]
        if f.dim() != 1:
            raise NotImplementedError('per-spectrum frequency grids are not implemented in Inverter yet '
                                      '(use capi.build_A for batched grids)')
        if f.shape[0] != Zt.shape[1]:
            raise ValueError('Length of frequencies and Z must be equal')  # inversion.py:2128-2129
        # sort by descending frequency (inversion.py:2138-2141)
        idx = torch.argsort(f, descending=True)
        f = f[idx].contiguous()
        Zt = Zt.to(self.device)[:, idx.to(self.device)].contiguous()
        return f, Zt

    def _scale_Z(self, Z, scale_Z):
        if not scale_Z:
            self._Z_scale = torch.ones(Z.shape[0], dtype=torch.float64, device=Z.device)
            return Z
        # series / mixed branch (inversion.py:2437-2441): std(|Z|) / sqrt(Nf / 81), population std like np.std
        self._Z_scale = Z.abs().std(dim=1, unbiased=False) / np.sqrt(Z.shape[1] / 81)
        return Z / self._Z_scale[:, None]

    def _grid(self, freq, name):
        """tau, epsilon of a distribution (inversion.py:2191-2209) and its cached kernel / penalty matrices."""
        info = self.distributions[name]
        bf = info.get('basis_freq', self.basis_freq)
        fn = freq.numpy()
        if bf is None:
            tmin = np.log10(1 / (2 * np.pi * np.max(fn))) - 1
            tmax = np.log10(1 / (2 * np.pi * np.min(fn))) + 1
            tau = np.logspace(tmin, tmax, int(10 * (tmax - tmin) + 1))
        else:
            tau = 1 / (2 * np.pi * np.asarray(torch.as_tensor(bf).cpu(), dtype=np.float64))
        eps = info.get('epsilon', self.epsilon)
        if eps is None:
            eps = 1 / np.mean(np.diff(np.log(tau)))
        info['tau'], info['epsilon'] = tau, float(eps)
        key = (name, fn.tobytes(), tau.tobytes(), float(eps), info['kernel'], info['dist_type'],
               info.get('symmetry'), info.get('bc'), info.get('ct', False), info.get('k_ct'))
        m = self.distribution_matrices.setdefault(name, {})
        if m.get('_key') != key:
            t = torch.as_tensor(tau)
            A_re, A_im = capi.build_A(freq, t, eps, kernel=info['kernel'], dist_type=info['dist_type'],
                                      symmetry=info.get('symmetry') or 'planar', bc=info.get('bc') or 'transmissive',
                                      ct=info.get('ct', False), k_ct=info.get('k_ct'), device=self.device)
            bft = torch.as_tensor(1 / (2 * np.pi * tau))
            m.clear()
            m.update(_key=key, A_re=A_re, A_im=A_im,
                     L0=capi.build_L(bft, t, eps, 0, device=self.device),
                     L1=capi.build_L(bft, t, eps, 1, device=self.device),
                     L2=capi.build_L(bft, t, eps, 2, device=self.device))
        self._recalc_mat = False
        return tau, float(eps), m

    # ------------------------------------------------------------------------------------------------------------
    # hierarchical-Bayes fit (Inverter.fit, inversion.py:1072-1289)
    # ------------------------------------------------------------------------------------------------------------
    def fit(self, frequencies, Z, part='both', scale_Z=True, nonneg=False, outliers=False, check_outliers=True,
            init_from_ridge=False, ridge_kw={}, sigma_min=0.002, inductance_scale=1, outlier_lambda=None,
            mode='optimize', random_seed=1234, max_iter=50000, warmup=200, samples=200, chains=2,
            add_stan_data={}, model_str=None, fitY=False, SA=False, SASY=False,
            init=None, polish=False, spectrum_offset=0, keep_draws=True):
        """Same arguments as the reference plus: ``init`` (explicit unconstrained initial points [B, D] or
        [B, chains, D]; Stan accepts an init dict the same way), ``polish`` (damped-Newton refinement of the MAP
        estimate to the exact optimum), ``spectrum_offset`` (global index of the first spectrum, for sharded batches),
        ``keep_draws`` (keep the HMC draws for percentile queries)."""
        if part != 'both':
            raise NotImplementedError("part != 'both' is not implemented (and is inconsistent in the reference's "
                                      "Series models, inversion.py:1721-1723)")
        if fitY or SA or SASY:
            raise NotImplementedError('fitY / SA / SASY are not implemented')
        if model_str is not None or add_stan_data:
            raise NotImplementedError('model_str / add_stan_data (Stan escape hatches) are not available')
        if mode not in ('optimize', 'sample'):
            raise ValueError(f"Invalid mode {mode}. Options are 'optimize', 'sample'")
        if len(self.distributions) != 1:
            raise NotImplementedError('multi-distribution (Series-Parallel / Series-2Parallel) models are not '
                                      'implemented in this build')
        name = list(self.distributions.keys())[0]
        info = self.distributions[name]
        if info['dist_type'] != 'series' or info['kernel'] != 'DRT':
            raise NotImplementedError("only the single-DRT 'Series' model family is implemented in this build")
        if outliers == 'auto' or init_from_ridge:
            raise NotImplementedError("outliers='auto' / init_from_ridge need the ridge initialisation chain, "
                                      "which this build does not wire up yet")
        freq, Zb = self._to_batch(frequencies, Z)
        self.f_train = freq.numpy()
        self.Z_train = Zb
        Zs = self._scale_Z(Zb, scale_Z)
        tau, eps, m = self._grid(freq, name)
        c = _MODE[mode]
        L = torch.stack([c['l'][0] * m['L0'], c['l'][1] * m['L1'], c['l'][2] * m['L2']])
        Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
        prob = capi.SeriesProblem(torch.cat((m['A_re'], m['A_im'])), Zst, freq, L, nonneg=bool(nonneg),
                                  outliers=bool(outliers), sigma_min=sigma_min, ups_alpha=c['ups_alpha'],
                                  ups_beta=c['ups_beta'], induc_scale=float(inductance_scale),
                                  sigma_out_lambda=10.0 if outlier_lambda is None else float(outlier_lambda),
                                  sigma_out_alpha=c['sigma_out_alpha'], sigma_out_beta=1.0, device=self.device)
        self._problem = prob
        # the reference's model file name (Inverter._get_stan_model, inversion.py:1576-1610)
        self.stan_model_name = 'Series' + ('_pos' if nonneg else '') + ('_outliers' if outliers else '') \
            + '_StanModel.pkl'
        B, D = prob.B, prob.D
        self.distribution_fits, self.error_fit = {}, {}
        if mode == 'optimize':
            u0 = _hash_uniform(random_seed, spectrum_offset, B, D, self.device) if init is None else \
                torch.as_tensor(init, dtype=torch.float64, device=self.device).reshape(B, D)
            r = prob.map_lbfgs(u0, max_iter=max_iter)
            if polish:
                p = prob.map_newton(r['u'])
                r.update(u=p['u'], lp=p['lp'], gnorm=p['gnorm'], newton_iters=p['iters'])
            self._opt_result = r
            out = prob.split_outputs(prob.constrain(r['u']))
            point = out
            self.fit_type = 'map'
            self._sample_result = None
        else:
            u0 = _hash_uniform(random_seed, spectrum_offset * chains, B * chains, D, self.device).reshape(
                B, chains, D) if init is None else torch.as_tensor(init, dtype=torch.float64,
                                                                    device=self.device).reshape(B, chains, D)
            r = prob.nuts(u0, chains=chains, warmup=warmup, samples=samples, seed=random_seed,
                          spectrum_offset=spectrum_offset)
            spec = torch.arange(B, dtype=torch.int32, device=self.device).repeat_interleave(chains * samples)
            cons = prob.constrain(r['draws'].reshape(B * chains * samples, D), spec=spec).reshape(
                B, chains * samples, prob.P)
            draws = prob.split_outputs(cons)
            # posterior mean over the merged chains (Inverter._extract_parameter, inversion.py:2514-2519)
            point = {k: v.mean(dim=1) for k, v in draws.items()}
            self._sample_result = draws if keep_draws else None
            self._sample_stats = {k: r[k] for k in ('stepsize', 'n_leapfrog', 'n_divergent', 'n_maxdepth', 'accept')}
            if not keep_draws:
                r['draws'] = None
            self.fit_type = 'bayes'
        s = self._Z_scale
        self.distribution_fits[name] = {'coef': point['x'] * s[:, None]}
        self.R_inf = point['Rinf'] * s
        self.inductance = point['induc'] * s
        self.error_fit['sigma_min'] = sigma_min * s
        self.error_fit['sigma_tot'] = point['sigma_tot'] * s[:, None]
        self.error_fit['sigma_res'] = point['sigma_res'] * s
        for k in ('alpha_prop', 'alpha_re', 'alpha_im'):
            self.error_fit[k] = point[k]
        if outliers:
            self.error_fit['sigma_out'] = point['sigma_out'] * s[:, None]
        if check_outliers and not outliers:
            idx = self.check_outliers(threshold=3.5)
            if self._single and len(idx) > 0:
                warnings.warn(f'Possible outliers were identified at indices {idx}. Check the residuals and consider '
                              f're-running with outliers=True')
        if self._single:
            self._squeeze()
        return self

    def _squeeze(self):
        """single-spectrum call: reference shapes as numpy arrays"""
        def sq(t):
            return t[0].cpu().numpy() if torch.is_tensor(t) and t.dim() > 0 else t
        for nm in self.distribution_fits:
            self.distribution_fits[nm] = {k: sq(v) for k, v in self.distribution_fits[nm].items()}
        self.R_inf, self.inductance = float(sq(self.R_inf)), float(sq(self.inductance))
        self.error_fit = {k: (float(sq(v)) if sq(v).ndim == 0 else sq(v)) for k, v in self.error_fit.items()}

    def _coef_batch(self, name):
        c = self.distribution_fits[name]['coef']
        c = torch.as_tensor(c, dtype=torch.float64, device=self.device)
        return c[None, :] if c.dim() == 1 else c

    def _ret(self, t):
        return t[0].cpu().numpy() if self._single else t

    # ------------------------------------------------------------------------------------------------------------
    # post-fit queries
    # ------------------------------------------------------------------------------------------------------------
    def coef_percentile(self, distribution_name, percentile):
        """inversion.py:2547-2566: per-coefficient np.percentile (linear interpolation) of the merged draws."""
        if self.fit_type != 'bayes' or self._sample_result is None:
            raise ValueError('Percentile prediction is only available for bayes_fit')
        x = self._sample_result['x']
        q = torch.quantile(x, percentile / 100.0, dim=1, interpolation='linear') * self._Z_scale[:, None]
        return self._ret(q)

    def predict_distribution(self, name=None, eval_tau=None, percentile=None, time=None):
        """inversion.py:3162 (generic branch :3298-3311): F = Phi @ coef, Phi[i, m] = exp(-(eps ln(eval_tau_i/tau_m))^2)."""
        if time is not None:
            raise NotImplementedError('drift fits are out of scope')
        if name is None:
            name = list(self.distributions.keys())[0]
        info = self.distributions[name]
        basis_tau = torch.as_tensor(info['tau'], dtype=torch.float64, device=self.device)
        et = basis_tau if eval_tau is None else torch.as_tensor(eval_tau, dtype=torch.float64, device=self.device)
        coef = self._coef_batch(name) if percentile is None else \
            torch.as_tensor(self.coef_percentile(name, percentile), device=self.device).reshape(-1, len(basis_tau))
        phi = torch.exp(-(info['epsilon'] * torch.log(et[:, None] / basis_tau[None, :])) ** 2)
        return self._ret(coef @ phi.T)

    def predict_Z(self, frequencies, times=None, distributions=None, include_offsets=True, percentile=None):
        """inversion.py:2669 (generic branch :2942-2959) for the single series distribution."""
        if times is not None:
            raise NotImplementedError('drift fits are out of scope')
        name = list(self.distributions.keys())[0]
        info = self.distributions[name]
        f = torch.as_tensor(np.asarray(frequencies, dtype=np.float64))
        A_re, A_im = capi.build_A(f, torch.as_tensor(info['tau']), info['epsilon'], device=self.device)
        fd = f.to(self.device)
        if percentile is not None:
            if self.fit_type != 'bayes' or self._sample_result is None:
                raise ValueError('Percentile prediction is only available for bayes_fit results')
            s = self._Z_scale[:, None, None]
            x = self._sample_result['x'] * s
            Zr = x @ A_re.T
            Zi = x @ A_im.T
            if include_offsets:
                Zr = Zr + (self._sample_result['Rinf'][..., None] * s)
                Zi = Zi + 2 * np.pi * fd * (self._sample_result['induc'][..., None] * s)
            Zp = torch.complex(torch.quantile(Zr, percentile / 100.0, dim=1),
                               torch.quantile(Zi, percentile / 100.0, dim=1))
            return self._ret(Zp)
        coef = self._coef_batch(name)
        Zr, Zi = coef @ A_re.T, coef @ A_im.T
        if include_offsets:
            Rinf = torch.as_tensor(self.R_inf, dtype=torch.float64, device=self.device).reshape(-1, 1)
            ind = torch.as_tensor(self.inductance, dtype=torch.float64, device=self.device).reshape(-1, 1)
            Zr = Zr + Rinf
            Zi = Zi + 2 * np.pi * fd * ind
        return self._ret(torch.complex(Zr, Zi))

    def predict_Rp(self, distributions=None, percentile=None, time=None):
        """inversion.py:3033: area under the DRT, sum(coef) sqrt(pi) / epsilon."""
        name = list(self.distributions.keys())[0]
        eps = self.distributions[name]['epsilon']
        if percentile is None:
            rp = self._coef_batch(name).sum(dim=1) * np.pi ** 0.5 / eps
        else:
            if self.fit_type != 'bayes' or self._sample_result is None:
                raise ValueError('Percentile prediction is only available for bayes_fit results')
            arr = (self._sample_result['x'] * self._Z_scale[:, None, None]).sum(dim=2) * np.pi ** 0.5 / eps
            rp = torch.quantile(arr, percentile / 100.0, dim=1)
        return float(rp[0]) if self._single else rp

    def predict_sigma(self, frequencies=None, percentile=None, times=None):
        """inversion.py:3089 for the training frequencies: the fitted sigma_tot split into real / imaginary parts."""
        if self.fit_type not in ('map', 'bayes'):
            raise ValueError('Error scale prediction only available for bayes_fit and map_fit')
        if frequencies is not None and not np.array_equal(mat.rel_round(np.asarray(frequencies), 10),
                                                           mat.rel_round(self.f_train, 10)):
            raise NotImplementedError('predict_sigma at frequencies other than the training grid is not implemented')
        if percentile is not None:
            if self.fit_type != 'bayes' or self._sample_result is None:
                raise ValueError('Percentile prediction is only available for bayes_fit')
            st = torch.quantile(self._sample_result['sigma_tot'], percentile / 100.0, dim=1) * self._Z_scale[:, None]
        else:
            st = torch.as_tensor(self.error_fit['sigma_tot'], device=self.device).reshape(-1, 2 * len(self.f_train))
        nf = len(self.f_train)
        return self._ret(st[:, :nf]), self._ret(st[:, nf:])

    def check_outliers(self, frequencies=None, Z=None, threshold=3.5, use_existing_fit=True, **ridge_kw):
        """inversion.py:3313-3376, existing MAP / HMC fit branch: combined z-score of the residuals under the fitted
        error model.  Returns indices ([n] for one spectrum, [n, 2] (spectrum, frequency) for a batch)."""
        if self.fit_type not in ('map', 'bayes'):
            raise NotImplementedError('check_outliers from a ridge fit needs ridge_fit (not wired up in this build)')
        name = list(self.distributions.keys())[0]
        m = self.distribution_matrices[name]
        coef = torch.as_tensor(self.distribution_fits[name]['coef'], device=self.device).reshape(-1, m['A_re'].shape[1])
        Rinf = torch.as_tensor(self.R_inf, dtype=torch.float64, device=self.device).reshape(-1, 1)
        ind = torch.as_tensor(self.inductance, dtype=torch.float64, device=self.device).reshape(-1, 1)
        fd = torch.as_tensor(self.f_train, device=self.device)
        err_re = coef @ m['A_re'].T + Rinf - self.Z_train.real
        err_im = coef @ m['A_im'].T + 2 * np.pi * fd * ind - self.Z_train.imag
        st = torch.as_tensor(self.error_fit['sigma_tot'], device=self.device).reshape(-1, 2 * len(self.f_train))
        nf = len(self.f_train)
        zs = torch.sqrt(((err_re / st[:, :nf]) ** 2 + (err_im / st[:, nf:]) ** 2) / 2)
        idx = torch.nonzero(zs > threshold)
        return idx[:, 1].cpu().numpy() if self._single else idx

    def ridge_fit(self, frequencies, Z, **kw):
        from .ridge import ridge_fit
        return ridge_fit(self, frequencies, Z, **kw)
