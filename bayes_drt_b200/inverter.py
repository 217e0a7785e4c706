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


# Series-Parallel constants (inversion.py:1913-1930): note the parallel L0 multiplier 1.5 * 0.36 in 'optimize'
_MODE_SP = {
    'sample': dict(ls=(1.0, 1.0, 0.75), lp=(1.0, 1.0, 0.75), x_sum_invscale=1.0),
    'optimize': dict(ls=(1.5 * 0.24, 1.5 * 0.16, 1.5 * 0.08), lp=(1.5 * 0.36, 1.5 * 0.16, 1.5 * 0.08),
                     x_sum_invscale=0.0),
}


def _hash_uniform(seed, start, n, D, device, lo=-2.0, hi=2.0, ids=None):
    """U(lo, hi) initial points keyed by the *global* spectrum index (splitmix64), so a spectrum gets the same init
    whatever the sharding (Stan: init='random' draws U(-2, 2) per unconstrained coordinate [Stan-upstream]).
    ``ids``: explicit global row indices (overrides start .. start + n)."""
    rows = torch.arange(start, start + n, dtype=torch.int64, device=device) if ids is None else \
        torch.as_tensor(ids, dtype=torch.int64, device=device)
    idx = rows[:, None] * D + torch.arange(D, dtype=torch.int64, device=device)[None, :]
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

    def get_distributions(self):
        return self._distributions

    def get_basis(self):
        return self._basis

    def get_basis_freq(self):
        return self._basis_freq

    def set_basis_freq(self, basis_freq):
        self._basis_freq = basis_freq
        self._recalc_mat = True

    basis_freq = property(get_basis_freq, set_basis_freq)

    def get_epsilon(self):
        return self._epsilon

    def set_epsilon(self, epsilon, override_distributions=False):
        self._epsilon = epsilon
        self._recalc_mat = True
        if override_distributions:  # inversion.py:4095-4097
            for name in self.distributions.keys():
                self.distributions[name]['epsilon'] = epsilon

    epsilon = property(get_epsilon, set_epsilon)

    def get_fit_inductance(self):
        return self._fit_inductance_

    def set_fit_inductance(self, fit_inductance):
        self._fit_inductance_ = fit_inductance

    fit_inductance = property(get_fit_inductance, set_fit_inductance)

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
        if f.dim() == 2:
            # one grid per spectrum ([B, Nf]): every row sorted on its own
            if self._single or tuple(f.shape) != tuple(Zt.shape):
                raise ValueError('Length of frequencies and Z must be equal')  # inversion.py:2128-2129
            idx = torch.argsort(f, dim=1, descending=True)
            f = torch.gather(f, 1, idx).contiguous()
            Zt = torch.gather(Zt.to(self.device), 1, idx.to(self.device)).contiguous()
            return f, Zt
        if f.dim() != 1:
            raise ValueError('frequencies must be [Nf] (shared grid) or [B, Nf] (one grid per spectrum)')
        if f.shape[0] != Zt.shape[1]:
            raise ValueError('Length of frequencies and Z must be equal')  # inversion.py:2128-2129
        # sort by descending frequency (inversion.py:2138-2141)
        idx = torch.argsort(f, descending=True)
        f = f[idx].contiguous()
        Zt = Zt.to(self.device)[:, idx.to(self.device)].contiguous()
        return f, Zt

    def _scale_Z(self, Z, scale_Z, fit_type='ridge'):
        if not scale_Z:
            self._Z_scale = torch.ones(Z.shape[0], dtype=torch.float64, device=Z.device)
            return Z
        infos = list(self.distributions.values())
        if len(infos) == 1 and infos[0]['dist_type'] == 'parallel' and fit_type != 'ridge' \
                and infos[0]['kernel'] == 'DDT' and infos[0].get('symmetry', 'planar') == 'planar':
            # pure parallel planar DDT (inversion.py:2417-2434): scale so that the scaled admittance has a fixed std
            Ystar_std = 14.0 if infos[0]['bc'] == 'transmissive' else 2.4
            self._Z_scale = Ystar_std * np.sqrt(Z.shape[1] / 81) / (1 / Z).abs().std(dim=1, unbiased=False)
            return Z / self._Z_scale[:, None]
        # series / mixed branch (inversion.py:2437-2441): std(|Z|) / sqrt(Nf / 81), population std like np.std
        self._Z_scale = Z.abs().std(dim=1, unbiased=False) / np.sqrt(Z.shape[1] / 81)
        return Z / self._Z_scale[:, None]

    def _grid(self, freq, name):
        """tau, epsilon of a distribution (inversion.py:2191-2209) and its cached kernel / penalty matrices."""
        info = self.distributions[name]
        bf = info.get('basis_freq', self.basis_freq)
        fn = freq.numpy()
        if bf is None and fn.ndim == 2:
            # default basis of every spectrum (inversion.py:2191-2199 applied row by row).  The penalty matrices are
            # shared by the batch, so the rows' bases must be log-shifts of one another (same K, same spacing).
            tmin = np.log10(1 / (2 * np.pi * fn.max(axis=1))) - 1
            tmax = np.log10(1 / (2 * np.pi * fn.min(axis=1))) + 1
            Ks = (10 * (tmax - tmin) + 1).astype(int)
            if np.any(Ks != Ks[0]) or np.ptp(tmax - tmin) > 1e-9 * abs(tmax[0] - tmin[0]):
                raise NotImplementedError('per-spectrum grids with default bases must span the same number of '
                                          'decades (pass a shared basis_freq otherwise)')
            tau = np.stack([np.logspace(a, b, int(Ks[0])) for a, b in zip(tmin, tmax)])
        elif bf is None:
            tmin = np.log10(1 / (2 * np.pi * np.max(fn))) - 1
            tmax = np.log10(1 / (2 * np.pi * np.min(fn))) + 1
            tau = np.logspace(tmin, tmax, int(10 * (tmax - tmin) + 1))
        else:
            tau = 1 / (2 * np.pi * np.asarray(torch.as_tensor(bf).cpu(), dtype=np.float64))
        eps = info.get('epsilon', self.epsilon)
        if eps is None:
            eps = 1 / np.mean(np.diff(np.log(tau[0] if tau.ndim == 2 else tau)))
        info['tau'], info['epsilon'] = tau, float(eps)
        key = (name, fn.tobytes(), tau.tobytes(), float(eps), info['kernel'], info['dist_type'],
               info.get('symmetry'), info.get('bc'), info.get('ct', False), info.get('k_ct'))
        m = self.distribution_matrices.setdefault(name, {})
        if m.get('_key') != key:
            t = torch.as_tensor(tau)
            A_re, A_im = capi.build_A(freq, t, eps, kernel=info['kernel'], dist_type=info['dist_type'],
                                      symmetry=info.get('symmetry') or 'planar', bc=info.get('bc') or 'transmissive',
                                      ct=info.get('ct', False), k_ct=info.get('k_ct'), device=self.device)
            if tau.ndim == 2:  # L depends on ratios of the basis time constants only: the first row stands for all
                t = t[0]
            bft = 1 / (2 * np.pi * t)
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
            init=None, polish=False, spectrum_offset=0, spectrum_ids=None, keep_draws=True):
        """Same arguments as the reference plus: ``init`` (explicit unconstrained initial points [B, D] or
        [B, chains, D]; Stan accepts an init dict the same way), ``polish`` (damped-Newton refinement of the MAP
        estimate to the exact optimum), ``spectrum_offset`` (global index of the first spectrum, for sharded batches) or
        ``spectrum_ids`` (global index of every spectrum: strided shards, ``distributed.shard_indices``),
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
        model_type, ser, par = self._model_type()
        if model_type != 'Series' and outliers:
            raise NotImplementedError("Parallel_outliers / Series-Parallel*_outliers are dimensionally inconsistent as "
                                      "shipped by the reference (N override vs matrix[N,K] A) and are not implemented")
        if init_from_ridge and model_type == 'Parallel':
            raise NotImplementedError('ridge initialisation is implemented for the DRT (series) kernel only')
        if init_from_ridge and len(self.distributions) > 1:
            raise ValueError('Ridge initialization can only be performed for single-distribution fits')  # :1155-1156
        if outliers == 'auto' and model_type != 'Series':
            raise NotImplementedError("outliers='auto' is implemented for single-distribution fits")
        name = ser[0] if ser else par[0]
        freq, Zb = self._to_batch(frequencies, Z)
        single = self._single
        B = Zb.shape[0]
        ids = self._global_ids(B, spectrum_offset, spectrum_ids)
        # ---- ridge initialisation and automatic outlier detection (inversion.py:1154-1187)
        ridge_init, flags, init_flags = None, None, None
        if init_from_ridge:
            ridge_init = self._get_init_from_ridge(freq, Zb, nonneg, inductance_scale, ridge_kw)
            if outliers:  # True or 'auto': likely outliers start with a large sigma_out_raw; a less stringent threshold
                # than the model choice below, so that none is missed (inversion.py:1668-1675)
                init_flags = self._ridge_outlier_flags(freq, Zb, threshold=3, use_existing_fit=True)
        if outliers == 'auto':
            # existing ridge fit if one was just made, else a fresh one with preset='Huang'; stringent threshold 4
            flags = self._ridge_outlier_flags(freq, Zb, threshold=4, use_existing_fit=init_from_ridge, **ridge_kw)
            has = flags.any(dim=1)
            if single and bool(has[0]):
                idx = torch.nonzero(flags[0])[:, 0].cpu().numpy()
                warnings.warn('Identified likely outliers at indices {}, f={} Hz. An outlier-robust error model will '
                              'be used. To disable this behavior, pass outliers=False.'.format(idx, freq.numpy()[idx]))
            groups = [(torch.nonzero(has)[:, 0], True), (torch.nonzero(~has)[:, 0], False)]
            groups = [(ix, fl) for ix, fl in groups if ix.numel() > 0]
        else:
            groups = [(None, bool(outliers))]
        self._single = False
        self.f_train = freq.numpy()
        self.Z_train = Zb
        Zs = self._scale_Z(Zb, scale_Z, fit_type='map' if mode == 'optimize' else 'bayes')
        self._outlier_model = torch.zeros(B, dtype=torch.bool, device=self.device)
        results = []
        for ix, fl in groups:
            sel = slice(None) if ix is None else ix
            u_init = None
            if init is not None:
                u_init = torch.as_tensor(init, dtype=torch.float64, device=self.device)
                u_init = u_init.reshape(B, -1, u_init.shape[-1])[sel]
            fsel = freq if (ix is None or freq.dim() == 1) else freq[ix.cpu()]  # per-spectrum grids follow their rows
            res = self._fit_core(fsel, Zs[sel], ids[sel], model_type, name, par, mode, nonneg, fl, sigma_min,
                                 inductance_scale, outlier_lambda, random_seed, max_iter, warmup, samples, chains,
                                 u_init, None if ridge_init is None else {k: v[sel] for k, v in ridge_init.items()},
                                 init_flags[sel] if (init_flags is not None and fl) else None, polish, keep_draws)
            results.append((ix, fl, res))
            if ix is None:
                self._outlier_model[:] = fl
            else:
                self._outlier_model[ix] = fl
        if freq.dim() == 2 and len(groups) > 1:  # the sub-batches left their own grids behind: back to the whole batch
            for nm in self.distributions:
                self._grid(freq, nm)
        self._merge_results(results, B, model_type, name, par, mode, sigma_min, keep_draws)
        self.stan_model_name = model_type + ('_pos' if nonneg and ser else '') + \
            ('_outliers' if bool(self._outlier_model.any()) else '') + '_StanModel.pkl'
        self._single = single
        if check_outliers and not bool(self._outlier_model.all()):
            idx = self.check_outliers(threshold=3.5)
            if single and len(idx) > 0 and not bool(self._outlier_model[0]):
                warnings.warn(f'Possible outliers were identified at indices {idx}. Check the residuals and consider '
                              f're-running with outliers=True')
        if single:
            self._squeeze()
        return self

    # ------------------------------------------------------------------------------------------------------------
    def _model_type(self):
        """model selection (Inverter._get_stan_model, inversion.py:1576-1610) -> (model_type, series names, parallel names)"""
        ser = [k for k, v in self.distributions.items() if v['dist_type'] == 'series']
        par = [k for k, v in self.distributions.items() if v['dist_type'] == 'parallel']
        if len(ser) == 1 and len(par) == 0:
            return 'Series', ser, par
        if len(ser) == 0 and len(par) == 1:
            return 'Parallel', ser, par
        if len(ser) == 1 and len(par) == 1:
            return 'Series-Parallel', ser, par
        if len(ser) == 1 and len(par) == 2:
            par = sorted(par)  # the reference orders the parallel distributions by name (inversion.py:1963-1968)
            self.distributions[par[0]]['order'], self.distributions[par[1]]['order'] = 1, 2
            return 'Series-2Parallel', ser, par
        raise NotImplementedError("the 'MultiDist' model (arbitrary numbers of distributions) is a placeholder in the "
                                  "reference (its Stan file is not shipped) and is not implemented")

    def prepare(self, frequencies, Z, mode='optimize', nonneg=False, outliers=False, scale_Z=True, sigma_min=0.002,
                inductance_scale=1, outlier_lambda=None, random_seed=1234, chains=2, spectrum_offset=0,
                spectrum_ids=None):
        """Everything ``fit`` does before it calls a solver, as a public hook (bench.py times the solver alone with it):
        sorting, scaling, kernel / penalty matrices, the Stan data of the selected program, Stan-style random initial
        points.  Returns (problem, u0): a ``capi.SeriesProblem`` (``map_lbfgs`` / ``map_newton`` / ``nuts`` /
        ``logpost_grad`` / ``constrain``) and the initial points [B, D] ('optimize') or [B, chains, D] ('sample')."""
        if mode not in ('optimize', 'sample'):
            raise ValueError(f"Invalid mode {mode}. Options are 'optimize', 'sample'")
        model_type, ser, par = self._model_type()
        name = ser[0] if ser else par[0]
        freq, Zb = self._to_batch(frequencies, Z)
        self.f_train, self.Z_train = freq.numpy(), Zb
        Zs = self._scale_Z(Zb, scale_Z, fit_type='map' if mode == 'optimize' else 'bayes')
        B = Zs.shape[0]
        ids = self._global_ids(B, spectrum_offset, spectrum_ids)
        prob = self._build_problem(freq, Zs, model_type, name, par, mode, nonneg, bool(outliers), sigma_min,
                                   inductance_scale, outlier_lambda)
        nch = 1 if mode == 'optimize' else chains
        u0 = self._initial_points(prob, ids, nch, random_seed, 0)
        return prob, (u0.reshape(B, prob.D) if mode == 'optimize' else u0)

    def _global_ids(self, B, spectrum_offset, spectrum_ids):
        if spectrum_ids is None:
            return torch.arange(spectrum_offset, spectrum_offset + B, dtype=torch.int64, device=self.device)
        ids = torch.as_tensor(spectrum_ids, dtype=torch.int64, device=self.device).reshape(-1)
        if ids.shape[0] != B:
            raise ValueError(f'spectrum_ids must have one entry per spectrum ({B})')
        return ids

    def _build_problem(self, freq, Zs, model_type, name, par, mode, nonneg, outliers, sigma_min, inductance_scale,
                       outlier_lambda):
        """The Stan data of one program for one (sub-)batch (Inverter._prep_stan_data, inversion.py:1684-2122)."""
        tau, eps, m = self._grid(freq, name)
        c = _MODE[mode]
        Zst = torch.cat((Zs.real, Zs.imag), dim=1).contiguous()
        common = dict(nonneg=bool(nonneg), sigma_min=sigma_min, ups_alpha=c['ups_alpha'], ups_beta=c['ups_beta'],
                      induc_scale=float(inductance_scale), device=self.device)
        if model_type in ('Series', 'Parallel'):  # same constants (inversion.py:1714-1754)
            L = torch.stack([c['l'][0] * m['L0'], c['l'][1] * m['L1'], c['l'][2] * m['L2']])
            return capi.SeriesProblem(torch.cat((m['A_re'], m['A_im']), dim=-2), Zst, freq, L, outliers=bool(outliers),
                                      parallel=model_type == 'Parallel',
                                      sigma_out_lambda=10.0 if outlier_lambda is None else float(outlier_lambda),
                                      sigma_out_alpha=c['sigma_out_alpha'], sigma_out_beta=1.0, **common)
        _, _, mp = self._grid(freq, par[0])
        csp = _MODE_SP[mode]
        Ls = torch.stack([csp['ls'][j] * m[f'L{j}'] for j in range(3)])
        Lp = torch.stack([csp['lp'][j] * mp[f'L{j}'] for j in range(3)])
        extra = {}
        x_sum_invscale = csp['x_sum_invscale']
        if model_type == 'Series-2Parallel':  # inversion.py:1961-2049
            _, _, mp2 = self._grid(freq, par[1])
            extra = dict(Ap2=torch.cat((mp2['A_re'], mp2['A_im']), dim=-2),
                         Lp2=torch.stack([csp['lp'][j] * mp2[f'L{j}'] for j in range(3)]),
                         xp2_scale=float(self.distributions[par[1]].get('x_scale', 1)))
            x_sum_invscale = 0.1 if mode == 'sample' else 0.0
        return capi.SeriesProblem(torch.cat((m['A_re'], m['A_im']), dim=-2), Zst, freq, Ls,
                                  Ap=torch.cat((mp['A_re'], mp['A_im']), dim=-2), Lp=Lp, x_sum_invscale=x_sum_invscale,
                                  xp_scale=float(self.distributions[par[0]].get('x_scale', 1)), **extra, **common)

    def _initial_points(self, prob, ids, nch, random_seed, attempt):
        """Stan: init='random' -> U(-2, 2) per unconstrained coordinate; rows keyed by the global (spectrum, chain) index
        (and by the attempt number when a point is redrawn), so a spectrum's start does not depend on the sharding."""
        rows = (ids[:, None] * nch + torch.arange(nch, device=self.device)[None, :]).reshape(-1)
        seed = int(random_seed) + 7919 * int(attempt)
        return _hash_uniform(seed, 0, 0, prob.D, self.device, ids=rows).reshape(len(ids), nch, prob.D)

    MAX_INIT_ATTEMPTS = 100  # Stan's own limit [Stan-upstream]

    # ------------------------------------------------------------------------------------------------------------
    def _fit_core(self, freq, Zs, ids, model_type, name, par, mode, nonneg, outliers, sigma_min, inductance_scale,
                  outlier_lambda, random_seed, max_iter, warmup, samples, chains, init, ridge_init, flags, polish,
                  keep_draws):
        """One Stan program on one (sub-)batch: build the problem, initialise, run the solver, read back."""
        prob = self._build_problem(freq, Zs, model_type, name, par, mode, nonneg, outliers, sigma_min,
                                   inductance_scale, outlier_lambda)
        self._problem = prob
        B, D, K, Nf = prob.B, prob.D, prob.K, prob.Nf
        nch = 1 if mode == 'optimize' else chains
        if mode == 'sample' and chains * samples > capi.SUMMARIZE_MAX_DRAWS:
            raise ValueError(f'chains * samples = {chains * samples} exceeds the {capi.SUMMARIZE_MAX_DRAWS} merged draws '
                             f'per spectrum that the on-device percentile kernel holds')

        def ridge_overlay(u0):
            # partial init from the ridge fit (inversion.py:1649-1677); Stan draws the remaining parameters randomly
            if ridge_init is None:
                return u0
            x = ridge_init['x']
            if nonneg:  # exact zeros of the active-set QP cannot be log-transformed: floor them (cvxopt's interior
                x = torch.clamp(x, min=1e-8 * x.abs().max(dim=1, keepdim=True).values)  # iterates sit ~1e-9 above)
                x = torch.log(x)
            u0[:, :, 2:2 + K] = x[:, None, :]
            u0[:, :, 0] = torch.log(ridge_init['Rinf_raw'])[:, None]
            u0[:, :, 1] = torch.log(ridge_init['induc_raw'])[:, None]
            if outliers:
                so = torch.full((B, Nf), 0.1, dtype=torch.float64, device=self.device)
                if flags is not None:
                    so[flags] = 1.0
                u0[:, :, 6 + K:6 + K + Nf] = torch.log(so)[:, None, :]
            return u0

        if init is not None:
            if init.shape[-1] != D or init.numel() not in (B * D, B * nch * D):
                raise ValueError(f'init must be [{B}, {D}] or [{B}, {nch}, {D}] (unconstrained parameters, Stan order)')
            u0 = init.reshape(B, -1, D)
            u0 = (u0.expand(B, nch, D) if u0.shape[1] == 1 else u0).clone()  # one point per spectrum: every chain starts there
        else:
            u0 = ridge_overlay(self._initial_points(prob, ids, nch, random_seed, 0))
        # Stan rejects an initial point whose log density or gradient is not finite and, for random inits, draws again
        # (up to 100 times) before it gives up with an error; a user-supplied point is an error right away
        lp0, g0 = prob.logpost_grad(u0.reshape(B * nch, D), jacobian=mode == 'sample',
                                    spec=torch.arange(B, device=self.device).repeat_interleave(nch).to(torch.int32))
        bad = ~(torch.isfinite(lp0) & torch.isfinite(g0).all(dim=1)).reshape(B, nch)
        attempt = 0
        while bool(bad.any()):
            attempt += 1
            if init is not None or attempt >= self.MAX_INIT_ATTEMPTS:
                rows = torch.nonzero(bad.any(dim=1))[:, 0].tolist()
                raise RuntimeError(f'Initialization failed: log density or gradient not finite at the initial point of '
                                   f'spectra {rows[:10]}{" ..." if len(rows) > 10 else ""}')
            fresh = ridge_overlay(self._initial_points(prob, ids, nch, random_seed, attempt))
            u0 = torch.where(bad[:, :, None], fresh, u0)
            sel = torch.nonzero(bad.reshape(-1))[:, 0]
            lp1, g1 = prob.logpost_grad(u0.reshape(B * nch, D)[sel], jacobian=mode == 'sample',
                                        spec=(sel // nch).to(torch.int32))
            still = ~(torch.isfinite(lp1) & torch.isfinite(g1).all(dim=1))
            bad = torch.zeros(B * nch, dtype=torch.bool, device=self.device).index_put_((sel,), still).reshape(B, nch)
        if mode == 'optimize':
            r = prob.map_lbfgs(u0.reshape(B, D), max_iter=max_iter)
            failed = r['status'] < 0
            if bool(failed.any()):  # Stan raises on a failed line search; a batch keeps its other spectra and warns
                rows = torch.nonzero(failed)[:, 0].tolist()
                warnings.warn(f'L-BFGS terminated with an error (line search failed) for {len(rows)} spectra: '
                              f'{rows[:10]}{" ..." if len(rows) > 10 else ""}; see _opt_result["status"]')
            if polish:
                p = prob.map_newton(r['u'])
                r.update(u=p['u'], lp=p['lp'], gnorm=p['gnorm'], newton_iters=p['iters'])
            point = prob.split_outputs(prob.constrain(r['u']))
            return dict(point=point, opt=r, draws=None, stats=None)
        # When the draws are not kept, the batch goes through the sampler in pieces of eight waves of resident chains, so
        # that the raw and constrained draws (1.4 MB per spectrum at 2 x 200 draws of D = 209) never exceed a few GB:
        # the 1e5-spectrum sweep of config 4 needs 18 GB per GPU otherwise (SURVEY.md section 7 hard part 5).
        sm = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == 'cuda' else 148
        piece = B if keep_draws else (getattr(self, '_hmc_piece', None) or max(1, (8 * 16 * sm) // chains))
        parts = []
        for a in range(0, B, piece):
            b = min(B, a + piece)
            pb = prob if (a == 0 and b == B) else prob.subset(a, b)
            r = pb.nuts(u0[a:b], chains=chains, warmup=warmup, samples=samples, seed=random_seed, spectrum_ids=ids[a:b])
            lost = ~torch.isfinite(r['stepsize'])
            if bool(lost.any()):
                rows = (torch.nonzero(lost.any(dim=1))[:, 0] + a).tolist()
                raise RuntimeError(f'NUTS could not find a step size for chains of spectra {rows[:10]}')
            nb = b - a
            spec = torch.arange(nb, dtype=torch.int32, device=self.device).repeat_interleave(chains * samples)
            cons = pb.constrain(r['draws'].reshape(nb * chains * samples, D), spec=spec).reshape(
                nb, chains * samples, prob.P)
            draws = prob.split_outputs(cons)
            # posterior mean over the merged chains (Inverter._extract_parameter, inversion.py:2514-2519), on device
            pm, _ = capi.summarize(cons, device=self.device)
            stats = {k: r[k] for k in ('stepsize', 'n_leapfrog', 'n_divergent', 'n_maxdepth', 'accept')}
            stats['rhat'], stats['ess_bulk'] = self._diagnostics(draws, chains, samples)
            parts.append((pm, stats, draws if keep_draws else None))
            del r, cons
        pm = torch.cat([p[0] for p in parts])
        stats = {k: torch.cat([p[1][k] for p in parts]) for k in parts[0][1]}
        return dict(point=prob.split_outputs(pm), opt=None, draws=parts[0][2], stats=stats)

    def _diagnostics(self, draws, chains, samples):
        """Split R-hat and bulk ESS of every coefficient of every distribution, R_inf and the inductance (what pystan
        prints after sampling, inversion.py:1218-1221): [B, n_quantities] each, computed on the device
        (bdrt_diagnostics)."""
        keys = [k for k in ('x', 'xs', 'xp', 'xp2') if k in draws and not (k == 'x' and 'xs' in draws)]
        q = torch.cat([draws[k] for k in keys] + [draws['Rinf'][..., None], draws['induc'][..., None]], dim=-1)
        return capi.diagnostics(q.contiguous(), chains, device=self.device)

    def _merge_results(self, results, B, model_type, name, par, mode, sigma_min, keep_draws):
        """Scatter the sub-batch results (one per Stan program) back into batch order and rescale
        (Inverter._extract_parameter / _rescale_coef, inversion.py:2445-2450, :2494-2519)."""
        def merge(get, fill=float('nan')):
            parts = [(ix, get(res)) for ix, _, res in results if get(res) is not None]
            if not parts:
                return None
            if len(results) == 1 and results[0][0] is None:
                return parts[0][1]
            ref = parts[0][1]
            out = torch.full((B,) + tuple(ref.shape[1:]), fill, dtype=ref.dtype, device=ref.device)
            for ix, v in parts:
                out[ix] = v
            return out
        keys = list(results[0][2]['point'].keys())
        point = {k: merge(lambda r, k=k: r['point'].get(k)) for k in set(keys) | {'sigma_out'}}
        s = self._Z_scale
        self.distribution_fits, self.error_fit = {}, {}
        if model_type == 'Series':
            self.distribution_fits[name] = {'coef': point['x'] * s[:, None]}
        elif model_type == 'Parallel':
            self.distribution_fits[name] = {'coef': point['x'] / s[:, None]}
        else:  # series coef * scale, parallel coef / scale
            self.distribution_fits[name] = {'coef': point['xs'] * s[:, None]}
            self.distribution_fits[par[0]] = {'coef': point['xp'] / s[:, None]}
            if model_type == 'Series-2Parallel':
                self.distribution_fits[par[1]] = {'coef': point['xp2'] / s[:, None]}
        self.R_inf = point['Rinf'] * s
        self.inductance = point['induc'] * s
        self.error_fit['sigma_min'] = sigma_min * s
        self.error_fit['sigma_tot'] = point['sigma_tot'] * s[:, None]
        self.error_fit['sigma_res'] = point['sigma_res'] * s
        for k in ('alpha_prop', 'alpha_re', 'alpha_im'):
            self.error_fit[k] = point[k]
        if point.get('sigma_out') is not None:
            self.error_fit['sigma_out'] = point['sigma_out'] * s[:, None]
        if mode == 'optimize':
            self.fit_type = 'map'
            self._sample_result = None
            self._opt_result = {k: merge(lambda r, k=k: r['opt'].get(k), fill=0) for k in results[0][2]['opt']} \
                if len(results) == 1 else {k: merge(lambda r, k=k: r['opt'].get(k), fill=0)
                                           for k in ('lp', 'iters', 'n_eval', 'status')}
        else:
            self.fit_type = 'bayes'
            self._sample_stats = {k: merge(lambda r, k=k: r['stats'][k], fill=0) for k in results[0][2]['stats']}
            if keep_draws:
                dk = set().union(*[set(res['draws'].keys()) for _, _, res in results])
                self._sample_result = {k: merge(lambda r, k=k: r['draws'].get(k)) for k in dk}
            else:
                self._sample_result = None

    def _get_init_from_ridge(self, freq, Zb, nonneg, inductance_scale, ridge_kw):
        """inversion.py:1616-1682: under-fitted ridge solution -> Stan initial values (scaled units)."""
        from .ridge import ridge_fit
        kw = dict(penalty='integral', hyper_lambda=True, lambda_0=1, hl_beta=5, weights='modulus')  # :1642
        kw.update(ridge_kw)
        single = self._single
        self._single = False
        ridge_fit(self, freq, Zb, **kw)
        self._single = single
        name = list(self.distributions.keys())[0]
        s = self._Z_scale
        induc = self.inductance / s
        induc = torch.where(induc <= 0, torch.full_like(induc, 1e-10), induc)  # :1665-1667
        return {'x': self.distribution_fits[name]['coef'] / s[:, None], 'Rinf_raw': self.R_inf / s / 100.0,
                'induc_raw': induc / inductance_scale}

    def _ridge_outlier_flags(self, freq, Zb, threshold, use_existing_fit, **ridge_kw):
        """check_outliers for a ridge fit (inversion.py:3351-3367): no error model yet, so residuals relative to |Z| are
        flagged when they exceed the 75th percentile by ``threshold`` inter-quartile ranges (utils.py:143-146).
        Returns bool [B, Nf]."""
        from .ridge import ridge_fit
        single = self._single
        self._single = False
        if not (use_existing_fit and self.fit_type == 'ridge'):
            ridge_fit(self, freq, Zb, preset='Huang', **ridge_kw)
        Zerr = self.predict_Z(freq.numpy()) - self.Z_train
        self._single = single
        Zmod = self.Z_train.abs()
        er, ei = (Zerr.real / Zmod), (Zerr.imag / Zmod)

        def thresh(y):
            q = torch.quantile(y, torch.tensor([0.25, 0.75], dtype=torch.float64, device=y.device), dim=1)
            return q[1] + threshold * (q[1] - q[0])
        tr, ti = thresh(er.abs()), thresh(ei.abs())
        return er ** 2 + ei ** 2 >= (tr ** 2 + ti ** 2)[:, None]

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
    def _pct(self, draws, percentile):
        """np.percentile(draws, percentile, axis=1) of [B, S, P] (or [B, S]) draws, on device (bdrt_summarize)."""
        d3 = draws if draws.dim() == 3 else draws[:, :, None]
        q = capi.summarize(d3.contiguous(), percentiles=(percentile,), want_mean=False, device=self.device)[1][0]
        return q if draws.dim() == 3 else q[:, 0]

    def _draw_key(self, name):
        """name of a distribution's coefficient block among the draws (Inverter._get_stan_coef_name)"""
        if len(self.distribution_fits) == 1:
            return 'x'
        if self.distributions[name]['dist_type'] == 'series':
            return 'xs'
        return 'xp' if 'xp2' not in self._sample_result else f"xp{self.distributions[name].get('order', 1)}"

    def coef_percentile(self, distribution_name, percentile):
        """inversion.py:2547-2566: per-coefficient np.percentile (linear interpolation) of the merged draws."""
        if self.fit_type != 'bayes' or self._sample_result is None:
            raise ValueError('Percentile prediction is only available for bayes_fit')
        q = self._pct(self._sample_result[self._draw_key(distribution_name)], percentile)
        s = self._Z_scale[:, None]
        q = q / s if self.distributions[distribution_name]['dist_type'] == 'parallel' else q * s
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
            torch.as_tensor(self.coef_percentile(name, percentile), device=self.device).reshape(-1, basis_tau.shape[-1])
        if basis_tau.dim() == 2:  # one basis per spectrum: eval_tau [T] (shared) or [B, T]
            nb = basis_tau.shape[0]
            coef = coef.reshape(nb, -1)
            et = et if et.dim() == 2 else et[None, :].expand(nb, -1)
            phi = torch.exp(-(info['epsilon'] * torch.log(et[:, :, None] / basis_tau[:, None, :])) ** 2)
            return self._ret(self._apply(phi, coef))
        phi = torch.exp(-(info['epsilon'] * torch.log(et[:, None] / basis_tau[None, :])) ** 2)
        return self._ret(coef @ phi.T)

    @staticmethod
    def _apply(A, coef):
        """A @ coef for coefficients [B, K] or draws [B, S, K]; A [Nf, K] (shared grid) or [B, Nf, K]."""
        At = A.transpose(-1, -2)
        if A.dim() == 3 and coef.dim() == 2:
            return torch.bmm(coef[:, None, :], At)[:, 0, :]
        return coef @ At

    def _pred_matrices(self, f, name):
        info = self.distributions[name]
        tau = torch.as_tensor(info['tau'])
        if tau.dim() == 2 and f.dim() == 1:  # per-spectrum bases, one set of evaluation frequencies for all
            f = f[None, :].expand(tau.shape[0], -1).contiguous()
        return capi.build_A(f, tau, info['epsilon'], kernel=info['kernel'],
                            dist_type=info['dist_type'], symmetry=info.get('symmetry') or 'planar',
                            bc=info.get('bc') or 'transmissive', ct=info.get('ct', False), k_ct=info.get('k_ct'),
                            device=self.device)

    def predict_Z(self, frequencies, times=None, distributions=None, include_offsets=True, percentile=None):
        """inversion.py:2669 (generic branch :2942-2959): series distributions add A @ coef, parallel distributions add
        1 / (A @ coef); R_inf and the inductance are added when include_offsets."""
        if times is not None:
            raise NotImplementedError('drift fits are out of scope')
        if distributions is None:
            names = list(self.distribution_fits.keys())
        else:
            names = [distributions] if isinstance(distributions, str) else list(distributions)
        f = torch.as_tensor(np.asarray(frequencies, dtype=np.float64))
        fd = f.to(self.device)
        if percentile is not None:
            if self.fit_type != 'bayes' or self._sample_result is None:
                raise ValueError('Percentile prediction is only available for bayes_fit results')
            if len(self.distributions) != 1 or names != list(self.distribution_fits.keys()) or \
                    self.distributions[names[0]]['dist_type'] != 'series':
                # several distributions, a subset, or a parallel distribution (coefficients rescale as x / scale and enter
                # as Z = 1 / (A x), inversion.py:2712-2725): the impedance of every draw, then the percentile of its real
                # and imaginary parts (inversion.py:2705-2737)
                single = self._single
                self._single = False
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    Zd = self.predict_Z_distribution(frequencies, distributions=names, include_offsets=include_offsets)
                self._single = single
                return self._ret(torch.complex(self._pct(Zd.real.contiguous(), percentile),
                                               self._pct(Zd.imag.contiguous(), percentile)))
            A_re, A_im = self._pred_matrices(f, names[0])
            s = self._Z_scale[:, None, None]
            x = self._sample_result['x'] * s
            Zr = self._apply(A_re, x)
            Zi = self._apply(A_im, x)
            if include_offsets:
                Zr = Zr + (self._sample_result['Rinf'][..., None] * s)
                Zi = Zi + 2 * np.pi * (fd[:, None, :] if fd.dim() == 2 else fd) * \
                    (self._sample_result['induc'][..., None] * s)
            Zp = torch.complex(self._pct(Zr, percentile), self._pct(Zi, percentile))
            return self._ret(Zp)
        Zp = None
        for name in names:
            A_re, A_im = self._pred_matrices(f, name)
            coef = self._coef_batch(name)
            z = torch.complex(self._apply(A_re, coef), self._apply(A_im, coef))
            if self.distributions[name]['dist_type'] == 'parallel':
                z = 1.0 / z
            Zp = z if Zp is None else Zp + z
        if include_offsets:
            Rinf = torch.as_tensor(self.R_inf, dtype=torch.float64, device=self.device).reshape(-1, 1)
            ind = torch.as_tensor(self.inductance, dtype=torch.float64, device=self.device).reshape(-1, 1)
            Zp = Zp + torch.complex(Rinf.expand_as(Zp.real), (2 * np.pi * fd * ind).expand_as(Zp.real))
        return self._ret(Zp)

    def predict_Z_distribution(self, frequencies, distributions=None, include_offsets=True):
        """inversion.py:2963-3031: the impedance of every posterior draw, [B, chains * samples, Nf] complex
        ([chains * samples, Nf] for one spectrum)."""
        if self.fit_type != 'bayes' or self._sample_result is None:
            raise ValueError('predict_Z_distribution is only available for bayes_fit results')
        names = list(self.distribution_fits.keys()) if distributions is None else (
            [distributions] if isinstance(distributions, str) else list(distributions))
        if len(names) != len(self.distributions) or not include_offsets:
            warnings.warn('All distributions and offsets should be included for meaningful results from '
                          'predict_Z_distribution')
        f = torch.as_tensor(np.asarray(frequencies, dtype=np.float64))
        fd = f.to(self.device)
        s = self._Z_scale[:, None, None]
        Zm = None
        for name in names:
            A_re, A_im = self._pred_matrices(f, name)
            par = self.distributions[name]['dist_type'] == 'parallel'
            x = self._sample_result[self._draw_key(name)]
            x = x / s if par else x * s
            z = torch.complex(self._apply(A_re, x), self._apply(A_im, x))
            z = 1.0 / z if par else z
            Zm = z if Zm is None else Zm + z
        if include_offsets:
            Zr = (self._sample_result['Rinf'][..., None] * s).expand_as(Zm.real)
            Zi = 2 * np.pi * (fd[:, None, :] if fd.dim() == 2 else fd) * (self._sample_result['induc'][..., None] * s)
            Zm = Zm + torch.complex(Zr, Zi.expand_as(Zm.real))
        return Zm[0].cpu().numpy() if self._single else Zm

    def ridge_ReImCV(self, frequencies, Z, lambdas=None, **kw):
        """inversion.py:902-944: Re-Im cross-validation of lambda_0.  Returns the optimum (one value per spectrum of a
        batch); the table is in ``cv_result``.  The instance is left fitted at that lambda_0."""
        self.ridge_fit(frequencies, Z, lambda_0='cv', cv_lambdas=lambdas, **kw)
        return float(self._cv_lambda_0[0]) if self._single else self._cv_lambda_0

    def predict_Rp(self, distributions=None, percentile=None, time=None):
        """inversion.py:3033: area under the DRT, sum(coef) sqrt(pi) / epsilon."""
        names = list(self.distribution_fits.keys()) if distributions is None else (
            [distributions] if isinstance(distributions, str) else list(distributions))
        if len(names) > 1 or self.distributions[names[0]]['kernel'] != 'DRT':  # inversion.py:3050-3052
            single = self._single
            self._single = False
            Zr = self.predict_Z(np.array([1e20, 1e-20]), distributions=names, percentile=percentile)
            self._single = single
            rp = (Zr[:, 1] - Zr[:, 0]).real
            return float(rp[0]) if self._single else rp
        name = names[0]
        eps = self.distributions[name]['epsilon']
        if percentile is None:
            rp = self._coef_batch(name).sum(dim=1) * np.pi ** 0.5 / eps
        else:
            if self.fit_type != 'bayes' or self._sample_result is None:
                raise ValueError('Percentile prediction is only available for bayes_fit results')
            arr = (self._sample_result['x'] * self._Z_scale[:, None, None]).sum(dim=2) * np.pi ** 0.5 / eps
            rp = self._pct(arr, percentile)
        return float(rp[0]) if self._single else rp

    def predict_sigma(self, frequencies=None, percentile=None, times=None):
        """inversion.py:3089-3139.  At the training frequencies: the fitted sigma_tot split into real / imaginary parts.
        Elsewhere: rebuilt from the error-model parameters and the predicted impedance (baseline outlier level
        min(sigma_out), as the reference does -- it notes that this does not match sigma_tot exactly)."""
        if times is not None:
            raise NotImplementedError('drift fits are out of scope')
        if self.fit_type not in ('map', 'bayes'):
            raise ValueError('Error scale prediction only available for bayes_fit and map_fit')
        if percentile is not None and (self.fit_type != 'bayes' or self._sample_result is None):
            raise ValueError('Percentile prediction is only available for bayes_fit')
        nf = self.f_train.shape[-1]
        s = self._Z_scale
        if frequencies is None or (
                np.shape(frequencies) == np.shape(self.f_train) and
                np.array_equal(mat.rel_round(np.ravel(frequencies), 10), mat.rel_round(np.ravel(self.f_train), 10))):
            if percentile is not None:
                st = self._pct(self._sample_result['sigma_tot'], percentile) * s[:, None]
            else:
                st = torch.as_tensor(self.error_fit['sigma_tot'], device=self.device).reshape(-1, 2 * nf)
            return self._ret(st[:, :nf]), self._ret(st[:, nf:])
        dev = self.device

        def col(v):  # [B, 1]
            return torch.as_tensor(v, dtype=torch.float64, device=dev).reshape(-1, 1)
        if percentile is not None:
            r = self._sample_result

            def pct_all(d):  # np.percentile over every draw (and entry) of a spectrum -> [B, 1]
                return self._pct(d.reshape(d.shape[0], -1).contiguous(), percentile).reshape(-1, 1)
            sigma_res = pct_all(r['sigma_res']) * s[:, None]
            a_prop, a_re, a_im = pct_all(r['alpha_prop']), pct_all(r['alpha_re']), pct_all(r['alpha_im'])
            so = r.get('sigma_out')
            so_min = torch.zeros_like(sigma_res) if so is None or not bool(torch.isfinite(so).all()) else \
                (self._pct(so, percentile) * s[:, None]).min(dim=1, keepdim=True).values
        else:
            e = self.error_fit
            sigma_res, a_prop, a_re, a_im = col(e['sigma_res']), col(e['alpha_prop']), col(e['alpha_re']), col(e['alpha_im'])
            so = e.get('sigma_out')
            if so is None:
                so_min = torch.zeros_like(sigma_res)
            else:
                so = torch.as_tensor(so, dtype=torch.float64, device=dev).reshape(sigma_res.shape[0], -1)
                so_min = torch.nan_to_num(so, nan=0.0).min(dim=1, keepdim=True).values
        single = self._single
        self._single = False
        Zp = self.predict_Z(frequencies, percentile=percentile)
        self._single = single
        base2 = sigma_res ** 2 + so_min ** 2 + col(self.error_fit['sigma_min']) ** 2
        common = (a_re * Zp.real) ** 2 + (a_im * Zp.imag) ** 2
        sigma_re = torch.sqrt(base2 + (a_prop * Zp.real) ** 2 + common)
        sigma_im = torch.sqrt(base2 + (a_prop * Zp.imag) ** 2 + common)
        return self._ret(sigma_re), self._ret(sigma_im)

    def score(self, frequencies, Z, metric='chi_sq', weights=None, part='both', times=None):
        """inversion.py:3141-3160: weighted chi^2 per frequency or r^2 (utils.r2_score) of the predicted impedance.
        Returns a float for one spectrum, a tensor [B] for a batch."""
        from .ridge import _weights
        if part not in ('both', 'real', 'imag'):
            raise ValueError(f"Invalid part {part}. Options are 'both', 'real', or 'imag'")
        if metric not in ('chi_sq', 'r2'):
            raise ValueError(f"Invalid metric {metric}. Options are 'chi_sq', 'r2'")
        single = self._single
        self._single = False
        Zp = self.predict_Z(frequencies, times=times)
        self._single = single
        Zt = torch.as_tensor(np.asarray(Z) if not torch.is_tensor(Z) else Z).to(torch.complex128).to(self.device)
        Zt = Zt.reshape(Zp.shape)
        w_re, w_im = _weights(Zt, weights)
        if part == 'both':
            zp, zt, w = torch.cat((Zp.real, Zp.imag), 1), torch.cat((Zt.real, Zt.imag), 1), torch.cat((w_re, w_im), 1)
        elif part == 'real':
            zp, zt, w = Zp.real, Zt.real, w_re
        else:
            zp, zt, w = Zp.imag, Zt.imag, w_im
        if metric == 'chi_sq':
            out = (((zp - zt) * w) ** 2).sum(dim=1) / Zp.shape[1]
        else:
            avg = (w * zt).sum(dim=1, keepdim=True) / w.sum(dim=1, keepdim=True)
            out = 1 - (w * (zp - zt) ** 2).sum(dim=1) / (w * (zt - avg) ** 2).sum(dim=1)
        return float(out[0]) if self._single else out

    def check_outliers(self, frequencies=None, Z=None, threshold=3.5, use_existing_fit=True, **ridge_kw):
        """inversion.py:3313-3376.  An existing MAP / HMC fit of the same data: combined z-score of the residuals under
        the fitted error model.  Otherwise (no fit of this data yet, or ``use_existing_fit=False``, or a ridge fit):
        ridge fit (``preset='Huang'`` unless one exists) and the inter-quartile rule on the residuals relative to |Z|.
        ``frequencies`` / ``Z`` default to the training data.  Returns indices ([n] for one spectrum, [n, 2]
        (spectrum, frequency) for a batch)."""
        fit_exists = self.fit_type in ('ridge', 'map', 'bayes') and not self._recalc_mat
        if frequencies is not None and Z is not None and fit_exists:
            f_in = np.asarray(torch.as_tensor(frequencies).cpu(), dtype=np.float64)
            Z_in = torch.as_tensor(np.asarray(Z) if not torch.is_tensor(Z) else Z).to(torch.complex128)
            if f_in.ndim == 1 and f_in.shape == np.shape(self.f_train):  # the training data are stored sorted
                o = np.argsort(-f_in)
                f_in, Z_in = f_in[o], Z_in[..., torch.as_tensor(o.copy())]
            fit_exists = f_in.shape == np.shape(self.f_train) and np.array_equal(f_in, self.f_train) and \
                tuple(Z_in.reshape(-1).shape) == tuple(self.Z_train.reshape(-1).shape) and \
                bool(torch.equal(Z_in.reshape(-1).to(self.device), self.Z_train.reshape(-1)))
        elif frequencies is None or Z is None:
            if self.fit_type not in ('ridge', 'map', 'bayes'):
                raise ValueError('frequencies and Z must be given if the Inverter has not been fitted')
            frequencies, Z, fit_exists = self.f_train, self.Z_train, True
        if not (use_existing_fit and fit_exists) or self.fit_type == 'ridge':
            single = torch.as_tensor(np.asarray(Z) if not torch.is_tensor(Z) else Z).dim() == 1
            freq, Zb = self._to_batch(frequencies, Z)
            flags = self._ridge_outlier_flags(freq, Zb, threshold, use_existing_fit and fit_exists, **ridge_kw)
            self._single = single
            idx = torch.nonzero(flags)
            return idx[:, 1].cpu().numpy() if single else idx
        single = self._single
        self._single = False
        Zp = self.predict_Z(self.f_train)
        self._single = single
        err_re = Zp.real - self.Z_train.real
        err_im = Zp.imag - self.Z_train.imag
        nf = self.f_train.shape[-1]
        st = torch.as_tensor(self.error_fit['sigma_tot'], device=self.device).reshape(-1, 2 * nf)
        zs = torch.sqrt(((err_re / st[:, :nf]) ** 2 + (err_im / st[:, nf:]) ** 2) / 2)
        idx = torch.nonzero(zs > threshold)
        return idx[:, 1].cpu().numpy() if self._single else idx

    # ------------------------------------------------------------------------------------------------------------
    # saving and loading fits (inversion.py:3980-4064): the same attribute lists, tensors stored as numpy arrays so a
    # file can be read without a GPU (and by code written against the reference's pickles)
    # ------------------------------------------------------------------------------------------------------------
    def get_fit_attributes(self, which='all'):
        fit_attributes = {
            'common': {'core': ['distributions', 'distribution_fits', 'f_train', 'Z_train', '_Z_scale', 'fit_type',
                                'R_inf', 'inductance', '_single'],
                       'detail': ['distribution_matrices']},
            'ridge': {'core': [], 'detail': ['_ridge_iters', '_ridge_converged']},
            'map': {'core': ['stan_model_name', 'error_fit'], 'detail': ['_init_params', '_opt_result']},
            'bayes': {'core': ['stan_model_name', '_sample_result', 'error_fit'],
                      'detail': ['_init_params', '_sample_stats']},
        }
        if self.fit_type not in fit_attributes:
            raise ValueError('No fit to save')
        if which == 'all':
            return sum(fit_attributes['common'].values(), []) + sum(fit_attributes[self.fit_type].values(), [])
        if which not in ('core', 'detail'):
            raise ValueError(f"Invalid which argument {which}. Options are 'core', 'detail', 'all'")
        return fit_attributes['common'][which] + fit_attributes[self.fit_type][which]

    def save_fit_data(self, filename=None, which='all'):
        """inversion.py:4004-4036: pickle (or return) the dict of fit attributes; ``which`` in 'core', 'detail', 'all'."""
        import pickle

        def host(v):
            if torch.is_tensor(v):
                return v.detach().cpu().numpy()
            if isinstance(v, dict):
                return {k: host(x) for k, x in v.items()}
            return v
        fit_data = {att: host(getattr(self, att)) for att in self.get_fit_attributes(which) if hasattr(self, att)}
        if filename is None:
            return fit_data
        with open(filename, 'wb') as f:  # stan_models.save_pickle (stan_models.py:6-9)
            pickle.dump(fit_data, f, pickle.HIGHEST_PROTOCOL)

    def load_fit_data(self, data):
        """inversion.py:4038-4064: restore a fit saved by save_fit_data (file name or dict)."""
        import pickle
        if isinstance(data, str):
            with open(data, 'rb') as f:
                data = pickle.load(f)

        def dev(v):
            if isinstance(v, np.ndarray) and v.dtype.kind in 'fc':
                return torch.as_tensor(v, device=self.device)
            if isinstance(v, dict):
                return {k: dev(x) for k, x in v.items()}
            return v
        for k, v in data.items():
            if k in ('_Z_scale', 'Z_train', '_sample_result', '_sample_stats', '_opt_result', 'distribution_matrices'):
                v = dev(v)
            elif k in ('distribution_fits', 'error_fit', 'R_inf', 'inductance') and not data.get('_single', False):
                v = dev(v)
            setattr(self, k, v)
        if 'distributions' in data:
            self._distributions = data['distributions']
        if torch.is_tensor(self._Z_scale) and self._Z_scale.dim() == 0:
            self._Z_scale = self._Z_scale.reshape(1)
        if 'distribution_matrices' not in data:
            self.distribution_matrices = {k: {} for k in self._distributions}  # rebuilt on demand
        return self

    def ridge_fit(self, frequencies, Z, **kw):
        from .ridge import ridge_fit
        return ridge_fit(self, frequencies, Z, **kw)
