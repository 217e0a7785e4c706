"""TEST DOUBLE for the device library, for CPU tests of the Python host code only.

`OracleProblem` has the interface of bayes_drt_b200.capi.SeriesProblem and answers it with the oracle (numpy model,
Stan-semantics L-BFGS and NUTS restatements); `install(monkeypatch)` puts it, oracle kernel matrices and torch
percentiles behind bayes_drt_b200.inverter so that the shipped Inverter runs end to end on CPU tensors.  Nothing in the
package imports this module: the product has no CPU path (bayes_drt_b200/_lib.py raises without the CUDA library / a GPU).
"""
import numpy as np
import torch

from bayes_drt_b200 import capi as _capi
from oracle import lbfgs as olb, matrices as om, model as omod, model_sp as osp, nuts as onuts

_SPLIT_OUTPUTS = _capi.SeriesProblem.split_outputs  # the shipped layout code (taken before install() swaps the class)


class OracleProblem:
    def __init__(self, A, Z, freq, L, nonneg=False, outliers=False, parallel=False, sigma_min=0.002, ups_alpha=0.05,
                 ups_beta=0.1, induc_scale=1.0, sigma_out_lambda=10.0, sigma_out_alpha=2.0, sigma_out_beta=1.0, Ap=None,
                 Lp=None, x_sum_invscale=0.0, xp_scale=1.0, Ap2=None, Lp2=None, xp2_scale=1.0, device=None):
        A, Z, freq, L = (np.asarray(t, dtype=np.float64) for t in (A, Z, freq, L))
        self.B, self.Nf, self.K = Z.shape[0], Z.shape[1] // 2, A.shape[-1]
        self.outliers, self.parallel = bool(outliers), bool(parallel)
        self.series_parallel, self.two_parallel = Ap is not None, Ap2 is not None
        self.Kp = 0 if Ap is None else np.shape(Ap)[-1]
        self.Kp2 = 0 if Ap2 is None else np.shape(Ap2)[-1]
        self.mod = osp if self.series_parallel else omod
        self.ds = []
        for b in range(self.B):
            Ab = A[b] if A.ndim == 3 else A
            fb = freq[b] if freq.ndim == 2 else freq
            if self.series_parallel:
                d = dict(Nf=self.Nf, Ks=self.K, Kp=self.Kp, freq=fb, As=Ab, Ap=np.asarray(Ap, dtype=np.float64), Z=Z[b],
                         Ls=list(L), Lp=list(np.asarray(Lp, dtype=np.float64)), sigma_min=sigma_min, ups_alpha=ups_alpha,
                         ups_beta=ups_beta, induc_scale=induc_scale, x_sum_invscale=x_sum_invscale, xp_scale=xp_scale,
                         pos=bool(nonneg))
                if self.two_parallel:
                    d.update(Kp2=self.Kp2, Ap2=np.asarray(Ap2, dtype=np.float64), Lp2=list(np.asarray(Lp2, dtype=np.float64)),
                             xp2_scale=xp2_scale)
            else:
                d = dict(Nf=self.Nf, K=self.K, freq=fb, A=Ab, Z=Z[b], L0=L[0], L1=L[1], L2=L[2], sigma_min=sigma_min,
                         ups_alpha=ups_alpha, ups_beta=ups_beta, induc_scale=induc_scale, pos=bool(nonneg) or self.parallel,
                         outliers=self.outliers, sigma_out_lambda=sigma_out_lambda, sigma_out_alpha=sigma_out_alpha,
                         sigma_out_beta=sigma_out_beta, parallel=self.parallel)
            self.ds.append(d)
        self.D = self.mod.n_params(self.ds[0])
        self.P = self.K + self.Kp + self.Kp2 + 6 + 2 * self.Nf + (self.Nf if self.outliers else 0)

    def _f(self, d, jacobian=False):
        def f(u):
            with np.errstate(all='ignore'):
                lp, g = self.mod.logpost(u, d, jacobian=jacobian)
            if not np.isfinite(lp) or not np.all(np.isfinite(g)):
                return None
            return -lp, -g
        return f

    def logpost_grad(self, u, spec=None, jacobian=False):
        u = np.asarray(u, dtype=np.float64).reshape(-1, self.D)
        lp, g = np.empty(len(u)), np.empty_like(u)
        for i in range(len(u)):
            d = self.ds[int(spec[i]) if spec is not None else i % self.B]
            with np.errstate(all='ignore'):
                lp[i], g[i] = self.mod.logpost(u[i], d, jacobian=jacobian)
        return torch.tensor(lp), torch.tensor(g)

    def map_lbfgs(self, u0, max_iter=2000, **kw):
        u0 = np.asarray(u0, dtype=np.float64).reshape(self.B, self.D)
        rs = [olb.minimize(self._f(d), u0[b], max_iter=max_iter) for b, d in enumerate(self.ds)]
        return dict(u=torch.tensor(np.stack([r['x'] for r in rs])), lp=torch.tensor([-r['f'] for r in rs]),
                    iters=torch.tensor([r['iters'] for r in rs], dtype=torch.int32),
                    n_eval=torch.tensor([r['n_eval'] for r in rs], dtype=torch.int32),
                    status=torch.tensor([r['code'] for r in rs], dtype=torch.int32))

    def subset(self, a, b):
        """spectra a .. b - 1 as a problem of their own (capi.SeriesProblem.subset)"""
        import copy
        sub = copy.copy(self)
        sub.ds, sub.B = self.ds[a:b], b - a
        return sub

    def nuts(self, u0, chains=2, warmup=200, samples=200, seed=0, spectrum_ids=None, **kw):
        # (random streams keyed by the GLOBAL spectrum index, like the CUDA sampler's Philox key)
        ids = np.arange(self.B) if spectrum_ids is None else np.asarray(spectrum_ids).reshape(-1)
        u0 = np.asarray(u0, dtype=np.float64).reshape(self.B, chains, self.D)
        draws = np.empty((self.B, chains, samples, self.D))
        st = {k: np.zeros((self.B, chains)) for k in ('stepsize', 'n_leapfrog', 'n_divergent', 'n_maxdepth', 'accept')}
        for b, d in enumerate(self.ds):
            def lg(u, d=d):
                with np.errstate(all='ignore'):
                    return self.mod.logpost(u, d, jacobian=True)
            for c in range(chains):
                r = onuts.sample_chain(lg, u0[b, c], warmup=warmup, samples=samples,
                                       seed=seed * 1000 + int(ids[b]) * chains + c)
                draws[b, c] = r['draws']
                for k in st:
                    st[k][b, c] = r.get(k, 0.0)
        out = {k: torch.tensor(v) for k, v in st.items()}
        out['draws'] = torch.tensor(draws)
        return out

    def constrain(self, u, spec=None):
        u = np.asarray(u, dtype=np.float64).reshape(-1, self.D)
        rows = []
        for i in range(u.shape[0]):
            d = self.ds[int(spec[i]) if spec is not None else (i if u.shape[0] == self.B else 0)]
            o = self.mod.constrain(u[i], d)
            x = [o['xs'], o['xp1'] if self.two_parallel else o['xp']] + ([o['xp2']] if self.two_parallel else []) \
                if self.series_parallel else [o['x']]
            tail = [o['sigma_out']] if self.outliers else []
            rows.append(np.concatenate(x + [[o['Rinf'], o['induc'], o['sigma_res'], o['alpha_prop'], o['alpha_re'],
                                             o['alpha_im']], o['sigma_tot']] + tail))
        return torch.tensor(np.stack(rows))

    def split_outputs(self, out):
        return _SPLIT_OUTPUTS(self, out)


def install(monkeypatch):
    """put the oracle behind the seams of bayes_drt_b200.inverter / .ridge (capi entry points, device context)"""
    from bayes_drt_b200 import inverter

    def build_A(freq, tau, eps, kernel='DRT', dist_type='series', symmetry='planar', bc='transmissive', ct=False,
                k_ct=None, device=None):
        kw = dict(epsilon=float(eps), kernel=kernel, dist_type=dist_type, symmetry=symmetry, bc=bc, ct=ct, k_ct=k_ct)
        f, t = np.asarray(freq, dtype=np.float64), np.asarray(tau, dtype=np.float64)
        if f.ndim == 2:
            rows = [(om.construct_A(f[g], 'real', tau=t[g] if t.ndim == 2 else t, **kw),
                     om.construct_A(f[g], 'imag', tau=t[g] if t.ndim == 2 else t, **kw)) for g in range(f.shape[0])]
            return torch.tensor(np.stack([r[0] for r in rows])), torch.tensor(np.stack([r[1] for r in rows]))
        return torch.tensor(om.construct_A(f, 'real', tau=t, **kw)), torch.tensor(om.construct_A(f, 'imag', tau=t, **kw))

    def build_L(freq, tau, eps, order, device=None):
        return torch.tensor(om.construct_L(np.asarray(freq, dtype=np.float64), tau=np.asarray(tau, dtype=np.float64),
                                           epsilon=float(eps), order=order))

    def summarize(draws, percentiles=(), want_mean=True, device=None):
        q = torch.stack([torch.quantile(draws, float(p) / 100.0, dim=1) for p in percentiles]) if len(percentiles) else None
        return (draws.mean(dim=1) if want_mean else None), q
    def diagnostics(draws, chains, device=None):
        d = np.asarray(draws, dtype=np.float64)
        G, S, P = d.shape
        x = d.reshape(G, chains, S // chains, P)
        ess = np.array([[onuts.ess_bulk(x[g, :, :, p]) for p in range(P)] for g in range(G)])
        h = (S // chains) // 2
        z = np.concatenate((x[:, :, :h], x[:, :, h:2 * h]), axis=1)  # split chains [G, 2 chains, h, P]
        W = z.var(axis=2, ddof=1).mean(axis=1)
        rhat = np.sqrt(((h - 1) / h * W + z.mean(axis=2).var(axis=1, ddof=1)) / W)
        return torch.tensor(rhat), torch.tensor(ess)
    monkeypatch.setattr(inverter, 'context', lambda device=None: type('Ctx', (), {'device': torch.device('cpu')})())
    monkeypatch.setattr(inverter.capi, 'build_A', build_A)
    monkeypatch.setattr(inverter.capi, 'build_L', build_L)
    monkeypatch.setattr(inverter.capi, 'summarize', summarize)
    monkeypatch.setattr(inverter.capi, 'diagnostics', diagnostics)
    monkeypatch.setattr(inverter.capi, 'SeriesProblem', OracleProblem)
    return inverter
