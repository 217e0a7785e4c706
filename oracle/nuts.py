"""Oracle restatement of Stan 2.19.1's adaptive NUTS sampler (what ``StanModel.sampling`` runs, inversion.py:1218-1221).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

[Stan-upstream] -- pystan==2.19.1.1 (setup.py:19) is absent from the reference tree and from this image.  Restated from
the published algorithm (Betancourt 2017, "A conceptual introduction to HMC", appendix A; Stan reference manual
"HMC algorithm parameters" / "Automatic parameter tuning"):
  * multinomial NUTS with biased progressive sampling between subtrees and uniform (multinomial) sampling inside,
    generalised no-U-turn criterion on rho = sum p and p_sharp = M^-1 p, max_treedepth 10, divergence at H - H0 > 1000;
  * diagonal Euclidean metric; step-size dual averaging (delta, gamma 0.05, t0, kappa 0.75, mu = log(10 eps0));
  * windowed adaptation: init_buffer 75 / base_window 25 (doubling) / term_buffer 50; regularised variance
    (n/(n+5)) var + 1e-3 * 5/(n+5); step-size re-initialised (doubling/halving heuristic) after every metric update.
The reference calls it with warmup=200, iter=warmup+samples, chains=2, control={'adapt_delta': 0.9, 'adapt_t0': 10}.
**Parity unpinned** (no Stan here): the CUDA sampler (csrc/nuts.cu) is compared with this file statistically.

The target is lp(u) = log_prob(jacobian=True); V = -lp.
"""
import numpy as np


class _State:
    __slots__ = ('q', 'p', 'g', 'lp')

    def __init__(self, q, p, g, lp):
        self.q, self.p, self.g, self.lp = q, p, g, lp

    def copy(self):
        return _State(self.q.copy(), self.p.copy(), self.g.copy(), self.lp)


class NUTS:
    def __init__(self, logp_grad, D, rng, max_depth=10, delta=0.9, gamma=0.05, t0=10.0, kappa=0.75):
        self.f = logp_grad  # u -> (lp, grad)
        self.D = D
        self.rng = rng
        self.max_depth = max_depth
        self.delta, self.gamma, self.t0, self.kappa = delta, gamma, t0, kappa
        self.inv_metric = np.ones(D)
        self.eps = 1.0
        self.n_grad = 0
        self.restart_da()

    # ---- Hamiltonian pieces
    def _eval(self, q):
        self.n_grad += 1
        with np.errstate(all='ignore'):
            lp, g = self.f(q)
        return lp, g

    def H(self, z):
        return -z.lp + 0.5 * np.dot(z.p, self.inv_metric * z.p)

    def leapfrog(self, z, eps):
        z.p = z.p + 0.5 * eps * z.g
        z.q = z.q + eps * self.inv_metric * z.p
        z.lp, z.g = self._eval(z.q)
        z.p = z.p + 0.5 * eps * z.g

    def sample_p(self):
        return self.rng.standard_normal(self.D) / np.sqrt(self.inv_metric)

    # ---- step size
    def restart_da(self):
        self.counter, self.s_bar, self.x_bar = 0, 0.0, 0.0
        self.mu = np.log(10 * self.eps)

    def learn_stepsize(self, accept):
        self.counter += 1
        accept = min(1.0, accept)
        eta = 1.0 / (self.counter + self.t0)
        self.s_bar = (1 - eta) * self.s_bar + eta * (self.delta - accept)
        x = self.mu - self.s_bar * np.sqrt(self.counter) / self.gamma
        x_eta = self.counter ** (-self.kappa)
        self.x_bar = (1 - x_eta) * self.x_bar + x_eta * x
        self.eps = np.exp(x)

    def init_stepsize(self, q, lp, g):
        z0 = _State(q.copy(), self.sample_p(), g.copy(), lp)
        z = z0.copy()
        H0 = self.H(z)
        self.leapfrog(z, self.eps)
        h = self.H(z)
        if np.isnan(h):
            h = np.inf
        direction = 1 if (H0 - h) > np.log(0.8) else -1
        while True:
            z = _State(q.copy(), self.sample_p(), g.copy(), lp)
            H0 = self.H(z)
            self.leapfrog(z, self.eps)
            h = self.H(z)
            if np.isnan(h):
                h = np.inf
            dH = H0 - h
            if direction == 1 and not (dH > np.log(0.8)):
                break
            if direction == -1 and not (dH < np.log(0.8)):
                break
            self.eps = 2 * self.eps if direction == 1 else 0.5 * self.eps
            if self.eps > 1e7 or self.eps == 0:
                raise RuntimeError('step size heuristic failed')

    # ---- one transition
    def _criterion(self, ps_minus, ps_plus, rho):
        return np.dot(ps_plus, rho) > 0 and np.dot(ps_minus, rho) > 0

    def _build(self, depth, z, prop, ps_left, ps_right, rho, H0, sign, st):
        """Stan base_nuts::build_tree.  z evolves in place; prop/ps_left/ps_right are 1-element lists (by-ref)."""
        if depth == 0:
            self.leapfrog(z, sign * self.eps)
            st['n_leapfrog'] += 1
            h = self.H(z)
            if np.isnan(h):
                h = np.inf
            if (h - H0) > 1000:
                st['divergent'] = True
            st['lsw'] = np.logaddexp(st['lsw'], H0 - h)
            st['sum_metro'] += 1.0 if H0 - h > 0 else np.exp(H0 - h)
            prop[0] = (z.q.copy(), z.g.copy(), z.lp)
            rho += z.p
            ps_left[0] = self.inv_metric * z.p
            ps_right[0] = ps_left[0]
            return not st['divergent']
        lsw_outer = st['lsw']
        st['lsw'] = -np.inf
        rho_left = np.zeros(self.D)
        dummy = [None]
        if not self._build(depth - 1, z, prop, ps_left, dummy, rho_left, H0, sign, st):
            st['lsw'] = np.logaddexp(lsw_outer, st['lsw'])
            return False
        lsw_left = st['lsw']
        st['lsw'] = -np.inf
        prop_right = [None]
        rho_right = np.zeros(self.D)
        dummy2 = [None]
        if not self._build(depth - 1, z, prop_right, dummy2, ps_right, rho_right, H0, sign, st):
            st['lsw'] = np.logaddexp(lsw_outer, np.logaddexp(lsw_left, st['lsw']))
            return False
        lsw_right = st['lsw']
        lsw_sub = np.logaddexp(lsw_left, lsw_right)
        st['lsw'] = np.logaddexp(lsw_outer, lsw_sub)
        if lsw_right > lsw_sub or self.rng.uniform() < np.exp(lsw_right - lsw_sub):
            prop[0] = prop_right[0]
        rho_sub = rho_left + rho_right
        rho += rho_sub
        return self._criterion(ps_left[0], ps_right[0], rho_sub)

    def transition(self, q, lp, g):
        z = _State(q.copy(), self.sample_p(), g.copy(), lp)
        z_plus, z_minus = z.copy(), z.copy()
        sample = (q.copy(), g.copy(), lp)
        ps_plus = [self.inv_metric * z.p]
        ps_minus = [ps_plus[0].copy()]
        rho = z.p.copy()
        lsw = 0.0
        H0 = self.H(z)
        st = dict(n_leapfrog=0, divergent=False, sum_metro=0.0, lsw=-np.inf)
        depth = 0
        while depth < self.max_depth:
            rho_sub = np.zeros(self.D)
            st['lsw'] = -np.inf
            prop = [None]
            dummy = [None]
            if self.rng.uniform() > 0.5:
                valid = self._build(depth, z_plus, prop, dummy, ps_plus, rho_sub, H0, 1, st)
            else:
                valid = self._build(depth, z_minus, prop, dummy, ps_minus, rho_sub, H0, -1, st)
            if not valid:
                break
            depth += 1
            lsw_sub = st['lsw']
            if lsw_sub > lsw or self.rng.uniform() < np.exp(lsw_sub - lsw):
                sample = prop[0]
            lsw = np.logaddexp(lsw, lsw_sub)
            rho += rho_sub
            if not self._criterion(ps_minus[0], ps_plus[0], rho):
                break
        accept = st['sum_metro'] / st['n_leapfrog']
        return sample, dict(accept=accept, depth=depth, n_leapfrog=st['n_leapfrog'], divergent=st['divergent'])


class _Windows:
    """Stan windowed_adaptation + var_adaptation bookkeeping."""

    def __init__(self, num_warmup, D, init_buffer=75, term_buffer=50, base_window=25):
        self.num_warmup = num_warmup
        if num_warmup < 20:
            self.enabled = False
            return
        self.enabled = True
        if init_buffer + base_window + term_buffer > num_warmup:
            init_buffer = int(0.15 * num_warmup)
            term_buffer = int(0.1 * num_warmup)
            base_window = num_warmup - (init_buffer + term_buffer)
        self.init_buffer, self.term_buffer = init_buffer, term_buffer
        self.window_size = base_window
        self.counter = 0
        self.next_window = init_buffer + base_window - 1
        self.n, self.mean, self.m2 = 0, np.zeros(D), np.zeros(D)

    def learn_variance(self, q):
        """returns the new inverse metric when a window closes, else None"""
        if not self.enabled:
            return None
        c = self.counter
        in_window = (c >= self.init_buffer) and (c < self.num_warmup - self.term_buffer) and (c != self.num_warmup)
        if in_window:
            self.n += 1
            d = q - self.mean
            self.mean = self.mean + d / self.n
            self.m2 = self.m2 + (q - self.mean) * d
        out = None
        if c == self.next_window and c != self.num_warmup:
            self._next()
            n = self.n
            var = self.m2 / (n - 1.0)
            out = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0))
            self.n, self.mean, self.m2 = 0, np.zeros_like(self.mean), np.zeros_like(self.m2)
        self.counter += 1
        return out

    def _next(self):
        last = self.num_warmup - self.term_buffer - 1
        if self.next_window == last:
            return
        self.window_size *= 2
        self.next_window = self.counter + self.window_size
        if self.next_window == last:
            return
        if self.next_window + 2 * self.window_size >= self.num_warmup - self.term_buffer:
            self.next_window = last


def sample_chain(logp_grad, u0, warmup=200, samples=200, seed=0, delta=0.9, t0=10.0, max_depth=10):
    """One adaptive NUTS chain.  Returns dict(draws [samples, D], stepsize, n_leapfrog, n_divergent, n_maxdepth, ...)."""
    rng = np.random.default_rng(seed)
    D = len(u0)
    s = NUTS(logp_grad, D, rng, max_depth=max_depth, delta=delta, t0=t0)
    q = np.array(u0, dtype=np.float64)
    lp, g = s._eval(q)
    if not np.isfinite(lp):
        raise RuntimeError('non-finite log density at the initial point')
    s.init_stepsize(q, lp, g)
    s.restart_da()
    win = _Windows(warmup, D)
    draws = np.empty((samples, D))
    stats = dict(n_leapfrog=0, n_divergent=0, n_maxdepth=0, accept=0.0, warm_leapfrog=0)
    for it in range(warmup + samples):
        (q, g, lp), info = s.transition(q, lp, g)
        if it < warmup:
            stats['warm_leapfrog'] += info['n_leapfrog']
            s.learn_stepsize(info['accept'])
            new_metric = win.learn_variance(q)
            if new_metric is not None:
                s.inv_metric = new_metric
                s.init_stepsize(q, lp, g)
                s.restart_da()
            if it == warmup - 1:
                s.eps = np.exp(s.x_bar)
        else:
            draws[it - warmup] = q
            stats['n_leapfrog'] += info['n_leapfrog']
            stats['n_divergent'] += int(info['divergent'])
            stats['n_maxdepth'] += int(info['depth'] >= max_depth)
            stats['accept'] += info['accept'] / samples
    stats.update(draws=draws, stepsize=s.eps, inv_metric=s.inv_metric, n_grad=s.n_grad)
    return stats


def ess_bulk(x):
    """Rank-normalised split-chain bulk ESS of draws x [chains, n] (Vehtari et al. 2021)."""
    from scipy.stats import norm, rankdata
    c, n = x.shape
    half = n // 2
    z = np.concatenate((x[:, :half], x[:, half:2 * half]), axis=0)
    r = rankdata(z.ravel()).reshape(z.shape)
    z = norm.ppf((r - 0.375) / (z.size + 0.25))
    return _ess(z)


def _ess(z):
    m, n = z.shape
    zc = z - z.mean(axis=1, keepdims=True)
    nfft = 1 << int(np.ceil(np.log2(2 * n)))
    f = np.fft.rfft(zc, nfft, axis=1)
    acov = np.fft.irfft(f * np.conj(f), nfft, axis=1)[:, :n] / n
    chain_var = acov[:, 0] * n / (n - 1.0)
    W = chain_var.mean()
    var_plus = W * (n - 1.0) / n
    if m > 1:
        var_plus += z.mean(axis=1).var(ddof=1)
    rho = 1.0 - (W - acov.mean(axis=0)) / var_plus
    rho[0] = 1.0
    # Geyer initial monotone sequence
    tau = -1.0
    t = 0
    prev = np.inf
    while t + 1 < n:
        pair = rho[t] + rho[t + 1]
        if pair < 0:
            break
        pair = min(pair, prev)
        prev = pair
        tau += 2 * pair
        t += 2
    tau = max(tau, 1.0 / np.log10(m * n))
    return m * n / tau


def ess_indicator(x, thr):
    """Split-chain ESS of the indicator I(x <= thr) (the ESS behind a quantile estimate, Vehtari et al. 2021 sec. 4.3)."""
    c, n = x.shape
    half = n // 2
    z = np.concatenate((x[:, :half], x[:, half:2 * half]), axis=0)
    ind = (z <= thr).astype(np.float64)
    if ind.min() == ind.max():
        return float(z.size)
    return _ess(ind)


def mcse_quantile(x, p):
    """Monte-Carlo standard error of the p-quantile of draws x [chains, n] (posterior::mcse_quantile): the ESS of the
    indicator gives a Beta interval for the probability, mapped back through the empirical quantile function."""
    from scipy.stats import beta
    flat = np.sort(x.ravel())
    S = flat.size
    ess = ess_indicator(x, np.quantile(flat, p))
    a, b = beta.ppf([0.1586553, 0.8413447], ess * p + 1, ess * (1 - p) + 1)
    lo = flat[max(int(np.floor(a * S)), 1) - 1]
    hi = flat[min(int(np.ceil(b * S)), S) - 1]
    return 0.5 * (hi - lo)


def mcse_mean(x):
    """sd / sqrt(bulk ESS) with the un-normalised split-chain ESS of the draws themselves."""
    c, n = x.shape
    half = n // 2
    z = np.concatenate((x[:, :half], x[:, half:2 * half]), axis=0)
    return x.std(ddof=1) / np.sqrt(_ess(z))
