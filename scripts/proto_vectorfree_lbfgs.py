"""CPU prototype: L-BFGS direction from the Gram matrix of {s_i, y_i, g} (vector-free two-loop) against the two-loop
recursion, first on random vectors (algebra), then inside full Stan-semantics L-BFGS runs of the oracle model
(iteration counts / final objective).  Result of the run recorded in DESIGN.md section 8."""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import lbfgs as ol, model as om
from helpers import load_spectrum

def direction_gram(S, Y, RHO, SY, YY, g, gamma):
    m = len(S)
    Sg = np.array([s @ g for s in S]); Yg = np.array([y @ g for y in Y])
    a = np.zeros(m)
    for i in range(m - 1, -1, -1):
        sq = Sg[i] - sum(a[j] * SY[i][j] for j in range(i + 1, m))
        a[i] = RHO[i] * sq
    c = np.zeros(m)
    for i in range(m):
        yr = gamma * (Yg[i] - sum(a[j] * YY[i][j] for j in range(m))) + sum(c[j] * SY[j][i] for j in range(i))
        c[i] = a[i] - RHO[i] * yr
    p = -gamma * g
    for j in range(m):
        p = p + gamma * a[j] * Y[j] - c[j] * S[j]
    return p

def direction_two_loop(S, Y, RHO, g, gamma):
    pk = -g; al = [0.0]*len(S)
    for i in range(len(S)-1, -1, -1):
        al[i] = RHO[i]*(S[i]@pk); pk = pk - al[i]*Y[i]
    pk = pk*gamma
    for i in range(len(S)):
        beta = RHO[i]*(Y[i]@pk); pk = pk + (al[i]-beta)*S[i]
    return pk

# random check of the algebra
rng = np.random.RandomState(0)
D, m = 50, 5
S = [rng.randn(D) for _ in range(m)]; Y = [s + 0.3*rng.randn(D) for s in S]
RHO = [1/(s@y) for s, y in zip(S, Y)]
SY = [[s@y for y in Y] for s in S]; YY = [[y1@y2 for y2 in Y] for y1 in Y]
g = rng.randn(D)
p1 = direction_two_loop(S, Y, RHO, g, 0.7); p2 = direction_gram(S, Y, RHO, SY, YY, g, 0.7)
print('algebra check: max rel diff', np.abs(p1-p2).max()/np.abs(p1).max())

# ---- full L-BFGS runs: two-loop vs Gram-matrix direction (direct dots) vs recurrences
from helpers import oracle_batch
from oracle import lbfgs as olb

def minimize_variant(func, x0, variant, max_iter=50000, history=5, init_alpha=1e-3):
    EPS = 2.220446049250313e-16
    counter = [1]
    xk = np.array(x0, dtype=np.float64); fk, gk = func(x0)
    pk = -gk; S, Y, RHO, SY, YY = [], [], [], [], []
    gamma = 1.0; alpha = init_alpha; it = 0; code = 0
    Sg_old = Yg_old = None
    while code == 0:
        it += 1; reset = (it == 1)
        while True:
            if reset: pk = -gk
            if it > 1 and not reset:
                alpha = min(1.0, 1.01 * olb.cubic_interp(gk_1 @ pk_1, alpha, fk - fk_1, gk @ pk_1, 1e-12, 1.0))
            else: alpha = init_alpha
            rc, out = olb.wolfe_line_search(func, alpha, pk, xk, fk, gk, counter=counter)
            if rc:
                if reset: return dict(x=xk, f=fk, iters=it, n_eval=counter[0], code=-1)
                reset = True; continue
            break
        alpha, xn, fn, gn = out
        xk_1, fk_1, gk_1, pk_1 = xk, fk, gk, pk
        xk, fk, gk = xn, fn, gn
        sk, yk = xk - xk_1, gk - gk_1
        grad_norm, step_norm = np.linalg.norm(gk), np.linalg.norm(sk)
        skyk = yk @ sk
        if reset:
            b0 = (yk @ yk) / skyk; S, Y, RHO, SY, YY = [], [], [], [], []
            pk_1 = pk_1 / b0; alpha = alpha * b0
        gamma = skyk / (yk @ yk)
        # Gram update (direct dots)
        m = len(S)
        for i in range(m):
            SY[i].append(S[i] @ yk); YY[i].append(Y[i] @ yk)
        SY.append([sk @ Y[j] for j in range(m)] + [skyk]); YY.append([yk @ Y[j] for j in range(m)] + [yk @ yk])
        S.append(sk); Y.append(yk); RHO.append(1.0 / skyk)
        if len(S) > history:
            S.pop(0); Y.pop(0); RHO.pop(0); SY.pop(0); YY.pop(0)
            for r in SY: r.pop(0)
            for r in YY: r.pop(0)
        pk = direction_two_loop(S, Y, RHO, gk, gamma) if variant == 'two_loop' else direction_gram(S, Y, RHO, SY, YY, gk, gamma)
        df = abs(fk_1 - fk)
        if df < 1e-12: code = 1
        elif df < 1e4 * max(abs(fk_1), abs(fk), 1.0) * EPS: code = 2
        elif grad_norm < 1e-8: code = 3
        elif -(gk @ pk) / max(abs(fk), 1.0) < 1e7 * EPS: code = 4
        elif step_norm < 1e-8: code = 5
        elif it >= max_iter: code = 6
    return dict(x=xk, f=fk, iters=it, n_eval=counter[0], code=code)

def _func(d):
    def f(u):
        lp, g = om.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(g)): return None
        return -lp, -g
    return f

names = ['ZARC_uniform_0.25', '2ZARC_uniform_0.25', 'ZARC-RL_uniform_0.25']
freq = load_spectrum(names[0])[0]
ds = oracle_batch(freq, [load_spectrum(n)[1] for n in names], mode='optimize')
rng = np.random.RandomState(1)
for b, d in enumerate(ds):
    for trial in range(2):
        u0 = rng.uniform(-2, 2, d['D'] if 'D' in d else 2 * d['K'] + 9)
        with np.errstate(all='ignore'):
            t = time.time(); r1 = minimize_variant(_func(d), u0, 'two_loop'); t1 = time.time() - t
            t = time.time(); r2 = minimize_variant(_func(d), u0, 'gram'); t2 = time.time() - t
        print(names[b], trial, 'two_loop: iters %d f %.6f code %d (%.1fs) | gram: iters %d f %.6f code %d (%.1fs)' % (r1['iters'], r1['f'], r1['code'], t1, r2['iters'], r2['f'], r2['code'], t2), flush=True)
