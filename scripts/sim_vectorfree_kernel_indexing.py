"""CPU check of the index logic of the candidate kernel profiles/r02_candidate_vectorfree_lbfgs.patch: direction_gram()
is transcribed here statement by statement with its lane-distributed 5x5 tables (tab_get / tab_set), ring slots and
age -> slot mapping, driven through a sequence of L-BFGS history updates (including a reset, after which stale table
entries must never be read), and compared with the plain two-loop recursion over the same ring."""
import numpy as np

HH = 5


def two_loop(S, Y, rho, g, nh, head, gamma):
    hidx = lambda j: (head - 1 - j) % HH
    p = -g.copy()
    al = np.zeros(HH)
    for j in range(nh):
        h = hidx(j)
        al[h] = rho[h] * (S[h] @ p)
        p -= al[h] * Y[h]
    p *= gamma
    for j in range(nh - 1, -1, -1):
        h = hidx(j)
        p += (al[h] - rho[h] * (Y[h] @ p)) * S[h]
    return p


def direction_gram(S, Y, rho, g, nh, head, gamma, skyk, yyn, sy, yy):
    """sy, yy: arrays of 32 'lanes' (persist across calls).  Mirrors the CUDA function line by line."""
    hidx = lambda t: (head - 1 - t) + (HH if head - 1 - t < 0 else 0)
    hn = hidx(0)
    sv, yv, gv = S[hn].copy(), Y[hn].copy(), g.copy()
    sg, yg = np.zeros(32), np.zeros(32)
    sg[hn], yg[hn] = sv @ gv, yv @ gv
    sy[hn * HH + hn], yy[hn * HH + hn] = skyk, yyn
    for t in range(1, nh):
        h = hidx(t)
        se, ye = S[h], Y[h]
        sg[h], yg[h] = se @ gv, ye @ gv
        sy[hn * HH + h] = sv @ ye
        sy[h * HH + hn] = se @ yv
        yy[hn * HH + h] = yy[h * HH + hn] = yv @ ye
    a, cc = np.zeros(HH), np.zeros(HH)
    for t in range(HH):
        if t < nh:
            ht = hidx(t)
            sq = sg[ht]
            for u in range(t):
                sq -= a[u] * sy[ht * HH + hidx(u)]
            a[t] = rho[ht] * sq
    for t in range(HH - 1, -1, -1):
        if t < nh:
            ht = hidx(t)
            acc = yg[ht]
            for u in range(HH):
                if u < nh:
                    acc -= a[u] * yy[ht * HH + hidx(u)]
            yr = gamma * acc
            for u in range(t + 1, HH):
                if u < nh:
                    yr += cc[u] * sy[hidx(u) * HH + ht]
            cc[t] = a[t] - rho[ht] * yr
    pr = gamma * a[0] * yv - cc[0] * sv - gamma * gv
    for t in range(1, HH):
        if t < nh:
            h = hidx(t)
            pr += gamma * a[t] * Y[h] - cc[t] * S[h]
    return pr


rng = np.random.RandomState(0)
D = 37
S, Y, rho = np.zeros((HH, D)), np.zeros((HH, D)), np.zeros(HH)
sy, yy = np.full(32, np.nan), np.full(32, np.nan)  # NaN: reading an entry that was never written would show
nh, head = 0, 0
worst = 0.0
for it in range(40):
    reset = it in (0, 17, 18)
    s = rng.randn(D)
    y = s * rng.uniform(0.5, 2.0, D) + 0.1 * rng.randn(D)  # s.y > 0
    S[head], Y[head] = s, y
    skyk, yyn = s @ y, y @ y
    if reset:  # lbfgs.cu: keep the entry just written as the only one, in slot 0
        nh = 0
        if head != 0:
            S[0], Y[0] = S[head].copy(), Y[head].copy()
            head = 0
    gamma = skyk / yyn
    rho[head] = 1.0 / skyk
    head = 0 if head + 1 == HH else head + 1
    nh = min(nh + 1, HH)
    g = rng.randn(D)
    p_ref = two_loop(S, Y, rho, g, nh, head, gamma)
    p_new = direction_gram(S, Y, rho, g, nh, head, gamma, skyk, yyn, sy, yy)
    assert np.all(np.isfinite(p_new)), it
    worst = max(worst, np.abs(p_new - p_ref).max() / np.abs(p_ref).max())
print('iterations 40 (3 resets), history', HH, ': max relative difference to the two-loop recursion', worst)
assert worst < 1e-12
