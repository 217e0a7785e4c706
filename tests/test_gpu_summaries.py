"""GPU: on-device posterior summaries (bdrt_summarize) against numpy -- the reductions the reference applies to Stan's
draws: np.mean(axis=0) (inversion.py:2514-2519) and np.percentile(axis=0), linear interpolation (:2560, :2702, :3096)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('G,S,P', [(3, 400, 246), (1, 1, 5), (2, 37, 33), (5, 800, 64), (2, 1000, 7), (4, 512, 32), (2, 4000, 9)])
def test_summarize_matches_numpy(G, S, P):
    from bayes_drt_b200 import capi
    rng = np.random.RandomState(G * 1000 + S + P)
    x = rng.standard_normal((G, S, P)) * np.exp(rng.uniform(-3, 3, (1, 1, P)))
    x[0, :, 0] = 1.0  # ties
    pcts = (0, 2.5, 25, 50, 97.5, 99.9, 100)
    mean, quant = capi.summarize(torch.tensor(x), percentiles=pcts)
    assert np.allclose(mean.cpu().numpy(), x.mean(axis=1), rtol=1e-12, atol=1e-14)
    want = np.percentile(x, pcts, axis=1)  # [nq, G, P]
    assert np.array_equal(quant.cpu().numpy(), want) or np.max(np.abs(quant.cpu().numpy() - want)) <= 1e-15 * np.abs(want).max()
    if S > 2:  # a NaN draw poisons that parameter only, like numpy
        xn = x.copy()
        xn[G - 1, 1, P - 1] = np.nan
        mn, qn = capi.summarize(torch.tensor(xn), percentiles=(50,))
        assert torch.isnan(mn[G - 1, P - 1]) and torch.isnan(qn[0, G - 1, P - 1])
        assert torch.isfinite(mn).sum().item() == G * P - 1 and torch.isfinite(qn).sum().item() == G * P - 1
    m2, q2 = capi.summarize(torch.tensor(x), percentiles=(), want_mean=True)
    assert q2 is None and torch.equal(m2, mean)


def test_summarize_errors_and_large_batch():
    from bayes_drt_b200 import capi
    from bayes_drt_b200._lib import BdrtError
    with pytest.raises(BdrtError, match='range'):
        capi.summarize(torch.zeros(1, 4, 3, dtype=torch.float64), percentiles=(101,))
    with pytest.raises(BdrtError, match='shared-memory'):
        capi.summarize(torch.zeros(1, 20000, 3, dtype=torch.float64), percentiles=(50,))
    with pytest.raises(ValueError):
        capi.summarize(torch.zeros(4, 3, dtype=torch.float64))
    # more tiles than resident CTAs; median of an arithmetic progression per column
    G, S, P = 600, 400, 40
    x = torch.arange(S, dtype=torch.float64, device='cuda')[None, :, None] * torch.ones(G, 1, P, dtype=torch.float64, device='cuda')
    x = x[:, torch.randperm(S, device='cuda'), :] + torch.arange(G, dtype=torch.float64, device='cuda')[:, None, None]
    mean, quant = capi.summarize(x.contiguous(), percentiles=(50, 100))
    assert torch.allclose(quant[0], (S - 1) / 2 + torch.arange(G, dtype=torch.float64, device='cuda')[:, None].expand(G, P))
    assert torch.equal(quant[1][:, 0], S - 1 + torch.arange(G, dtype=torch.float64, device='cuda'))
    assert torch.allclose(mean, quant[0])


def test_diagnostics_match_oracle_estimators():
    """bdrt_diagnostics: split R-hat and rank-normalised bulk ESS of every column against oracle/nuts.py: ess_bulk (the
    estimator the golden NUTS files and the benchmark's ESS/s are made with) and the textbook split R-hat, on AR(1)
    chains with different autocorrelations, a stuck chain (ties) and an offset chain (R-hat >> 1)."""
    from bayes_drt_b200 import capi
    from oracle.nuts import ess_bulk
    rng = np.random.RandomState(3)
    G, chains, n, P = 3, 4, 200, 7
    x = np.empty((G, chains, n, P))
    for g in range(G):
        for p in range(P):
            phi = [0.0, 0.3, 0.6, 0.9, 0.95, -0.4, 0.5][p]
            e = rng.standard_normal((chains, n))
            for c in range(chains):
                v = np.empty(n)
                v[0] = e[c, 0]
                for i in range(1, n):
                    v[i] = phi * v[i - 1] + np.sqrt(1 - phi * phi) * e[c, i]
                x[g, c, :, p] = v * (1 + g) + p
    x[1, 2, :, 3] += 4.0           # one chain elsewhere: R-hat well above 1
    x[2, 1, 50:120, 5] = x[2, 1, 50, 5]  # a stuck stretch: tied values
    draws = torch.tensor(x.reshape(G, chains * n, P))
    rhat, ess = capi.diagnostics(draws, chains)
    rhat, ess = rhat.cpu().numpy(), ess.cpu().numpy()
    for g in range(G):
        for p in range(P):
            col = x[g, :, :, p]
            assert ess[g, p] == pytest.approx(ess_bulk(col), rel=1e-9), (g, p)
            h = n // 2
            z = np.concatenate((col[:, :h], col[:, h:2 * h]), axis=0)
            W = z.var(axis=1, ddof=1).mean()
            ref = np.sqrt(((h - 1) / h * W + z.mean(axis=1).var(ddof=1)) / W)
            assert rhat[g, p] == pytest.approx(ref, rel=1e-10), (g, p)
    assert rhat[1, 3] > 1.5 and np.all(np.delete(rhat.ravel(), 1 * P + 3) < 1.3)  # (phi = 0.95: 4 x 200 draws are few)
    # odd chain length: the last draw is dropped, like the split of the oracle
    r2, e2 = capi.diagnostics(torch.tensor(x[:, :, :199].reshape(G, chains * 199, P)), chains)
    assert e2[0, 1].item() == pytest.approx(ess_bulk(x[0, :, :199, 1]), rel=1e-9)
