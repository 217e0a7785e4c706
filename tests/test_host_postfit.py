"""CPU: the host side of the post-fit path -- coefficient rescaling and extraction (Inverter._merge_results), predict_*,
coef_percentile, score, check_outliers -- against the reference's own arithmetic.

tests/golden/postfit.npz (scripts/make_golden_postfit.py) holds what the unmodified reference computes after a fit whose
Stan result is a known synthetic one.  Here the same synthetic result is handed to the host code of
bayes_drt_b200.Inverter, running on CPU tensors: the two device calls it would make are replaced by their oracle
equivalents (kernel matrices from oracle.matrices, percentiles from torch.quantile), everything else is the shipped code."""
import os

import numpy as np
import pytest
import torch

from oracle import matrices as om

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'postfit.npz'))
FREQ, Z, F_PRED, EVAL_TAU = G['freq'], G['Z'], G['f_pred'], G['eval_tau']
BF = np.logspace(6, -2, 81)
TP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel', 'basis_freq': BF}
BP = {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'blocking', 'dist_type': 'parallel', 'basis_freq': BF}
DRT = {'kernel': 'DRT', 'dist_type': 'series', 'basis_freq': BF}
CASES = {
    'series_opt': ({'DRT': {'kernel': 'DRT', 'dist_type': 'series'}}, 'Series', 'optimize', False),
    'series_out_opt': ({'DRT': {'kernel': 'DRT', 'dist_type': 'series'}}, 'Series', 'optimize', True),
    'series_sample': ({'DRT': {'kernel': 'DRT', 'dist_type': 'series'}}, 'Series', 'sample', False),
    'series_out_sample': ({'DRT': {'kernel': 'DRT', 'dist_type': 'series'}}, 'Series', 'sample', True),
    'parallel_opt': ({'TP-DDT': dict(TP)}, 'Parallel', 'optimize', False),
    'parallel_sample': ({'TP-DDT': dict(TP)}, 'Parallel', 'sample', False),
    'sp_opt': ({'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}, 'Series-Parallel', 'optimize', False),
    'sp_sample': ({'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8)}, 'Series-Parallel', 'sample', False),
    's2p_opt': ({'DRT': dict(DRT), 'TP-DDT': dict(TP, x_scale=0.8), 'BP-DDT': dict(BP)}, 'Series-2Parallel', 'optimize',
                False),
}


def _host_inverter(case):
    """bayes_drt_b200.Inverter with the synthetic Stan result of `case` merged in, on CPU tensors."""
    from bayes_drt_b200.inverter import Inverter
    dists, model_type, mode, outliers = CASES[case]
    p = case + '/'
    inv = Inverter.__new__(Inverter)  # no GPU context: only host logic is exercised
    inv.device = torch.device('cpu')
    inv._distributions = {k: dict(v) for k, v in dists.items()}
    for name, info in inv._distributions.items():
        info['tau'], info['epsilon'] = G[p + 'tau/' + name], float(G[p + 'epsilon/' + name])
    ser = [k for k, v in dists.items() if v['dist_type'] == 'series']
    par = sorted(k for k, v in dists.items() if v['dist_type'] == 'parallel')
    if model_type == 'Series-2Parallel':
        inv._distributions[par[0]]['order'], inv._distributions[par[1]]['order'] = 1, 2
    inv.distribution_matrices, inv._recalc_mat, inv._single = {}, False, False
    inv._Z_scale = torch.tensor([float(G[p + 'Z_scale'])], dtype=torch.float64)
    order = np.argsort(FREQ)[::-1]
    inv.f_train = FREQ[order].copy()
    inv.Z_train = torch.tensor(Z[order].copy())[None, :]
    inv._outlier_model = torch.tensor([outliers])

    def pred_matrices(f, name):  # the device call of Inverter._pred_matrices, through the oracle
        info = inv._distributions[name]
        kw = dict(tau=info['tau'], epsilon=info['epsilon'], kernel=info['kernel'], dist_type=info['dist_type'],
                  symmetry=info.get('symmetry', 'planar'), bc=info.get('bc'))
        fn = np.asarray(f, dtype=np.float64)
        return torch.tensor(om.construct_A(fn, 'real', **kw)), torch.tensor(om.construct_A(fn, 'imag', **kw))
    inv._pred_matrices = pred_matrices
    inv._pct = lambda d, q: torch.quantile(d, q / 100.0, dim=1)  # bdrt_summarize == np.percentile (tested on the GPU)
    stan = {k[len(p + 'stan/'):]: torch.tensor(G[k]) for k in G.files if k.startswith(p + 'stan/')}
    rename = {'xp1': 'xp'} if model_type == 'Series-2Parallel' else {}
    if mode == 'optimize':
        point = {rename.get(k, k): v[None] for k, v in stan.items()}
        res = dict(point=point, opt=dict(lp=torch.zeros(1)), draws=None, stats=None)
    else:
        draws = {k: v[None] for k, v in stan.items() if k != 'Z_hat'}
        point = {k: v.mean(dim=1) for k, v in draws.items()}  # Inverter._fit_core: posterior mean over merged chains
        res = dict(point=point, opt=None, draws=draws, stats=dict(accept=torch.zeros(1)))
    name = ser[0] if ser else par[0]
    inv._merge_results([(None, outliers, res)], 1, model_type, name, par, mode, 0.002, True)
    return inv


@pytest.mark.parametrize('case', sorted(CASES))
def test_extraction_and_queries_match_the_reference(case):
    inv = _host_inverter(case)
    p = case + '/'
    mode = CASES[case][2]
    close = lambda a, b, tol=1e-12: np.max(np.abs(np.asarray(a) - b)) <= tol * max(np.max(np.abs(b)), 1e-300)
    for name in inv.distributions:
        assert close(inv.distribution_fits[name]['coef'][0], G[p + 'coef/' + name]), name  # _rescale_coef :2445-2450
        assert close(inv.predict_distribution(name, eval_tau=EVAL_TAU)[0], G[p + 'gamma/' + name], 1e-11), name
        if mode == 'sample':
            assert close(inv.coef_percentile(name, 2.5)[0], G[p + 'coef_p2.5/' + name]), name
            assert close(inv.predict_distribution(name, eval_tau=EVAL_TAU, percentile=90)[0], G[p + 'gamma_p90/' + name], 1e-11)
    assert float(inv.R_inf[0]) == pytest.approx(float(G[p + 'R_inf']), rel=1e-13)
    assert float(inv.inductance[0]) == pytest.approx(float(G[p + 'inductance']), rel=1e-13)
    for k in ('sigma_min', 'sigma_res', 'alpha_prop', 'alpha_re', 'alpha_im', 'sigma_tot', 'sigma_out'):
        if p + 'error_fit/' + k in G.files:
            assert close(torch.as_tensor(inv.error_fit[k]).reshape(-1), G[p + 'error_fit/' + k].reshape(-1)), k
    assert close(inv.predict_Z(inv.f_train)[0], G[p + 'Z_pred_train'], 1e-10)
    assert close(inv.predict_Z(F_PRED)[0], G[p + 'Z_pred'], 1e-10)
    assert close(inv.predict_Z(F_PRED, include_offsets=False)[0], G[p + 'Z_pred_no_offsets'], 1e-10)
    if np.isfinite(G[p + 'Rp']):
        assert float(inv.predict_Rp()[0]) == pytest.approx(float(G[p + 'Rp']), rel=1e-12)
    s_re, s_im = inv.predict_sigma(inv.f_train)
    assert close(torch.cat((s_re[0], s_im[0])), G[p + 'sigma_train'])
    s_re, s_im = inv.predict_sigma(F_PRED)
    assert close(torch.cat((s_re[0], s_im[0])), G[p + 'sigma_pred'], 1e-10)
    assert float(inv.score(FREQ, Z)[0]) == pytest.approx(float(G[p + 'score_chi_sq']), rel=1e-9)
    assert float(inv.score(FREQ, Z, metric='r2', weights='modulus')[0]) == pytest.approx(float(G[p + 'score_r2_modulus']), rel=1e-9)
    idx = inv.check_outliers(threshold=1.0)
    assert np.array_equal(idx[:, 1].numpy(), G[p + 'outlier_idx_z1'])
    if p + 'Z_pred_p25' in G.files:  # single- and multi-distribution fits (:2705-2737)
        assert close(inv.predict_Z(F_PRED, percentile=25)[0], G[p + 'Z_pred_p25'], 1e-10)
    if p + 'sigma_pred_p60' in G.files and p + 'Rp_p75' not in G.files:  # single parallel distribution (:2712-2725)
        s_re, s_im = inv.predict_sigma(F_PRED, percentile=60)
        assert close(torch.cat((s_re[0], s_im[0])), G[p + 'sigma_pred_p60'], 1e-10)
    if p + 'Rp_p75' in G.files:
        assert float(inv.predict_Rp(percentile=75)[0]) == pytest.approx(float(G[p + 'Rp_p75']), rel=1e-12)
        s_re, s_im = inv.predict_sigma(inv.f_train, percentile=60)
        assert close(torch.cat((s_re[0], s_im[0])), G[p + 'sigma_train_p60'])
        s_re, s_im = inv.predict_sigma(F_PRED, percentile=60)
        assert close(torch.cat((s_re[0], s_im[0])), G[p + 'sigma_pred_p60'], 1e-10)
