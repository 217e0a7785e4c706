"""Golden fixture from REAL Stan output: the MAP fits the reference's authors saved for their paper.

``code_EchemActa/map_results/obj_<name>.pkl`` (written by the reference's ``save_fit_data``, inversion.py) holds, for the
paper's data sets, what ``StanModel.optimizing`` (pystan 2.19, Stan's L-BFGS) returned -- every parameter and every
transformed parameter of the Stan program (``_opt_result``) -- next to the kernel / penalty matrices the fit used
(``distribution_matrices``), the frequency grid, the scaling factor and the distribution settings.  pystan is not
installable here, so these files are the only numbers in the tree that Stan itself computed.  This script reads them
in place (build container only) together with the spectra (``data/simulated/Z_<name>.csv``, three files of
``data/experimental/``) and writes
``tests/golden/stan_map.npz``:

  <name>/model            Stan program (file name the reference pickled)
  <name>/freq, Z          the fitted spectrum (descending frequency), complex
  <name>/meta             JSON: sigma_min, Z_scale and, per distribution in the reference's order: name, kernel, dist_type,
                          symmetry, bc, ct, x_scale, epsilon
  <name>/tau/<dist>       basis time constants
  <name>/proj/<dist>/<M>  the reference's stored matrices M = A_re, A_im, L0, L1, L2 as products with fixed seeded
                          random vectors (M r and l M; the matrices themselves are 50 kB each)
  <name>/stan/<param>     Stan's result: parameters and the transformed parameters the tests compare (Z_hat, sigma_tot,
                          q*, ups*, dups*, xp, x_sum_raw)

The Stan programs of that version of the package differ from the current ones only in declaring ``induc`` directly
instead of ``induc_raw * induc_scale`` (diff of stan_model_files/*_modelcode.txt against
code_EchemActa/bayes-drt_20201113/stan_model_files/), i.e. not at all for the default ``induc_scale = 1``.

    python scripts/make_golden_stan_map.py
"""
import glob
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference'
PROJ_SEED = 20201113


class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, st):
        self.__dict__.update(st if isinstance(st, dict) else {'_state': st})


class _Unpickler(pickle.Unpickler):  # the pickled object is an instance of the authors' own class (module `drt`)
    def find_class(self, module, name):
        try:
            return super().find_class(module, name)
        except Exception:
            return type(name, (_Stub,), {'__module__': module})


def load_obj(path):
    with open(path, 'rb') as fh:
        return _Unpickler(fh).load().__dict__


def proj_vectors(n_rows, n_cols):
    rng = np.random.RandomState(PROJ_SEED + 1000 * n_rows + n_cols)
    return rng.standard_normal(n_rows), rng.standard_normal(n_cols)


EXPERIMENTAL = {'PDAC': 'PDAC_COM3_02109_Contact10_2065C_500C.txt', 'LIB_data': 'DRTtools_LIB_data.txt',
                'LIB_data_qtr': 'DRTtools_LIB_data_qtr.csv',
                # (not PDAC_DRT-TpDDT: on that measured, not exactly log-uniform grid the stored DDT matrix is up to 1.2 % off
                # today's matrices.construct_A at low frequency x long tau -- that version integrated the general path
                # differently; today's code is the specification, tests/golden/matrices.npz)
                'LIB_data_DRT-TpDDT': 'DRTtools_LIB_data.txt', 'LIB_data_qtr_DRT-TpDDT': 'DRTtools_LIB_data_qtr.csv'}


def read_experimental(path):
    """Freq, Z of the three experimental files the paper fits (a Gamry table, a 3-column text file, a csv)."""
    import pandas as pd
    if path.endswith('.csv'):
        df = pd.read_csv(path)
        return df['Freq'].values, df['Zreal'].values + 1j * df['Zimag'].values
    with open(path, encoding='latin-1') as fh:
        lines = fh.read().splitlines()
    hdr = [i for i, ln in enumerate(lines) if 'Freq' in ln.split('\t') and 'Zreal' in ln.split('\t')]
    if hdr:  # Gamry: header row, units row, data rows (leading tab)
        cols = lines[hdr[0]].split('\t')
        iF, iR, iI = cols.index('Freq'), cols.index('Zreal'), cols.index('Zimag')
        rows = [ln.split('\t') for ln in lines[hdr[0] + 2:] if ln.strip()]
        a = np.array([[float(r[iF]), float(r[iR]), float(r[iI])] for r in rows])
    else:
        a = np.array([[float(v) for v in ln.split()] for ln in lines if ln.strip()])
    return a[:, 0], a[:, 1] + 1j * a[:, 2]


KEEP = ('Rinf_raw', 'induc', 'induc_raw', 'x', 'xs', 'xp_raw', 'xp1_raw', 'xp2_raw', 'sigma_res_raw', 'alpha_prop_raw',
        'alpha_re_raw', 'alpha_im_raw', 'ups_raw', 'ups_s_raw', 'ups_p_raw', 'ups_p1_raw', 'ups_p2_raw', 'd0_strength',
        'd1_strength', 'd2_strength', 'd0s_strength', 'd1s_strength', 'd2s_strength', 'd0p_strength', 'd1p_strength',
        'd2p_strength', 'd0p1_strength', 'd1p1_strength', 'd2p1_strength', 'd0p2_strength', 'd1p2_strength',
        'd2p2_strength', 'Rinf', 'q', 'qs', 'qp', 'qp1', 'qp2', 'xp', 'xp1', 'xp2', 'x_sum_raw', 'Z_hat', 'sigma_tot',
        'ups', 'ups_s', 'ups_p', 'ups_p1', 'ups_p2', 'dups', 'dups_s', 'dups_p', 'dups_p1', 'dups_p2', 'sigma_out_raw',
        'sigma_out_scale', 'sigma_out')


def main():
    import pandas as pd
    out = {}
    names = []
    for path in sorted(glob.glob(os.path.join(REF, 'code_EchemActa/map_results/obj_*.pkl'))):
        name = os.path.basename(path)[4:-4]
        d = load_obj(path)
        if 'outliers' in d['stan_model_name']:
            # the outlier programs of that version differ from the current ones (one sigma_out per stacked element, no
            # sigma_out_scale: diff of Series_outliers_modelcode.txt), so their results are not results of today's program
            continue
        csv = os.path.join(REF, 'data/simulated', f'Z_{name}.csv')
        if os.path.exists(csv):
            df = pd.read_csv(csv)
            f_all = df['Freq'].values
            Z_all = df['Zreal'].values + 1j * df['Zimag'].values
        elif name in EXPERIMENTAL:
            f_all, Z_all = read_experimental(os.path.join(REF, 'data/experimental', EXPERIMENTAL[name]))
        else:
            continue
        ft = np.asarray(d['f_train'], dtype=np.float64)
        idx = [int(np.argmin(np.abs(np.log(f_all) - np.log(f)))) for f in ft]
        if not np.allclose(f_all[idx], ft, rtol=1e-9):
            print('skip (training grid not in the data file):', name)
            continue
        p = name + '/'
        out[p + 'model'] = np.array(d['stan_model_name'])
        out[p + 'freq'] = ft
        out[p + 'Z'] = Z_all[idx]
        meta = dict(sigma_min=float(d['sigma_min']), Z_scale=float(d['_Z_scale']), dists=[])
        for dn, info in d['_distributions'].items():
            meta['dists'].append(dict(name=dn, kernel=info['kernel'], dist_type=info['dist_type'],
                                      symmetry=info.get('symmetry'), bc=info.get('bc'), ct=bool(info.get('ct', False)),
                                      x_scale=float(info.get('x_scale', 1.0)), epsilon=float(info['epsilon'])))
            out[p + 'tau/' + dn] = np.asarray(info['tau'], dtype=np.float64)
            for mn, M in d['distribution_matrices'][dn].items():
                M = np.asarray(M, dtype=np.float64)
                l, r = proj_vectors(*M.shape)
                out[p + f'proj/{dn}/{mn}/r'] = M @ r
                out[p + f'proj/{dn}/{mn}/l'] = l @ M
        out[p + 'meta'] = np.array(json.dumps(meta))
        for k, v in d['_opt_result'].items():
            if k in KEEP:
                out[p + 'stan/' + k] = np.asarray(v, dtype=np.float64)
        names.append(name)
    out['names'] = np.array(names)
    dst = os.path.join(ROOT, 'tests', 'golden', 'stan_map.npz')
    np.savez_compressed(dst, **out)
    print(f'wrote {dst}: {len(names)} fits, {os.path.getsize(dst) / 1024:.0f} kB')


if __name__ == '__main__':
    main()
