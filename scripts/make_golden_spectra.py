"""Copy a handful of the reference's simulated spectra (inputs) and the paper's published MAP / HMC outputs
(loose goldens) into tests/golden/spectra.npz so GPU-box tests need no reference tree.

    python scripts/make_golden_spectra.py      (build container only; reads /root/reference)

Sources: data/simulated/Z_<name>.csv (columns Freq,Zreal,Zimag), code_EchemActa/map_results/Gout_<name>.csv
(tau,gamma), code_EchemActa/bayes_results/Gout_<name>.csv (tau,gamma,gamma_lo,gamma_hi).
"""
import os

import numpy as np
import pandas as pd

REF = '/root/reference'
names = ['ZARC_uniform_0.25', 'ZARC-RL_uniform_0.25', '2ZARC_uniform_0.25', 'RC-ZARC_Macdonald_0.25',
         'ZARC_noiseless', 'Gerischer_uniform_0.25']
out = {}
for n in names:
    df = pd.read_csv(f'{REF}/data/simulated/Z_{n}.csv')
    out[f'{n}/freq'] = df['Freq'].values
    out[f'{n}/Z'] = df['Zreal'].values + 1j * df['Zimag'].values
    for kind in ('map', 'bayes'):
        p = f'{REF}/code_EchemActa/{kind}_results/Gout_{n}.csv'
        if os.path.exists(p):
            g = pd.read_csv(p)
            out[f'{n}/{kind}_tau'] = g['tau'].values
            for c in g.columns[1:]:
                out[f'{n}/{kind}_{c}'] = g[c].values
    print(n, len(df), [k for k in out if k.startswith(n + '/')])
dst = os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'spectra.npz')
np.savez_compressed(dst, **out)
print('wrote', dst, os.path.getsize(dst) / 1e3, 'kB')
