"""Synthetic benchmark spectra: ZARC / RC spectra on one shared frequency grid.

Recipe of the reference's simulated data (code_EchemActa/Data simulation.ipynb, 'uniform_0.25' noise):
Z = R0 + R1 / (1 + (j 2 pi f tau0)^n),  additive i.i.d. normal noise sigma = 0.0025 * R1 on Z' and Z''.
The generator is a CPU torch.Generator so the CPU baseline and every GPU shard see identical inputs.
"""
import math

import torch


def bench_grid():
    """freq = 10**(5 - arange(70)/10) (Nf = 70), basis_freq = 10**(6 - arange(100)/10) (K = 100)."""
    freq = 10.0 ** (5.0 - torch.arange(70, dtype=torch.float64) / 10.0)
    basis_freq = 10.0 ** (6.0 - torch.arange(100, dtype=torch.float64) / 10.0)
    return freq, basis_freq


def make_spectra(B, freq=None, seed=20240601, noise=0.0025):
    """Returns (freq [Nf], Z [B, Nf] complex128, params dict) on the CPU."""
    if freq is None:
        freq, _ = bench_grid()
    g = torch.Generator(device='cpu').manual_seed(seed)
    R0 = 0.5 + 1.5 * torch.rand(B, generator=g, dtype=torch.float64)
    R1 = 0.5 + 1.5 * torch.rand(B, generator=g, dtype=torch.float64)
    lt = -4.0 + 4.0 * torch.rand(B, generator=g, dtype=torch.float64)
    n = 0.6 + 0.4 * torch.rand(B, generator=g, dtype=torch.float64)
    is_rc = torch.rand(B, generator=g, dtype=torch.float64) < 0.2
    n = torch.where(is_rc, torch.ones_like(n), n)
    tau0 = 10.0 ** lt
    jw = (2j * math.pi) * freq.to(torch.complex128)[None, :] * tau0.to(torch.complex128)[:, None]
    Z = R0[:, None] + R1[:, None] / (1.0 + jw ** n.to(torch.complex128)[:, None])
    sig = (noise * R1)[:, None]
    Z = Z + sig * torch.randn(B, len(freq), generator=g, dtype=torch.float64) \
        + 1j * sig * torch.randn(B, len(freq), generator=g, dtype=torch.float64)
    return freq, Z, dict(R0=R0, R1=R1, tau0=tau0, n=n)


def make_spectra_sp(B, seed=20240605, noise=0.002, nf=81):
    """Config 5 of BASELINE.json: every spectrum on its own frequency grid
    freq_b = 10**(6 - delta_b - arange(81) / 10), delta_b ~ U(0, 1); the cell is a ZARC in series with a finite-length
    (transmissive) Warburg element, the shape the paper fits with a DRT plus a planar transmissive DDT in parallel form
    (code_EchemActa/Run fits.ipynb cell 20).  Returns (freq [B, Nf], Z [B, Nf] complex128) on the CPU."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    rnd = lambda: torch.rand(B, generator=g, dtype=torch.float64)  # noqa: E731
    delta = rnd()
    freq = 10.0 ** (6.0 - delta[:, None] - torch.arange(nf, dtype=torch.float64)[None, :] / 10.0)
    Rinf, R1, Rd = 0.3 + 0.4 * rnd(), 0.6 + 0.8 * rnd(), 0.4 + 0.6 * rnd()
    tau0, td, n = 10.0 ** (-4.0 + 2.0 * rnd()), 10.0 ** (-1.5 + 1.5 * rnd()), 0.7 + 0.25 * rnd()
    jw = (2j * math.pi) * freq.to(torch.complex128)
    x = torch.sqrt(jw * td.to(torch.complex128)[:, None])
    Z = Rinf[:, None] + R1[:, None] / (1.0 + (jw * tau0.to(torch.complex128)[:, None]) ** n.to(torch.complex128)[:, None]) \
        + Rd[:, None] * torch.tanh(x) / x
    Z = Z + noise * (torch.randn(B, nf, generator=g, dtype=torch.float64)
                     + 1j * torch.randn(B, nf, generator=g, dtype=torch.float64))
    return freq, Z


def sp_distributions(x_scale=0.8):
    """The paper's two distributions on one shared basis: any shift of a spectrum's log-uniform grid against it keeps the
    kernel matrices Toeplitz."""
    import numpy as np
    bf = np.logspace(6, -2, 81)
    return {'DRT': {'kernel': 'DRT', 'dist_type': 'series', 'basis_freq': bf},
            'TP-DDT': {'kernel': 'DDT', 'symmetry': 'planar', 'bc': 'transmissive', 'dist_type': 'parallel',
                       'basis_freq': bf, 'x_scale': x_scale}}
