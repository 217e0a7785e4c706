"""Sampler diagnostics on the device (torch): rank-normalised split-chain bulk ESS (Vehtari et al. 2021), batched."""
import math

import torch


def _autocov(z):
    """z [..., n] -> biased autocovariance [..., n] via FFT"""
    n = z.shape[-1]
    nfft = 1 << int(math.ceil(math.log2(2 * n)))
    zc = z - z.mean(dim=-1, keepdim=True)
    f = torch.fft.rfft(zc, nfft, dim=-1)
    return torch.fft.irfft(f * f.conj(), nfft, dim=-1)[..., :n] / n


def ess_bulk(x):
    """x [..., chains, n] -> ESS [...] (rank-normalised, split chains, Geyer initial monotone sequence)."""
    *lead, c, n = x.shape
    half = n // 2
    z = torch.cat((x[..., :half], x[..., half:2 * half]), dim=-2)  # [..., 2c, half]
    m, n2 = z.shape[-2], z.shape[-1]
    flat = z.reshape(*lead, m * n2)
    ranks = flat.argsort(dim=-1).argsort(dim=-1).to(torch.float64) + 1.0
    p = (ranks - 0.375) / (m * n2 + 0.25)
    z = (math.sqrt(2.0) * torch.erfinv(2 * p - 1)).reshape(*lead, m, n2)
    acov = _autocov(z)
    chain_var = acov[..., 0] * n2 / (n2 - 1.0)
    W = chain_var.mean(dim=-1)
    var_plus = W * (n2 - 1.0) / n2 + z.mean(dim=-1).var(dim=-1, unbiased=True)
    rho = 1.0 - (W[..., None] - acov.mean(dim=-2)) / var_plus[..., None]  # [..., n2]
    rho[..., 0] = 1.0
    npair = n2 // 2
    pairs = rho[..., 0:2 * npair:2] + rho[..., 1:2 * npair:2]  # Gamma_k
    # initial positive sequence: keep pairs until the first negative one; monotone: running minimum
    positive = (pairs >= 0).to(torch.float64).cumprod(dim=-1)
    mono = torch.cummin(pairs, dim=-1).values
    tau = -1.0 + 2.0 * (mono * positive).sum(dim=-1)
    tau = torch.clamp(tau, min=1.0 / math.log10(m * n2))
    return m * n2 / tau
