"""Host side of ``Inverter.ridge_fit`` (inversion.py:142-900): argument handling and the assembly of the weighted,
augmented system on the device; the hyper-lambda loop and the QP run in ridge.cu through the C ABI."""
import warnings

import numpy as np
import torch

from . import capi
from . import matrices as mat


def _weights(Zs, weights):
    """Inverter._format_weights for part='both' (inversion.py:2338-2395) -> (w_re, w_im) [B, Nf]."""
    if weights is None or (isinstance(weights, str) and weights == 'unity'):
        w = torch.ones_like(Zs.real)
        return w, w.clone()
    if isinstance(weights, str):
        if weights == 'modulus':
            w = 1.0 / Zs.abs()
            return w, w.clone()
        if weights == 'Orazem':
            w = 1.0 / (Zs.real.abs() + Zs.imag.abs())
            return w, w.clone()
        if weights == 'proportional':
            return 1.0 / Zs.real.abs(), 1.0 / Zs.imag.abs()
        raise ValueError(f"Invalid weights argument {weights}. String options are 'unity', 'modulus', 'proportional', "
                         f"and 'prop_adj'")
    raise NotImplementedError('array / scalar weights are not implemented in this build')


def ridge_fit(inv, frequencies, Z, part='both', penalty='discrete', reg_ord=2, L1_penalty=0, scale_Z=True, nonneg=True,
              weights=None, preset=None, hyper_lambda=True, hl_solution='analytic', hl_beta=2.5, hl_fbeta=None,
              lambda_0=1e-2, cv_lambdas=None, hyper_weights=False, hw_beta=2, hw_wbar=1, xtol=1e-3, max_iter=20,
              hyper_a=False, alpha_a=2, hl_beta_a=2, hyper_b=False, sb=1, correct_phase_offset=False, IERange=None,
              lambda_phz=1, init_phase_offset=False, x0=None, dZ=False, dZ_power=0.5):
    presets = ['Ciucci', 'Huang']
    if preset is not None:
        if preset not in presets:
            raise ValueError('Invalid preset {}. Options are {}'.format(preset, presets))
        if preset == 'Ciucci':
            penalty, lambda_0, hl_fbeta = 'discrete', 'cv', 0.1
        else:  # inversion.py:278-282
            penalty, hl_beta, lambda_0, weights = 'integral', 2.5, 1e-2, 'modulus'
    if penalty in ('discrete', 'cholesky'):
        if hl_beta <= 1:
            raise ValueError("hl_beta must be greater than 1 for penalty 'cholesky' and 'discrete'")
    elif penalty == 'integral':
        if hl_beta <= 2:
            raise ValueError("hl_beta must be greater than 2 for penalty 'integral'")
    else:
        raise ValueError(f"Invalid penalty argument {penalty}. Options are 'integral', 'discrete', and 'cholesky'")
    if hyper_lambda and hyper_weights:
        raise ValueError('hyper_lambda and hyper_weights fits cannot be performed simultaneously')
    if len(inv.distributions) > 1:
        raise ValueError('ridge_fit cannot be used to fit multiple distributions')
    if correct_phase_offset and IERange is None:
        raise ValueError('IERange must be provided if correct_phase_offset==True')
    # options outside the hot path: loud, never a silent fallback
    for flag, nm in ((penalty == 'cholesky', "penalty='cholesky'"), (lambda_0 == 'cv', "lambda_0='cv' (Re-Im CV)"),
                     (hl_fbeta is not None, 'hl_fbeta'), (hl_solution != 'analytic', "hl_solution='lm'"),
                     (hyper_weights, 'hyper_weights'), (hyper_a or hyper_b, 'hyper_a / hyper_b'),
                     (correct_phase_offset, 'correct_phase_offset'), (dZ, 'dZ'), (x0 is not None, 'x0'),
                     (part != 'both', "part != 'both'")):
        if flag:
            raise NotImplementedError(f'ridge_fit option {nm} is not implemented in this build')
    name = list(inv.distributions.keys())[0]
    info = inv.distributions[name]
    if info['dist_type'] != 'series' or info['kernel'] != 'DRT':
        raise NotImplementedError('ridge_fit is implemented for the DRT (series) kernel only in this build')

    dev = inv.device
    freq, Zb = inv._to_batch(frequencies, Z)
    if freq.dim() == 2:
        raise NotImplementedError('ridge_fit takes one frequency grid per batch (per-spectrum grids: Inverter.fit)')
    inv.f_train, inv.Z_train = freq.numpy(), Zb
    Zs = inv._scale_Z(Zb, scale_Z)
    tau, eps, m = inv._grid(freq, name)
    B, Nf = Zs.shape
    K = len(tau)
    n = K + 2
    # augmented matrices: [1 | 0 | A_re], [0 | 2 pi f 1e-4 | A_im]   (inversion.py:401-417)
    A_re = torch.zeros((Nf, n), dtype=torch.float64, device=dev)
    A_im = torch.zeros((Nf, n), dtype=torch.float64, device=dev)
    A_re[:, 2:], A_im[:, 2:] = m['A_re'], m['A_im']
    A_re[:, 0] = 1.0
    if inv.fit_inductance:
        A_im[:, 1] = 2 * np.pi * freq.to(dev) * 1e-4
    w_re, w_im = _weights(Zs, weights)
    shared_w = weights is None or (isinstance(weights, str) and weights == 'unity')
    if shared_w:
        WA_re, WA_im = A_re, A_im
    else:
        WA_re, WA_im = w_re[:, :, None] * A_re[None], w_im[:, :, None] * A_im[None]
    WZ_re, WZ_im = (w_re * Zs.real).contiguous(), (w_im * Zs.imag).contiguous()
    frac = np.zeros(3)
    if isinstance(reg_ord, (int, np.integer)):
        frac[int(reg_ord)] = 1
    else:
        frac[:] = np.asarray(reg_ord, dtype=np.float64)
    bft = torch.as_tensor(1 / (2 * np.pi * tau))
    Pen = torch.zeros((3, n, n), dtype=torch.float64, device=dev)
    Lmat = None
    if penalty == 'integral':
        toep = mat.is_loguniform(bft)
        for o in range(3):
            Pen[o, 2:, 2:] = capi.build_M(bft, eps, o, toep, device=dev)
            m[f'M{o}'] = Pen[o, 2:, 2:]
    else:
        Lmat = torch.zeros((3, K, n), dtype=torch.float64, device=dev)
        for o in range(3):
            Lmat[o, :, 2:] = m[f'L{o}']
            Pen[o] = Lmat[o].T @ Lmat[o]
    if hyper_lambda:
        r = capi.ridge_fit(WA_re, WA_im, WZ_re, WZ_im, Pen, Lmat, penalty=penalty, nonneg=nonneg, max_iter=max_iter,
                           xtol=xtol, hl_beta=float(hl_beta), lambda_0=float(lambda_0), reg_ord=frac,
                           L1_penalty=L1_penalty, epsilon=eps, fit_inductance=inv.fit_inductance, device=dev)
        coef, lam = r['coef'], r['lam']
        inv._ridge_iters, inv._ridge_converged = r['iters'], r['converged']
        if inv._single and not bool(r['converged'][0]):
            warnings.warn(f'Hyperparametric solution did not converge within {max_iter} iterations')
    else:
        # ordinary ridge: one QP with lambda = lambda_0 (inversion.py:835-850)
        G0 = (WA_re.transpose(-1, -2) @ WA_re + WA_im.transpose(-1, -2) @ WA_im)
        P = G0 + sum(frac[o] * lambda_0 * Pen[o] for o in range(3))
        P = P.expand(B, n, n).contiguous()
        L1_vec = torch.full((n,), np.pi ** 0.5 / eps * L1_penalty, dtype=torch.float64, device=dev)
        L1_vec[:2] = 0
        q = -(WA_re.transpose(-1, -2) @ WZ_re[:, :, None])[..., 0] - (WA_im.transpose(-1, -2) @ WZ_im[:, :, None])[..., 0] \
            + L1_vec
        lb = torch.zeros(n, dtype=torch.float64, device=dev)
        if not nonneg:
            lb[2:] = -10.0
        coef, _, _ = capi.qp_bound(P, q.contiguous(), lb, device=dev)
        lam = torch.full((B, 3, n), float(lambda_0), dtype=torch.float64, device=dev)
    # rescale (inversion.py:875-898)
    s = inv._Z_scale
    out = coef * s[:, None]
    out[:, 1] *= 1e-4
    if not inv.fit_inductance:
        out[:, 1] = 0
    inv.distribution_fits = {name: {'coef': out[:, 2:].contiguous(), 'scaled_coef': coef, 'lambda_vectors': lam}}
    inv.R_inf, inv.inductance = out[:, 0].contiguous(), out[:, 1].contiguous()
    inv.error_fit = {}
    inv.fit_type = 'ridge'
    inv._sample_result = None
    if inv._single:
        inv._squeeze()
    return inv
