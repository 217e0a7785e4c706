"""Host side of ``Inverter.ridge_fit`` (inversion.py:142-900): argument handling and the assembly of the weighted,
augmented system on the device; the hyper-lambda loop and the QP run in ridge.cu through the C ABI."""
import warnings

import numpy as np
import torch

from . import capi
from . import matrices as mat


def _weights(Zs, weights):
    """Inverter._format_weights (inversion.py:2338-2395) on a batch -> (w_re, w_im) [B, Nf]: named schemes, a real or
    complex constant, or an array ([Nf] or [B, Nf]; real: both parts, complex: real part weights Z', imaginary Z'')."""
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
        if weights == 'prop_adj':  # 25th percentile of |Z|^2, as the reference has it
            p25 = torch.quantile(Zs.abs() ** 2, 0.25, dim=1, keepdim=True)
            return 1.0 / (Zs.real.abs() + p25), 1.0 / (Zs.imag.abs() + p25)
        raise ValueError(f"Invalid weights argument {weights}. String options are 'unity', 'modulus', 'proportional', "
                         f"and 'prop_adj'")
    one = torch.ones_like(Zs.real)
    if isinstance(weights, (float, int)):
        return one * float(weights), one * float(weights)
    if isinstance(weights, complex):
        return one * weights.real, one * weights.imag
    w = torch.as_tensor(np.asarray(weights) if not torch.is_tensor(weights) else weights).to(Zs.device)
    if w.shape[-1] != Zs.shape[1] or w.dim() > 2 or (w.dim() == 2 and w.shape[0] != Zs.shape[0]):
        raise ValueError('Weights array must match length of data')  # inversion.py:2373-2374
    if w.is_complex():
        return (one * w.real.to(torch.float64)).contiguous(), (one * w.imag.to(torch.float64)).contiguous()
    w = (one * w.to(torch.float64)).contiguous()
    return w, w.clone()


def ridge_fit(inv, frequencies, Z, part='both', penalty='discrete', reg_ord=2, L1_penalty=0, scale_Z=True, nonneg=True,
              weights=None, preset=None, hyper_lambda=True, hl_solution='analytic', hl_beta=2.5, hl_fbeta=None,
              lambda_0=1e-2, cv_lambdas=None, hyper_weights=False, hw_beta=2, hw_wbar=1, xtol=1e-3, max_iter=20,
              hyper_a=False, alpha_a=2, hl_beta_a=2, hyper_b=False, sb=1, correct_phase_offset=False, IERange=None,
              lambda_phz=1, init_phase_offset=False, x0=None, dZ=False, dZ_power=0.5, stop_rule='unchanged'):
    """Inverter.ridge_fit (inversion.py:142-900) for a batch.  ``frequencies`` may be [Nf] (one grid) or [B, Nf] (one grid
    per spectrum: kernel matrices built per row on the GPU; the rows' bases share one spacing, so the penalty matrices
    are common).  ``stop_rule``: how the hyper-lambda stop test treats a coefficient that is exactly 0 in two consecutive
    iterations -- 'unchanged' (default): no change; 'nan': numpy's 0/0 = NaN never passes, which is what the
    reference's code does with an exact QP solver (it then always runs max_iter iterations; cvxopt's interior iterates
    are never exactly 0, SURVEY.md section 7 hard part 1b)."""
    if stop_rule not in ('unchanged', 'nan'):
        raise ValueError(f"Invalid stop_rule {stop_rule}. Options are 'unchanged', 'nan'")
    presets = ['Ciucci', 'Huang']
    if preset is not None:
        if preset not in presets:
            raise ValueError('Invalid preset {}. Options are {}'.format(preset, presets))
        if preset == 'Ciucci':  # inversion.py:274-277
            penalty, lambda_0, hl_fbeta = 'discrete', 'cv', 0.1
        else:  # inversion.py:278-282
            penalty, hl_beta, lambda_0, weights = 'integral', 2.5, 1e-2, 'modulus'
    if not isinstance(hl_beta, (int, float, np.floating, np.integer)):
        raise NotImplementedError('one hl_beta per derivative order is not implemented in this build (pass a scalar)')
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
    if part not in ('both', 'real', 'imag'):
        raise ValueError(f"Invalid part {part}. Options are 'both', 'real', 'imag'")
    cv = isinstance(lambda_0, str) and lambda_0 == 'cv'
    # options outside the hot path: loud, never a silent fallback
    for flag, nm in ((hl_fbeta is not None and (penalty == 'integral' or not hyper_lambda),
                      "hl_fbeta without the discrete hyper-lambda penalty"),
                     (hl_solution != 'analytic', "hl_solution='lm'"),
                     (hyper_weights, 'hyper_weights'), (hyper_a or hyper_b, 'hyper_a / hyper_b'),
                     (correct_phase_offset, 'correct_phase_offset'), (dZ, 'dZ'), (x0 is not None, 'x0')):
        if flag:
            raise NotImplementedError(f'ridge_fit option {nm} is not implemented in this build')
    name = list(inv.distributions.keys())[0]
    info = inv.distributions[name]
    if info['dist_type'] != 'series' or info['kernel'] != 'DRT':
        raise NotImplementedError('ridge_fit is implemented for the DRT (series) kernel only in this build')

    dev = inv.device
    freq, Zb = inv._to_batch(frequencies, Z)
    per_grid = freq.dim() == 2
    inv.f_train, inv.Z_train = freq.numpy(), Zb
    Zs = inv._scale_Z(Zb, scale_Z)
    tau, eps, m = inv._grid(freq, name)
    B, Nf = Zs.shape
    K = tau.shape[-1]
    n = K + 2
    # augmented matrices: [1 | 0 | A_re], [0 | 2 pi f 1e-4 | A_im]   (inversion.py:401-417); one pair per spectrum when
    # every spectrum has its own grid
    lead = (B,) if per_grid else ()
    A_re = torch.zeros(lead + (Nf, n), dtype=torch.float64, device=dev)
    A_im = torch.zeros(lead + (Nf, n), dtype=torch.float64, device=dev)
    A_re[..., 2:], A_im[..., 2:] = m['A_re'], m['A_im']
    A_re[..., 0] = 1.0
    if inv.fit_inductance:
        A_im[..., 1] = 2 * np.pi * freq.to(dev) * 1e-4
    w_re, w_im = _weights(Zs, weights)
    shared_w = (weights is None or (isinstance(weights, str) and weights == 'unity')) and not per_grid
    if shared_w:
        WA_re, WA_im = A_re, A_im
    else:
        WA_re, WA_im = w_re[:, :, None] * (A_re if per_grid else A_re[None]), \
            w_im[:, :, None] * (A_im if per_grid else A_im[None])
    WZ_re, WZ_im = (w_re * Zs.real).contiguous(), (w_im * Zs.imag).contiguous()
    frac = np.zeros(3)
    if isinstance(reg_ord, (int, np.integer)):
        frac[int(reg_ord)] = 1
    else:
        frac[:] = np.asarray(reg_ord, dtype=np.float64)
    # penalty matrices depend on ratios of the basis time constants only: the first row stands for all
    bft = torch.as_tensor(1 / (2 * np.pi * (tau[0] if tau.ndim == 2 else tau)))
    Pen = torch.zeros((3, n, n), dtype=torch.float64, device=dev)
    Lmat = None
    if penalty in ('integral', 'cholesky'):
        toep = mat.is_loguniform(bft)
        for o in range(3):
            Pen[o, 2:, 2:] = capi.build_M(bft, eps, o, toep, device=dev)
            m[f'M{o}'] = Pen[o, 2:, 2:]
        if penalty == 'cholesky':  # inversion.py:2309-2321: M in the objective, L = chol(M) (upper) in the lambda rule
            Lmat = torch.zeros((3, K, n), dtype=torch.float64, device=dev)
            for o in range(3):
                Lmat[o, :, 2:] = torch.linalg.cholesky(Pen[o, 2:, 2:], upper=True)
    else:
        Lmat = torch.zeros((3, K, n), dtype=torch.float64, device=dev)
        for o in range(3):
            Lmat[o, :, 2:] = m[f'L{o}']
            Pen[o] = Lmat[o].T @ Lmat[o]

    def core(lam0, part_, sel=None):
        """One ridge fit of the (sub-)batch ``sel`` on one part of the data -> scaled coefficients [b, n], lambda
        vectors, hyper-iterations, convergence flags.  A part that is left out enters with zero rows (_convex_opt,
        inversion.py:1047-1052); the parameter it alone determines is then set by least squares (:855-873)."""
        pick = (lambda t: t) if sel is None else (lambda t: t[sel])
        war, wai = (WA_re, WA_im) if shared_w else (pick(WA_re), pick(WA_im))
        wzr, wzi, zs = pick(WZ_re), pick(WZ_im), pick(Zs)
        if part_ == 'real':
            wai, wzi = torch.zeros_like(wai), torch.zeros_like(wzi)
        elif part_ == 'imag':
            war, wzr = torch.zeros_like(war), torch.zeros_like(wzr)
        b = zs.shape[0]
        if hyper_lambda:
            r = capi.ridge_fit(war, wai, wzr, wzi, Pen, Lmat, nonneg=nonneg, max_iter=max_iter,
                               penalty='integral' if penalty == 'integral' else 'discrete',  # 'cholesky': discrete rule
                               xtol=xtol, hl_beta=float(hl_beta), lambda_0=float(lam0), reg_ord=frac,
                               L1_penalty=L1_penalty, epsilon=eps,
                               fit_inductance=inv.fit_inductance and part_ != 'real', hl_fbeta=hl_fbeta,
                               stop_rule=1 if stop_rule == 'unchanged' else 0, device=dev)
            coef, lam, iters, conv, nfac = r['coef'], r['lam'], r['iters'], r['converged'], r.get('n_factor')
        else:
            # ordinary ridge: one QP with lambda = lambda_0 (inversion.py:835-850)
            G0 = (war.transpose(-1, -2) @ war + wai.transpose(-1, -2) @ wai)
            P = G0 + sum(frac[o] * float(lam0) * Pen[o] for o in range(3))
            P = P.expand(b, n, n).contiguous()
            L1_vec = torch.full((n,), np.pi ** 0.5 / eps * L1_penalty, dtype=torch.float64, device=dev)
            L1_vec[:2] = 0
            q = -(war.transpose(-1, -2) @ wzr[:, :, None])[..., 0] - (wai.transpose(-1, -2) @ wzi[:, :, None])[..., 0] \
                + L1_vec
            lb = torch.zeros(n, dtype=torch.float64, device=dev)
            if not nonneg:
                lb[2:] = -10.0
            coef, _, _ = capi.qp_bound(P, q.contiguous(), lb, device=dev)
            lam = torch.full((b, 3, n), float(lam0), dtype=torch.float64, device=dev)
            iters = torch.ones(b, dtype=torch.int32, device=dev)
            conv = torch.ones(b, dtype=torch.int32, device=dev)
            nfac = None
        are, aim = (A_re, A_im) if not per_grid else (pick(A_re), pick(A_im))
        if part_ == 'imag':  # R_inf from the real part: the least-squares fit of a constant is the mean
            coef[:, 0] = (zs.real - inv._apply(are[..., 2:], coef[:, 2:])).mean(dim=1)
        elif part_ == 'real' and inv.fit_inductance:  # inductance from the imaginary part
            a_l = aim[..., 1]
            coef[:, 1] = ((zs.imag - inv._apply(aim[..., 2:], coef[:, 2:])) * a_l).sum(dim=-1) / (a_l * a_l).sum(dim=-1)
        if nfac is None:
            nfac = torch.zeros(b, dtype=torch.int32, device=dev)
        return coef, lam, iters, conv, nfac

    if cv:
        # Re-Im cross-validation of lambda_0 (Inverter.ridge_ReImCV, inversion.py:902-944): fit the real part and score
        # it on the imaginary part and vice versa, for every lambda_0 of the grid; each spectrum takes its own optimum.
        lambdas = np.logspace(-10, 5, 31) if cv_lambdas is None else np.asarray(cv_lambdas, dtype=np.float64)
        recv = torch.zeros((B, len(lambdas)), dtype=torch.float64, device=dev)
        imcv = torch.zeros_like(recv)
        for i, lam_i in enumerate(lambdas):
            c = core(lam_i, 'real')[0]
            imcv[:, i] = ((Zs.imag - inv._apply(A_im, c)) ** 2).sum(dim=1)
            c = core(lam_i, 'imag')[0]
            recv[:, i] = ((Zs.real - inv._apply(A_re, c)) ** 2).sum(dim=1)
        s2 = (inv._Z_scale ** 2)[:, None]  # the reference scores in unscaled units
        recv, imcv = recv * s2, imcv * s2
        totcv = recv + imcv
        best = torch.argmin(torch.nan_to_num(totcv, nan=float('inf')), dim=1).cpu().numpy()
        lam_b = lambdas[best]
        if inv._single:
            if best[0] in (int(np.argmin(lambdas)), int(np.argmax(lambdas))):
                warnings.warn('Optimal lambda_0 {} determined by Re-Im CV is at the boundary of the evaluated range. '
                              'Re-run with an expanded lambda_0 range to obtain an accurate estimate of the optimal '
                              'lambda_0.'.format(lam_b[0]))
            import pandas as pd
            inv.cv_result = pd.DataFrame(np.stack([lambdas, recv[0].cpu().numpy(), imcv[0].cpu().numpy(),
                                                   totcv[0].cpu().numpy()]).T,
                                         columns=['lambda', 'recv', 'imcv', 'totcv'])
        else:
            inv.cv_result = {'lambda': lambdas, 'recv': recv, 'imcv': imcv, 'totcv': totcv}
        inv._cv_lambda_0 = lam_b
        coef = torch.empty((B, n), dtype=torch.float64, device=dev)
        lam = torch.empty((B, 3, n), dtype=torch.float64, device=dev)
        iters = torch.empty(B, dtype=torch.int32, device=dev)
        conv = torch.empty(B, dtype=torch.int32, device=dev)
        nfac = torch.empty(B, dtype=torch.int32, device=dev)
        for lv in np.unique(lam_b):  # one launch per selected lambda_0
            sel = torch.as_tensor(np.nonzero(lam_b == lv)[0], device=dev)
            coef[sel], lam[sel], iters[sel], conv[sel], nfac[sel] = core(lv, part, sel)
    else:
        coef, lam, iters, conv, nfac = core(lambda_0, part)
    inv._ridge_factorisations = nfac  # Cholesky factorisations of the QP solver per spectrum, counted on the device
    if hyper_lambda:
        inv._ridge_iters, inv._ridge_converged = iters, conv
        if inv._single and not bool(conv[0]):
            warnings.warn(f'Hyperparametric solution did not converge within {max_iter} iterations')
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
