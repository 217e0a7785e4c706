#!/usr/bin/env python
"""Benchmark of the bayes-drt inversion hot path on B200 (BASELINE.json metric: spectra/sec (MAP), ESS/sec (HMC)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-hmc] [--quick]

Workload (config 4 of BASELINE.json): synthetic ZARC / RC spectra sharing one frequency grid
(Nf = 70, K = 100 basis functions -> D = 209 parameters), reference defaults of Inverter.fit(mode='optimize'): model
'Series', Stan-semantics L-BFGS (history 5, Stan's tolerances, iter cap 50000) from Stan-style random inits U(-2, 2).
Weak scaling: every GPU gets --batch spectra per step (default 12500 = the per-GPU share of the 1e5-spectrum sweep at
8 GPUs, strided shards); one "step" is one full MAP inversion of that batch.  One JSON line is printed by rank 0:
`value` (inputs resident), `e2e` (Inverter.fit with host buffers), `roofline` (lbfgs_kernel against the measured FP64
DMMA peak), `cpu_baseline`, `hmc` (2 chains, one wave, with its own CPU baseline) and `hmc_4chains` (three waves),
`flows` (ridge-initialised / auto-outlier fits, truncated L-BFGS + Newton), `config5` (per-spectrum grids,
Series-Parallel_pos, MAP and HMC), `extras` (ridge fits/s with factorisations counted on device, kernel-matrix builds).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'spectra/sec (MAP)'
UNIT = 'spectra/s'
F_GRAD = 84000.0  # algorithmic flop per log-posterior+gradient, B shape, banded-L accounting


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=3)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    p.add_argument('--batch', type=int, default=12500, help='spectra per GPU per step')
    p.add_argument('--max-iter', type=int, default=50000)
    p.add_argument('--no-hmc', action='store_true')
    p.add_argument('--hmc-batch', type=int, default=1184)
    p.add_argument('--hmc4-batch', type=int, default=1776, help='spectra of the 4-chain HMC leg (3 waves of 2368 chains)')
    p.add_argument('--config5-batch', type=int, default=2368)
    p.add_argument('--config5-hmc-batch', type=int, default=592)
    p.add_argument('--quick', action='store_true', help='headline legs only (MAP value / e2e, 2-chain HMC, extras)')
    p.add_argument('--cpu-sample', type=int, default=0, help='spectra in the CPU-baseline sample (0: 2 per core)')
    return p.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference path (numpy model + Stan-semantics L-BFGS), all host cores.
# pystan / cvxopt are not installable here (no network), so this is kind="port"; see DESIGN.md.
# ----------------------------------------------------------------------------------------------------------------------
def _cpu_one(args):
    freq, Z, bf, seed, max_iter = args
    from oracle import lbfgs as olb, model as omod
    d = omod.prep_series(freq, Z, basis_freq=bf, mode='optimize')
    D = omod.n_params(d)
    u0 = np.random.RandomState(seed).uniform(-2, 2, D)

    def f(u):
        with np.errstate(all='ignore'):
            lp, g = omod.logpost(u, d)
        if not np.isfinite(lp) or not np.all(np.isfinite(g)):
            return None
        return -lp, -g
    r = olb.minimize(f, u0, max_iter=max_iter)
    return r['n_eval']


def cpu_map_throughput(n_spectra, max_iter, cores=None):
    from multiprocessing import Pool
    from bayes_drt_b200 import synth
    cores = cores or os.cpu_count()
    freq, Z, _ = synth.make_spectra(n_spectra, seed=20240601)
    _, bf = synth.bench_grid()
    jobs = [(freq.numpy(), Z[i].numpy(), bf.numpy(), 1234 + i, max_iter) for i in range(n_spectra)]
    t = time.time()
    with Pool(cores) as p:
        nev = p.map(_cpu_one, jobs)
    dt = time.time() - t
    return n_spectra / dt, cores, dt, float(np.mean(nev))


def run_reference(a):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count()
    n = a.cpu_sample or 2 * cores
    times = []
    for _ in range(a.warmup if a.warmup < 1 else 1):  # one untimed pass is enough to warm the process pool / BLAS
        cpu_map_throughput(min(n, cores), a.max_iter, cores)
    vals = []
    for _ in range(a.steps):
        v, c, dt, nev = cpu_map_throughput(n, a.max_iter, cores)
        vals.append(v)
        times.append(dt)
    value = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
        'warmup': a.warmup, 'ms_per_step': 1e3 * float(np.mean(times)), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'config4 sample: ZARC/RC spectra, Nf=70, K=100, Series MAP (oracle port of the Stan '
                               'model + Stan-semantics L-BFGS, numpy), %d spectra per step on %d host cores'
                               % (n, cores)},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d spectra per step, one process per core' % n},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {'sm_mhz': float(np.median(busy)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def _cpu_hmc_one(args):
    """one chain of the oracle's NUTS restatement (2 x (200 + 200) per inversion, like the GPU leg)"""
    freq, Z, bf, seed, warmup, samples = args
    from oracle import model as omod, nuts
    d = omod.prep_series(freq, Z, basis_freq=bf, mode='sample')
    D = omod.n_params(d)
    u0 = np.random.RandomState(seed).uniform(-2, 2, D)

    def lg(u):
        with np.errstate(all='ignore'):
            return omod.logpost(u, d, jacobian=True)
    r = nuts.sample_chain(lg, u0, warmup=warmup, samples=samples, seed=seed)
    K = d['K']
    return r['draws'][:, 2:2 + K], r['n_grad']


def cpu_hmc_throughput(n_spectra, chains=2, warmup=200, samples=200, cores=None):
    """ESS/s of the CPU arm: oracle NUTS, one chain per core, min-bulk-ESS over the K coefficients per inversion (the
    same estimator as the GPU leg), summed and divided by the wall time."""
    from multiprocessing import Pool
    from bayes_drt_b200 import synth
    from oracle.nuts import ess_bulk
    cores = cores or os.cpu_count()
    freq, Z, _ = synth.make_spectra(n_spectra, seed=20240601)
    _, bf = synth.bench_grid()
    jobs = [(freq.numpy(), Z[i].numpy(), bf.numpy(), 5000 + 10 * i + c, warmup, samples)
            for i in range(n_spectra) for c in range(chains)]
    t = time.time()
    with Pool(min(cores, len(jobs))) as p:
        res = p.map(_cpu_hmc_one, jobs, chunksize=1)
    dt = time.time() - t
    tot_ess, ngrad = 0.0, 0
    for i in range(n_spectra):
        x = np.stack([res[i * chains + c][0] for c in range(chains)])  # [chains, samples, K]
        tot_ess += min(ess_bulk(x[:, :, k]) for k in range(x.shape[2]))
        ngrad += sum(res[i * chains + c][1] for c in range(chains))
    return tot_ess / dt, n_spectra / dt, min(cores, len(jobs)), dt, ngrad


def run_ours(a):
    import torch.distributed as dist
    from bayes_drt_b200 import Inverter, capi, synth
    from bayes_drt_b200.distributed import gather_results, shard_indices
    from bayes_drt_b200._lib import context
    ws = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if ws > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        """device time of fn() in seconds: events on torch's current stream (the stream libbdrt launches on), max over
        ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        out = fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() * 1e-3, out

    Bg = a.batch
    # weak scaling: ws * Bg spectra, identical on every rank; rank r takes the strided shard r, r + ws, ... (SURVEY 8e:
    # interleaving balances the load); starts and random streams are keyed by the global index
    ids = shard_indices(ws * Bg, rank, ws, mode='strided')
    freq, Z_all, _ = synth.make_spectra(ws * Bg, seed=20240601)
    _, bf = synth.bench_grid()
    Z_host = Z_all[ids].contiguous().pin_memory()
    ids_dev = ids.to(dev)
    ctx = context(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # -------------------------------------------------------------------------------- kernel-only leg ("value")
    # inputs resident in HBM: the public preparation hook hands out the problem and the Stan-style random starts that
    # Inverter.fit uses; the timed step is the solver (bdrt_map_lbfgs), the read-out of the constrained parameters
    # (bdrt_constrain) and the one gather of the results
    inv = Inverter(basis_freq=bf.numpy(), device=dev)
    prob, u0 = inv.prepare(freq, Z_host.to(dev), mode='optimize', spectrum_ids=ids_dev)

    def step_resident():
        r = prob.map_lbfgs(u0, max_iter=a.max_iter)
        out = prob.constrain(r['u'])
        res = torch.cat((out[:, :prob.K + 6], r['lp'][:, None]), dim=1)
        return r, gather_results(res, indices=ids_dev) if ws > 1 else res

    for _ in range(a.warmup):
        step_resident()
    barrier()
    launches0 = ctx.launches
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, n_eval_tot, n_iter_tot = 0.0, 0, 0
    barrier()
    e0.record()
    for _ in range(a.steps):
        flush.zero_()  # L2 flush between timed iterations (inside the timed region: ~0.05 ms of a multi-second step)
        k0.record()
        r, res = step_resident()
        k1.record()
        k1.synchronize()
        kern_ms += k0.elapsed_time(k1)
        n_eval_tot += int(r['n_eval'].sum().item())
        n_iter_tot += int(r['iters'].sum().item())
    e1.record()
    barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    t_ms = t_ms.item()
    launches = ctx.launches - launches0
    clocks = clk.stop() if rank == 0 else None
    value = ws * Bg * a.steps / (t_ms * 1e-3)
    status = r['status']
    term = {int(k): int((status == k).sum().item()) for k in status.unique().tolist()}
    lp_random = r['lp'].clone()

    # roofline of the dominant kernel (lbfgs_kernel): algorithmic flops / device time of the step on this rank
    dfma, dmma = capi.peak_fp64(dev)
    flops = n_eval_tot * F_GRAD + n_iter_tot * 8.0 * 5 * prob.D  # gradients + two-loop recursion (4 m D MACs)
    achieved = flops / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of the dominant kernel: only a capture of THIS launch shape counts (scripts/gpu_r2_final.sh runs ncu on
    # the bench command and scripts/ncu_traffic.py writes profiles/summary.json)
    traffic, traffic_src = None, None
    pj = os.path.join(ROOT, 'profiles', 'summary.json')
    if os.path.exists(pj):
        try:
            ent = json.load(open(pj)).get('lbfgs_kernel_bench', {})
            if ent.get('batch') == Bg and ent.get('max_iter') == a.max_iter:
                traffic, traffic_src = ent.get('dram_bytes_per_launch'), ent.get('capture')
        except Exception:
            traffic = None
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': dmma, 'unit': 'TFLOP/s', 'frac': achieved / dmma,
                'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': 'FP64 DMMA (mma.sync.m8n8k4.f64) peak measured live by bdrt_peak_fp64 on this GPU; '
                               'MEASURED_PEAKS.json has no FP64 figure (bf16/HBM only); FP64 FMA-pipe peak measured '
                               '%.1f TFLOP/s' % dfma,
                'kernel': 'lbfgs_kernel', 'flop_model': 'n_eval*84000 (banded-L F_grad, SURVEY 8d) + n_iter*8*m*D',
                'grad_evals_per_spectrum': n_eval_tot / (Bg * a.steps),
                'grad_evals_per_s': n_eval_tot / (kern_ms * 1e-3)}

    # -------------------------------------------------------------------------------- end-to-end leg ("e2e")
    # the call a user makes: Inverter.fit(freq, Z) with HOST buffers; H2D of the spectra and D2H of the results inside
    out_host = torch.empty((Bg, prob.K + 6), dtype=torch.float64).pin_memory()

    def step_e2e():
        iv = Inverter(basis_freq=bf.numpy(), device=dev)  # fresh instance: matrices are rebuilt every step
        iv.fit(freq, Z_host, mode='optimize', max_iter=a.max_iter, spectrum_ids=ids_dev, check_outliers=False)
        packed = torch.cat((iv.distribution_fits['DRT']['coef'], iv.R_inf[:, None], iv.inductance[:, None],
                            iv.error_fit['sigma_res'][:, None], iv.error_fit['alpha_prop'][:, None],
                            iv.error_fit['alpha_re'][:, None], iv.error_fit['alpha_im'][:, None]), dim=1)
        if ws > 1:
            packed = gather_results(packed, indices=ids_dev)[ids_dev]
        out_host.copy_(packed, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    step_e2e()
    barrier()
    e0.record()
    for _ in range(a.steps):
        flush.zero_()
        step_e2e()
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = ws * Bg * a.steps / (t2.item() * 1e-3)
    e2e = {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(Z_host.numel() * 16 + freq.numel() * 8),
           'd2h_bytes_per_step': int(out_host.numel() * 8), 'api': 'bayes_drt_b200.Inverter.fit(freq, Z_host)'}

    # -------------------------------------------------------------------------------- HMC (second headline metric)
    def hmc_leg(Bh, chains):
        ivh = Inverter(basis_freq=bf.numpy(), device=dev)
        Zh = Z_host[:Bh]
        th, _ = timed(lambda: ivh.fit(freq, Zh, mode='sample', chains=chains, warmup=200, samples=200,
                                      spectrum_ids=ids_dev[:Bh], check_outliers=False, keep_draws=False))
        st = ivh._sample_stats
        K = prob.K
        min_ess = st['ess_bulk'][:, :K].min(dim=1).values  # min over the DRT coefficients (bdrt_diagnostics)
        tot = torch.tensor([float(st['n_leapfrog'].sum().item()), float(min_ess.sum().item()),
                            float(st['n_divergent'].sum().item()), float(st['n_maxdepth'].sum().item()),
                            float((st['rhat'][:, :K].max(dim=1).values < 1.05).sum().item())],
                           dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(tot)
        waves = Bh * chains / (16.0 * torch.cuda.get_device_properties(dev).multi_processor_count)  # 2 CTAs x 8 slots / SM
        return {'metric': 'ESS/sec (HMC)', 'value': tot[1].item() / th, 'unit': 'min-bulk-ESS/s',
                'inversions_per_s': ws * Bh / th, 'spectra': ws * Bh, 'chains': chains, 'warmup': 200, 'samples': 200,
                'waves_of_resident_chains': waves, 'seconds': th, 'grad_evals_per_s': tot[0].item() / th,
                'achieved_tflops': tot[0].item() * (F_GRAD + 6 * prob.D) / th / 1e12 / ws,
                'frac_of_dmma_peak': tot[0].item() * (F_GRAD + 6 * prob.D) / th / 1e12 / ws / dmma,
                'mean_min_ess_per_inversion': tot[1].item() / (ws * Bh),
                'frac_inversions_rhat_below_1.05': tot[4].item() / (ws * Bh),
                'divergent_frac': tot[2].item() / (ws * Bh * chains * 200),
                'maxdepth_frac': tot[3].item() / (ws * Bh * chains * 200)}

    hmc = hmc4 = None
    if not a.no_hmc:
        ivw = Inverter(basis_freq=bf.numpy(), device=dev)
        ivw.fit(freq, Z_host[:16], mode='sample', warmup=20, samples=10, check_outliers=False)  # warm-up launch
        hmc = hmc_leg(a.hmc_batch, 2)          # the reference's default: 2 chains, one wave of resident chains
        if not a.quick:
            hmc4 = hmc_leg(a.hmc4_batch, 4)    # config 4's 4-chain variant, three waves (tail / refill included)

    # -------------------------------------------------------------------------------- the reference's recommended flows
    flows = None
    if not a.quick:
        flows = {}
        med = lp_random.median()
        flows['random_init'] = {'spectra_per_s': value / ws, 'grad_evals_per_spectrum': n_eval_tot / (Bg * a.steps),
                                'poor_optima_frac': float((lp_random < med - 200).float().mean().item())}
        # (a) fit(init_from_ridge=True): under-fitted ridge solution as the start of x, R_inf, inductance
        #     (inversion.py:1154-1187, :1616-1682)
        for key, kw in (('init_from_ridge', dict(init_from_ridge=True)),
                        ('init_from_ridge_outliers_auto', dict(init_from_ridge=True, outliers='auto'))):
            ivf = Inverter(basis_freq=bf.numpy(), device=dev)
            with __import__('warnings').catch_warnings():
                __import__('warnings').simplefilter('ignore')
                tf, _ = timed(lambda: ivf.fit(freq, Z_host, mode='optimize', max_iter=a.max_iter, spectrum_ids=ids_dev,
                                              check_outliers=False, **kw))
            o = ivf._opt_result
            flows[key] = {'spectra_per_s': Bg / tf, 'seconds': tf,
                          'grad_evals_per_spectrum': float(o['n_eval'].float().mean().item()),
                          'max_iterations': int(o['iters'].max().item()),
                          'poor_optima_frac': float((o['lp'] < med - 200).float().mean().item()),
                          'outlier_model_frac': float(ivf._outlier_model.float().mean().item()),
                          'ridge_hyper_iterations': float(ivf._ridge_iters.float().mean().item())
                          if hasattr(ivf, '_ridge_iters') else None}
        # (b) information only (not Stan semantics): L-BFGS cut at 2000 iterations, then the Newton polish to the exact
        #     optimum; distance of Stan's own end point from that optimum
        def trunc_newton():
            r1 = prob.map_lbfgs(u0, max_iter=2000)
            return r1, prob.map_newton(r1['u'])
        tn, (r1, pn) = timed(trunc_newton)
        xs = prob.split_outputs(prob.constrain(pn['u']))['x']
        xl = prob.split_outputs(prob.constrain(r['u']))['x']
        okn = pn['gnorm'] < 1e-7
        dist_l = ((xl - xs).abs().max(dim=1).values / xs.abs().max(dim=1).values)[okn]
        flows['lbfgs2000_then_newton'] = {
            'spectra_per_s': Bg / tn, 'seconds': tn,
            'grad_evals_per_spectrum': float((r1['n_eval'].float() + pn['n_eval'].float()).mean().item()),
            'newton_converged_frac': float(okn.float().mean().item()),
            'stan_lbfgs_rel_distance_to_optimum_median': float(dist_l.median().item()) if dist_l.numel() else None,
            'stan_lbfgs_rel_distance_to_optimum_p95': float(dist_l.quantile(0.95).item()) if dist_l.numel() else None}

    # -------------------------------------------------------------------------------- other solvers of the path (info)
    extras = None
    if not a.no_hmc:
        # hyper-parametric ridge (Inverter.ridge_fit defaults: discrete penalty; preset 'Huang': integral penalty, modulus
        # weights -> per-spectrum Gram matrices), factorisations counted on the device
        Br = min(Bg, 8192)
        n_r, nf_r = len(bf) + 2, len(freq)
        ridge = {}
        for key, kw in (('default', dict()), ('default_reference_stop_rule', dict(stop_rule='nan')),
                        ('preset_Huang', dict(preset='Huang'))):
            ivr = Inverter(basis_freq=bf.numpy(), device=dev)
            ivr.ridge_fit(freq, Z_host[:Br], **kw)
            t_r, _ = timed(lambda: ivr.ridge_fit(freq, Z_host[:Br], **kw))
            it_r = float(ivr._ridge_iters.float().mean().item())
            fac = float(ivr._ridge_factorisations.float().mean().item())
            # Gram matrix 2 (2 Nf) n^2 (per spectrum only with per-spectrum weights), per hyper-iteration the
            # lambda update and the weighted penalty 4 n^2, per factorisation n^3 / 3 + two triangular solves and the
            # multipliers 4 n^2
            fl = (2.0 * 2 * nf_r * n_r ** 2 if 'preset' in kw else 0.0) + it_r * 4.0 * n_r ** 2 + \
                fac * (n_r ** 3 / 3.0 + 4.0 * n_r ** 2)
            ridge[key] = {'fits_per_s': ws * Br / t_r, 'hyper_iterations': it_r, 'factorisations': fac,
                          'converged_frac': float(ivr._ridge_converged.float().mean().item()),
                          'tflops': Br * fl / t_r / 1e12, 'frac_of_dmma_peak': Br * fl / t_r / 1e12 / dmma}
        # kernel matrices on per-spectrum grids (config 5 shape): A_re + A_im of 2048 grids, 81 x 81, trapezoid over the
        # reference's node set restricted to the Gaussian window (76 of 1000 nodes at the default epsilon)
        Gm = 2048
        fg = 10.0 ** (6.0 - torch.rand(Gm, 1, dtype=torch.float64) - torch.arange(81, dtype=torch.float64)[None, :] / 10.0)
        taum = torch.as_tensor(1.0 / (2 * np.pi * np.logspace(6, -2, 81)))
        epsm = 1.0 / float(np.mean(np.diff(np.log(taum.numpy()))))
        fgd = fg.to(dev)
        capi.build_A(fgd[:8], taum, epsm, device=dev)
        t_m, _ = timed(lambda: capi.build_A(fgd, taum, epsm, device=dev))
        nodes = 76
        extras = {'ridge': ridge, 'ridge_batch': Br,
                  'A_builds_per_s': ws * 2 * Gm / t_m, 'A_build_shape': [81, 81],
                  'A_build_tflops': Gm * 81 * 81 * nodes * 12 / t_m / 1e12,  # 12 flop per entry and node for both parts
                  'A_build_frac_of_dfma_peak': Gm * 81 * 81 * nodes * 12 / t_m / 1e12 / dfma,
                  # the same launches in the reference's own accounting (all 1000 trapezoid nodes, SURVEY 8d "Q_ref")
                  'A_build_tflops_reference_literal': Gm * 81 * 81 * 1000 * 12 / t_m / 1e12}

    # -------------------------------------------------------------------------------- config 5 (per-spectrum grids)
    config5 = None
    if not a.quick:
        from bayes_drt_b200.synth import make_spectra_sp, sp_distributions
        B5, B5h = a.config5_batch, a.config5_hmc_batch
        f5, Z5 = make_spectra_sp(max(B5, B5h), seed=20240605 + rank)
        iv5 = Inverter(distributions=sp_distributions(), device=dev)
        iv5.fit(f5[:8], Z5[:8], nonneg=True, mode='optimize', max_iter=50, check_outliers=False)
        t5, _ = timed(lambda: iv5.fit(f5[:B5], Z5[:B5], nonneg=True, mode='optimize', max_iter=a.max_iter,
                                      check_outliers=False))
        o5 = iv5._opt_result
        ev5 = float(o5['n_eval'].float().sum().item())
        t5h, _ = timed(lambda: iv5.fit(f5[:B5h], Z5[:B5h], nonneg=True, mode='sample', chains=2, warmup=200, samples=200,
                                       check_outliers=False, keep_draws=False))
        s5 = iv5._sample_stats
        config5 = {'workload': 'Series-Parallel_pos (DRT + planar transmissive DDT, Ks = Kp = 81, D = 336), every spectrum on '
                               'its own grid 10**(6 - delta - arange(81)/10), kernel matrices built per spectrum inside the '
                               'timed region; per GPU',
                   'map_spectra_per_s': B5 / t5, 'map_batch': B5, 'map_grad_evals_per_spectrum': ev5 / B5,
                   'map_grad_evals_per_s': ev5 / t5,
                   'hmc_inversions_per_s': B5h / t5h, 'hmc_batch': B5h, 'hmc_chains': 2,
                   'hmc_grad_evals_per_s': float(s5['n_leapfrog'].sum().item()) / t5h,
                   'hmc_min_bulk_ess_per_s': float(s5['ess_bulk'][:, :162].min(dim=1).values.sum().item()) / t5h,
                   'hmc_maxdepth_frac': float(s5['n_maxdepth'].float().mean().item()) / 200,
                   'hmc_divergent_frac': float(s5['n_divergent'].float().mean().item()) / 200}

    # -------------------------------------------------------------------------------- CPU baselines (rank 0, N = 1)
    cpu = None
    if rank == 0 and ws == 1:
        cores = os.cpu_count()
        n = a.cpu_sample or 2 * cores
        v, cused, dt, nev = cpu_map_throughput(n, a.max_iter, cores)
        cpu = {'value': v, 'unit': UNIT, 'cores': cused, 'kind': 'port',
               'sample': '%d spectra of the same workload, one process per core, %.1f s wall, %.0f gradient '
                         'evaluations per spectrum (oracle restatement of the Stan model + Stan L-BFGS in numpy; '
                         'pystan is not installable offline)' % (n, dt, nev)}
        if hmc is not None and not a.quick:
            nh = max(1, cores // 2)  # one chain per core, two chains per inversion
            ess_s, inv_s, cu, dth, ng = cpu_hmc_throughput(nh, 2, 200, 200, cores)
            hmc['cpu_baseline'] = {'value': ess_s, 'unit': 'min-bulk-ESS/s', 'inversions_per_s': inv_s, 'cores': cu,
                                   'kind': 'port',
                                   'sample': '%d inversions (2 chains x (200 + 200) each) of the same workload, one chain '
                                             'per core, %.1f s wall, %.2e gradient evaluations (oracle restatement of '
                                             "Stan's NUTS in numpy)" % (nh, dth, ng)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': ws, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': t_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'config4: %d spectra/GPU/step (1e5-spectrum sweep = 12500/GPU at 8 GPUs), shared '
                                   'grid Nf=70, K=100, D=209, model Series, mode=optimize, Stan-semantics L-BFGS '
                                   '(history 5, Stan tolerances, iter cap %d), random inits U(-2,2), strided shards'
                                   % (Bg, a.max_iter),
                       'l2_flush': 'explicit 256 MiB write between timed steps',
                       'termination': {str(k): v for k, v in term.items()}},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'hmc': hmc, 'hmc_4chains': hmc4, 'flows': flows, 'config5': config5, 'extras': extras,
        }
        _emit(line)
    if ws > 1:
        dist.destroy_process_group()


def _emit(line):
    """The JSON line is the only thing that goes to the real stdout (libraries such as NCCL print banners to fd 1)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + '\n').encode())


if __name__ == '__main__':
    # everything else that is written to stdout (NCCL version banner, warnings of child processes) goes to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
