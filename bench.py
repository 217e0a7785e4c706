#!/usr/bin/env python
"""Benchmark of the bayes-drt inversion hot path on B200 (BASELINE.json metric: spectra/sec (MAP), ESS/sec (HMC)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--no-hmc]

Workload (config 4 of BASELINE.json): synthetic ZARC / RC spectra sharing one frequency grid
(Nf = 70, K = 100 basis functions -> D = 209 parameters), reference defaults of Inverter.fit(mode='optimize'): model
'Series', Stan-semantics L-BFGS (history 5, Stan's tolerances, iter cap 50000) from Stan-style random inits U(-2, 2).
Weak scaling: every GPU gets --batch spectra per step (default 12500 = the per-GPU share of the 1e5-spectrum sweep at
8 GPUs); one "step" is one full MAP inversion of that batch.  One JSON line is printed by rank 0.
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


def run_ours(a):
    import torch.distributed as dist
    from bayes_drt_b200 import Inverter, capi, diagnostics, synth
    from bayes_drt_b200.distributed import gather_results
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

    Bg = a.batch
    start = rank * Bg
    freq, Z_all, _ = synth.make_spectra(ws * Bg, seed=20240601)  # identical on every rank; each takes its block
    _, bf = synth.bench_grid()
    Z_host = Z_all[start:start + Bg].contiguous().pin_memory()
    ctx = context(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # -------------------------------------------------------------------------------- kernel-only leg ("value")
    inv = Inverter(basis_freq=bf.numpy(), device=dev)
    Zd = Z_host.to(dev)
    from bayes_drt_b200.inverter import _MODE, _hash_uniform
    fsorted, Zb = inv._to_batch(freq, Zd)
    Zs = inv._scale_Z(Zb, True)
    tau, eps, m = inv._grid(fsorted, 'DRT')
    c = _MODE['optimize']
    L = torch.stack([c['l'][j] * m[f'L{j}'] for j in range(3)])
    prob = capi.SeriesProblem(torch.cat((m['A_re'], m['A_im'])), torch.cat((Zs.real, Zs.imag), dim=1).contiguous(),
                              fsorted, L, device=dev)
    u0 = _hash_uniform(1234, start, Bg, prob.D, dev)

    def step_resident():
        r = prob.map_lbfgs(u0, max_iter=a.max_iter)
        out = prob.constrain(r['u'])
        res = torch.cat((out[:, :prob.K + 6], r['lp'][:, None]), dim=1)
        return r, gather_results(res) if ws > 1 else res

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

    # roofline of the dominant kernel (lbfgs_kernel): algorithmic flops / device time of the step on this rank
    dfma, dmma = capi.peak_fp64(dev)
    flops = n_eval_tot * F_GRAD + n_iter_tot * 8.0 * 5 * prob.D  # gradients + two-loop recursion (4 m D MACs)
    achieved = flops / (kern_ms * 1e-3) / 1e12
    traffic = None
    pj = os.path.join(ROOT, 'profiles', 'summary.json')
    if os.path.exists(pj):
        try:
            traffic = json.load(open(pj)).get('lbfgs_kernel', {}).get('dram_bytes_per_launch')
        except Exception:
            traffic = None
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': dmma, 'unit': 'TFLOP/s', 'frac': achieved / dmma,
                'traffic': traffic,
                'peak_source': 'FP64 DMMA (mma.sync.m8n8k4.f64) peak measured live by bdrt_peak_fp64 on this GPU; '
                               'MEASURED_PEAKS.json has no FP64 figure (bf16/HBM only); FP64 FMA-pipe peak measured '
                               '%.1f TFLOP/s' % dfma,
                'kernel': 'lbfgs_kernel', 'flop_model': 'n_eval*84000 (banded-L F_grad, SURVEY 8d) + n_iter*8*m*D',
                'grad_evals_per_spectrum': n_eval_tot / (Bg * a.steps)}

    # -------------------------------------------------------------------------------- end-to-end leg ("e2e")
    # the call a user makes: Inverter.fit(freq, Z) with HOST buffers; H2D of the spectra and D2H of the results inside
    out_host = torch.empty((Bg, prob.K + 6), dtype=torch.float64).pin_memory()

    def step_e2e():
        iv = Inverter(basis_freq=bf.numpy(), device=dev)  # fresh instance: matrices are rebuilt every step
        iv.fit(freq, Z_host, mode='optimize', max_iter=a.max_iter, spectrum_offset=start, check_outliers=False)
        packed = torch.cat((iv.distribution_fits['DRT']['coef'], iv.R_inf[:, None], iv.inductance[:, None],
                            iv.error_fit['sigma_res'][:, None], iv.error_fit['alpha_prop'][:, None],
                            iv.error_fit['alpha_re'][:, None], iv.error_fit['alpha_im'][:, None]), dim=1)
        if ws > 1:
            packed = gather_results(packed)[start:start + Bg]
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
    hmc = None
    if not a.no_hmc:
        Bh = a.hmc_batch
        ivh = Inverter(basis_freq=bf.numpy(), device=dev)
        Zh = Z_host[:Bh]
        ivh.fit(freq, Zh[:16], mode='sample', warmup=20, samples=10, check_outliers=False)  # warm-up launch
        barrier()
        e0.record()
        ivh.fit(freq, Zh, mode='sample', chains=2, warmup=200, samples=200, spectrum_offset=start,
                check_outliers=False)
        e1.record()
        barrier()
        th = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(th, op=dist.ReduceOp.MAX)
        th = th.item() * 1e-3
        x = ivh._sample_result['x'].reshape(Bh, 2, 200, -1).permute(0, 3, 1, 2)  # [B, K, chains, n]
        ess = diagnostics.ess_bulk(x)  # [B, K]
        min_ess = ess.min(dim=1).values
        st = ivh._sample_stats
        ngrad = torch.tensor([float(st['n_leapfrog'].sum().item()), float(min_ess.sum().item()),
                              float(st['n_divergent'].sum().item()), float(st['n_maxdepth'].sum().item())],
                             dtype=torch.float64, device=dev)
        if ws > 1:
            dist.all_reduce(ngrad)
        hmc = {'metric': 'ESS/sec (HMC)', 'value': ngrad[1].item() / th, 'unit': 'min-bulk-ESS/s',
               'inversions_per_s': ws * Bh / th, 'spectra': ws * Bh, 'chains': 2, 'warmup': 200, 'samples': 200,
               'seconds': th, 'grad_evals_per_s': ngrad[0].item() / th,
               'achieved_tflops': ngrad[0].item() * (F_GRAD + 6 * prob.D) / th / 1e12 / ws,
               'frac_of_dmma_peak': ngrad[0].item() * (F_GRAD + 6 * prob.D) / th / 1e12 / ws / dmma,
               'mean_min_ess_per_inversion': ngrad[1].item() / (ws * Bh),
               'divergent_frac': ngrad[2].item() / (ws * Bh * 2 * 200),
               'maxdepth_frac': ngrad[3].item() / (ws * Bh * 2 * 200)}

    # -------------------------------------------------------------------------------- other solvers of the path (info)
    extras = None
    if not a.no_hmc:
        # hyper-parametric ridge (Inverter.ridge_fit defaults: discrete penalty, 20 hyper-iterations, exact QP)
        ivr = Inverter(basis_freq=bf.numpy(), device=dev)
        Br = min(Bg, 4096)
        ivr.ridge_fit(freq, Z_host[:64])
        barrier()
        e0.record()
        ivr.ridge_fit(freq, Z_host[:Br])
        e1.record()
        barrier()
        t_r = e0.elapsed_time(e1) * 1e-3
        # kernel matrices on per-spectrum grids (config 5 shape): A_re + A_im of 2048 grids, 81 x 81, trapezoid over the
        # reference's node set restricted to the Gaussian window (76 of 1000 nodes at the default epsilon)
        Gm = 2048
        fg = 10.0 ** (6.0 - torch.rand(Gm, 1, dtype=torch.float64) - torch.arange(81, dtype=torch.float64)[None, :] / 10.0)
        taum = torch.as_tensor(1.0 / (2 * np.pi * np.logspace(6, -2, 81)))
        epsm = 1.0 / float(np.mean(np.diff(np.log(taum.numpy()))))
        fgd = fg.to(dev)
        capi.build_A(fgd[:8], taum, epsm, device=dev)
        barrier()
        e0.record()
        capi.build_A(fgd, taum, epsm, device=dev)
        e1.record()
        barrier()
        t_m = e0.elapsed_time(e1) * 1e-3
        nodes = 76
        # ridge work, SURVEY 8d: P assembly 2 (2 Nf) n^2 once; per hyper-iteration the lambda-weighted penalty 2 n^2 and at
        # least one Cholesky n^3 / 3 + two triangular solves 2 n^2 (the active-set QP may factor more than once per
        # hyper-iteration and its pivots are not counted on device: this is a lower bound)
        n_r, nf_r = len(bf) + 2, len(freq)
        it_r = float(ivr._ridge_iters.float().mean().item())
        ridge_flop = 2.0 * (2 * nf_r) * n_r ** 2 + it_r * (n_r ** 3 / 3.0 + 4.0 * n_r ** 2)
        extras = {'ridge_fits_per_s': ws * Br / t_r, 'ridge_batch': Br,
                  'ridge_hyper_iterations': it_r,
                  'ridge_tflops_lower_bound': Br * ridge_flop / t_r / 1e12,
                  'ridge_frac_of_dfma_peak_lower_bound': Br * ridge_flop / t_r / 1e12 / dfma,
                  'A_builds_per_s': ws * 2 * Gm / t_m, 'A_build_shape': [81, 81],
                  'A_build_tflops': Gm * 81 * 81 * nodes * 12 / t_m / 1e12,  # 12 flop per entry and node for both parts
                  'A_build_frac_of_dfma_peak': Gm * 81 * 81 * nodes * 12 / t_m / 1e12 / dfma,
                  # the same launches in the reference's own accounting (all 1000 trapezoid nodes, SURVEY 8d "Q_ref")
                  'A_build_tflops_reference_literal': Gm * 81 * 81 * 1000 * 12 / t_m / 1e12}

    # -------------------------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and ws == 1:
        cores = os.cpu_count()
        n = a.cpu_sample or 2 * cores
        v, cused, dt, nev = cpu_map_throughput(n, a.max_iter, cores)
        cpu = {'value': v, 'unit': UNIT, 'cores': cused, 'kind': 'port',
               'sample': '%d spectra of the same workload, one process per core, %.1f s wall, %.0f gradient '
                         'evaluations per spectrum (oracle restatement of the Stan model + Stan L-BFGS in numpy; '
                         'pystan is not installable offline)' % (n, dt, nev)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': ws, 'steps': a.steps, 'warmup': a.warmup,
            'ms_per_step': t_ms / a.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': 'config4: %d spectra/GPU/step (1e5-spectrum sweep = 12500/GPU at 8 GPUs), shared '
                                   'grid Nf=70, K=100, D=209, model Series, mode=optimize, Stan-semantics L-BFGS '
                                   '(history 5, Stan tolerances, iter cap %d), random inits U(-2,2)' % (Bg, a.max_iter),
                       'l2_flush': 'explicit 256 MiB write between timed steps',
                       'termination': {str(k): v for k, v in term.items()}},
            'e2e': e2e, 'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'hmc': hmc, 'extras': extras,
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
