"""CPU: host-side logic of the package that needs no GPU -- shard arithmetic, the world_size-2 gather over gloo, the
hash-based Stan-style random init, the ESS estimator against the oracle's, and the reference-interface error checks."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything():
    from bayes_drt_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 100000):
        for ws in (1, 2, 3, 8):
            r = [shard_range(n, k, ws) for k in range(ws)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def _worker(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bayes_drt_b200.distributed import gather_results, shard_range
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=ws)
    n_total = 11  # ragged
    a, b = shard_range(n_total)
    local = torch.arange(a, b, dtype=torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
    out = gather_results(local, n_total)
    q.put((rank, out[:, 0].tolist()))
    dist.destroy_process_group()


def test_shard_indices_strided_and_block():
    from bayes_drt_b200.distributed import shard_indices, shard_range
    for n in (0, 1, 7, 11, 1000):
        for ws in (1, 2, 3, 8):
            parts = [shard_indices(n, k, ws) for k in range(ws)]
            assert sorted(sum((p.tolist() for p in parts), [])) == list(range(n))
            assert all(p.tolist() == list(range(k, n, ws)) for k, p in enumerate(parts))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for k in range(ws):
                a, b = shard_range(n, k, ws)
                assert shard_indices(n, k, ws, mode='block').tolist() == list(range(a, b))


def _worker_strided(rank, ws, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from bayes_drt_b200.distributed import gather_results, shard_indices
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=ws)
    n_total = 11  # ragged: rank 0 holds 6 rows, rank 1 holds 5
    idx = shard_indices(n_total)
    local = idx.to(torch.float64)[:, None] * torch.ones(1, 3, dtype=torch.float64)
    out = gather_results(local, n_total, indices=idx)
    q.put((rank, out[:, 0].tolist()))
    dist.destroy_process_group()


def test_gather_results_strided_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker_strided, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in ps]
    for rank, vals in res:
        assert vals == [float(i) for i in range(11)]


def test_gather_results_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in ps]
    for rank, vals in res:
        assert vals == [float(i) for i in range(11)]


def test_hash_uniform_is_index_keyed():
    from bayes_drt_b200.inverter import _hash_uniform
    u = _hash_uniform(1234, 0, 1000, 209, 'cpu')
    assert u.min() >= -2 and u.max() <= 2
    assert abs(u.mean().item()) < 0.02 and abs(u.std().item() - 4 / 12 ** 0.5) < 0.02
    # shard invariance: rows 300..400 generated on their own are identical
    assert torch.equal(u[300:400], _hash_uniform(1234, 300, 100, 209, 'cpu'))
    assert not torch.equal(u, _hash_uniform(1235, 0, 1000, 209, 'cpu'))
    # no visible correlation between neighbouring coordinates
    c = np.corrcoef(u[:, :-1].flatten().numpy(), u[:, 1:].flatten().numpy())[0, 1]
    assert abs(c) < 0.01


def test_ess_matches_oracle():
    from bayes_drt_b200.diagnostics import ess_bulk
    from oracle.nuts import ess_bulk as oess
    rng = np.random.RandomState(0)
    x = np.zeros((3, 4, 400))
    for i, phi in enumerate((0.0, 0.6, 0.95)):
        e = rng.standard_normal((4, 400))
        for t in range(1, 400):
            e[:, t] = phi * e[:, t - 1] + np.sqrt(1 - phi ** 2) * e[:, t]
        x[i] = e
    got = ess_bulk(torch.tensor(x)).numpy()
    want = np.array([oess(x[i]) for i in range(3)])
    assert np.allclose(got, want, rtol=0.02), (got, want)


def test_reference_interface_errors_without_gpu():
    from bayes_drt_b200 import matrices as m
    with pytest.raises(ValueError):
        m.construct_A(np.logspace(3, 0, 5), 'real', basis='Zic')
    with pytest.raises(ValueError):
        m.construct_A(np.logspace(3, 0, 5), 'real', kernel='DRT', dist_type='parallel')
    with pytest.raises(ValueError):
        m.construct_A(np.logspace(3, 0, 5), 'real', kernel='XYZ')
    with pytest.raises(ValueError):
        m.construct_M(np.logspace(3, 0, 5), order=3)
    assert m.is_loguniform(np.logspace(3, 0, 5)) and m.is_loguniform([9.0, 5.0, 4.9, 1.0])  # the reference's quirk
    r = m.rel_round(np.array([123456.789012345, 0.00123456789012345]), 10)
    # utils.py:113-131: round(x, precision - floor(log10 x)) decimals
    assert r[0] == 123456.78901 and r[1] == round(0.00123456789012345, 13)


def test_ridge_weight_forms_match_the_oracle():
    """ridge._weights (Inverter._format_weights, inversion.py:2338-2395) on CPU tensors: named schemes against the
    oracle's restatement, constants and arrays by their definition."""
    import numpy as np
    import torch
    from bayes_drt_b200.ridge import _weights
    from oracle import ridge as oridge
    rng = np.random.RandomState(0)
    Z = rng.randn(3, 11) + 1j * rng.randn(3, 11)
    Zt = torch.tensor(Z)
    for scheme in (None, 'unity', 'modulus', 'Orazem', 'proportional'):
        w_re, w_im = _weights(Zt, scheme)
        for b in range(3):
            w = oridge.format_weights(Z[b], scheme)
            assert np.allclose(w_re[b].numpy(), w.real, rtol=1e-15) and np.allclose(w_im[b].numpy(), w.imag, rtol=1e-15)
    w_re, w_im = _weights(Zt, 'prop_adj')
    p25 = np.percentile(np.abs(Z) ** 2, 25, axis=1)[:, None]
    assert np.allclose(w_re.numpy(), 1 / (np.abs(Z.real) + p25)) and np.allclose(w_im.numpy(), 1 / (np.abs(Z.imag) + p25))
    w_re, w_im = _weights(Zt, 2.5)
    assert bool((w_re == 2.5).all()) and bool((w_im == 2.5).all())
    w_re, w_im = _weights(Zt, 1 + 3j)
    assert bool((w_re == 1).all()) and bool((w_im == 3).all())
    a = rng.rand(11)
    w_re, w_im = _weights(Zt, a)
    assert tuple(w_re.shape) == (3, 11) and np.array_equal(w_re[2].numpy(), a) and np.array_equal(w_im[0].numpy(), a)
    ac = rng.rand(3, 11) + 1j * rng.rand(3, 11)
    w_re, w_im = _weights(Zt, ac)
    assert np.array_equal(w_re.numpy(), ac.real) and np.array_equal(w_im.numpy(), ac.imag)
    import pytest
    with pytest.raises(ValueError):
        _weights(Zt, 'nope')
    with pytest.raises(ValueError):
        _weights(Zt, np.ones(5))


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys, timed on a bounded sample of the workload."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup',
                        '0', '--cpu-sample', '2'], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'spectra/sec (MAP)' and d['unit'] == 'spectra/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']


def test_numeric_helpers_against_reference_vectors():
    """tests/golden/utils.npz holds outputs of the reference's utils.py (rel_round, is_loguniform, get_outlier_thresh,
    r2_score; generated by scripts/make_golden_utils.py).  The host mirrors and the torch expressions Inverter uses for
    the inter-quartile rule (check_outliers) and r^2 (score) reproduce them."""
    import os
    import numpy as np
    import torch
    from bayes_drt_b200 import matrices as m
    from oracle import matrices as om
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'utils.npz'))
    for prec in (3, 10):
        assert np.array_equal(m.rel_round(g['rr/x'], prec), g[f'rr/p{prec}'])
    i = 0
    while f'lu/g{i}' in g.files:
        assert m.is_loguniform(g[f'lu/g{i}']) == bool(g[f'lu/r{i}']) == om.is_loguniform(g[f'lu/g{i}']), i
        i += 1
    assert i == 7
    y = torch.tensor(g['ot/y'])[None, :]
    for f in (1.5, 3.5, 4.0):  # Inverter._ridge_outlier_flags: q75 + f * (q75 - q25), utils.py:143-146
        q = torch.quantile(y, torch.tensor([0.25, 0.75], dtype=torch.float64), dim=1)
        assert abs(float(q[1] + f * (q[1] - q[0])) - float(g[f'ot/t{f}'])) <= 1e-12 * float(g[f'ot/t{f}'])
    yt, yh, w = (torch.tensor(g[k])[None, :] for k in ('r2/y', 'r2/yhat', 'r2/w'))
    for wt, key in ((torch.ones_like(w), 'r2/plain'), (w, 'r2/weighted')):  # Inverter.score, utils.py:149-165
        avg = (wt * yt).sum(dim=1, keepdim=True) / wt.sum(dim=1, keepdim=True)
        r2 = 1 - (wt * (yh - yt) ** 2).sum(dim=1) / (wt * (yt - avg) ** 2).sum(dim=1)
        assert abs(float(r2) - float(g[key])) <= 1e-13


def test_bench_reference_arm_under_torchrun_prints_once():
    """launched the way the driver launches N > 1 (torchrun, one rank per GPU): rank 0 alone runs the CPU arm and prints
    its line, the other ranks exit 0 without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', '29537', os.path.join(root, 'bench.py'), '--impl',
                        'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--cpu-sample', '2'],
                       capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['value'] > 0
