"""Time bdrt_summarize (HBM-bound: every draw read once) against the measured copy bandwidth."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bayes_drt_b200 import capi
G, S, P = 4000, 400, 246   # 3.15 GB of constrained draws (2 chains x 200 draws, Series model at the benchmark shape)
x = torch.randn(G, S, P, dtype=torch.float64, device='cuda')
peak = 6551.7
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mean, q = capi.summarize(x, percentiles=(2.5, 50, 97.5))
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1)
    byt = x.numel() * 8 + 4 * G * P * 8
    print(f'summarize G={G} S={S} P={P}: {ms:.3f} ms  {byt/ms/1e6:.1f} GB/s  ({100*byt/ms/1e6/peak:.1f} % of measured HBM copy {peak} GB/s)')
t = time.time(); m2 = x.mean(dim=1); q2 = torch.quantile(x[:500], torch.tensor([0.025, 0.5, 0.975], dtype=torch.float64, device='cuda'), dim=1); torch.cuda.synchronize()
print('check', (mean - m2).abs().max().item(), (q[:, :500] - q2).abs().max().item())
