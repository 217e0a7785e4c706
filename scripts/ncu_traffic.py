"""DRAM traffic of lbfgs_kernel on the benchmark's own launch shape -> profiles/summary.json['lbfgs_kernel_bench'].

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:lbfgs_kernel -c 2 --csv --log-file gpurun_out/traffic_<tag>.csv python bench.py --quick --no-hmc ...
    python scripts/ncu_traffic.py <tag> gpurun_out/traffic_<tag>.csv <batch> <max_iter>
"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, path, batch, max_iter = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]
iN, iM, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
iI = hdr.index('ID')
mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6,
        'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
per = {}
for r in rows[1:]:
    if 'lbfgs_kernel' not in r[iN]:
        continue
    d = per.setdefault(r[iI], {})
    d[r[iM]] = float(r[iV].replace(',', '')) * mult.get(r[iU], 1.0)
vals = list(per.values())
assert vals, 'no lbfgs_kernel launch in ' + path
dram = sum(v['dram__bytes_read.sum'] + v['dram__bytes_write.sum'] for v in vals) / len(vals)
dur = sum(v['gpu__time_duration.sum'] for v in vals) / len(vals)
sj = os.path.join(ROOT, 'profiles', 'summary.json')
summ = json.load(open(sj)) if os.path.exists(sj) else {}
summ['lbfgs_kernel_bench'] = {'round': tag, 'batch': batch, 'max_iter': max_iter, 'launches': len(vals),
                              'dram_bytes_per_launch': dram, 'duration_ms_under_ncu': dur,
                              'dram_read_bytes': sum(v['dram__bytes_read.sum'] for v in vals) / len(vals),
                              'dram_write_bytes': sum(v['dram__bytes_write.sum'] for v in vals) / len(vals),
                              'capture': os.path.basename(path)}
json.dump(summ, open(sj, 'w'), indent=1, sort_keys=True)
print(summ['lbfgs_kernel_bench'])
