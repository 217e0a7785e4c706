"""Summarise ncu reports into profiles/: python scripts/ncu_summary.py <round tag> <kernel>=<report.ncu-rep> ...
Writes profiles/<tag>_<kernel>.txt (key raw metrics + hottest source lines) and updates profiles/summary.json
(per-kernel DRAM bytes per launch, duration, pipe utilisation) which bench.py reads for roofline.traffic."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        res.append({h: (vals[i], units[i]) for i, h in enumerate(hdr)})
    return res


def main():
    tag = sys.argv[1]
    sj = os.path.join(ROOT, 'profiles', 'summary.json')
    summ = json.load(open(sj)) if os.path.exists(sj) else {}
    for arg in sys.argv[2:]:
        name, rep = arg.split('=')
        launches = raw(rep)
        r = launches[0]
        lines = [f'# {name}  ({os.path.basename(rep)}, ncu --set full --clock-control none; round {tag})',
                 f'# kernel: {r["Kernel Name"][0]}  launches captured: {len(launches)}', '']
        for k in KEYS:
            if k in r:
                lines.append(f'{k:95s} {r[k][0]:>16s} {r[k][1]}')
        f = lambda k: float(r[k][0].replace(',', ''))
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        dram = f('dram__bytes_read.sum') * mult[r['dram__bytes_read.sum'][1]] + \
            f('dram__bytes_write.sum') * mult[r['dram__bytes_write.sum'][1]]
        tu = {'ns': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}
        dur = f('gpu__time_duration.sum') * tu.get(r['gpu__time_duration.sum'][1], 1.0)
        summ[name] = {'round': tag, 'dram_bytes_per_launch': dram, 'duration_ms_under_ncu': dur,
                      'dmma_pipe_pct': f('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'),
                      'fp64_pipe_pct': f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'),
                      'issue_active_pct': f('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                      'registers': int(f('launch__registers_per_thread')),
                      'capture': os.path.basename(rep)}
        lines += ['', '# hottest source lines (samples, %, file:line, instructions, top stall reasons)']
        top = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_lines.py'), rep, '40'],
                             capture_output=True, text=True).stdout
        lines += [ln[:210] for ln in top.splitlines()]
        with open(os.path.join(ROOT, 'profiles', f'{tag}_{name}.txt'), 'w') as fh:
            fh.write('\n'.join(lines) + '\n')
        print('wrote', f'profiles/{tag}_{name}.txt')
    json.dump(summ, open(sj, 'w'), indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
