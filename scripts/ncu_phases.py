"""Aggregate the per-line samples of an ncu report by engine phase: python scripts/ncu_phases.py report.ncu-rep"""
import csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
src = open(os.path.join(ROOT, 'bayes_drt_b200', 'csrc', 'engine.cuh')).read().splitlines()
marks = []  # (line, label)
for i, ln in enumerate(src, 1):
    mm = re.search(r'// -+ (phase \d[^\n]*)', ln)
    if mm: marks.append((i, mm.group(1)[:40]))
    if 'inline void engine_load' in ln: marks.append((i, 'engine_load'))
    if 'inline double engine_eval' in ln: marks.append((i, 'eval prologue'))
def phase(line):
    lab = 'engine other'
    for l, name in marks:
        if line >= l: lab = name
    return lab
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; agg = {}; stall = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-' and r[0].isdigit():
        try: s = int(r[hdr.index('# Samples')])
        except ValueError: continue
        key = phase(int(r[0])) if fname == 'engine.cuh' else fname
        agg[key] = agg.get(key, 0) + s
        d = dict(zip(hdr, r))
        st = stall.setdefault(key, {})
        for k in d:
            if k.startswith('stall_') and '(' not in k and d[k].isdigit():
                st[k[6:]] = st.get(k[6:], 0) + int(d[k])
tot = sum(agg.values())
print('total samples', tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    top = sorted(stall[k].items(), key=lambda kv: -kv[1])[:4]
    print(f'{v:8d} {100*v/tot:5.1f}%  {k:42s} ' + ' '.join(f'{a}={100*b/max(v,1):.0f}%' for a, b in top))
