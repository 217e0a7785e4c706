"""Per engine phase: warp instructions executed and stall samples of an ncu report (source page).
python scripts/ncu_phase_insts.py report.ncu-rep [n_evals]"""
import csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
nev = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
src = open(os.path.join(ROOT, 'bayes_drt_b200', 'csrc', 'engine.cuh')).read().splitlines()
marks = []
for i, ln in enumerate(src, 1):
    mm = re.search(r'// -+ (phase \d[^\n]*)', ln)
    if mm: marks.append((i, mm.group(1)[:34]))
    if 'inline void engine_load' in ln: marks.append((i, 'engine_load'))
    if 'inline void engine_fill_table' in ln: marks.append((i, 'engine_load'))
    if 'inline double engine_eval' in ln: marks.append((i, 'eval prologue'))
def phase(line):
    lab = 'engine other'
    for l, name in marks:
        if line >= l: lab = name
    return lab
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-' and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: s = int(d['# Samples']); ins = int(d['Instructions Executed'])
        except ValueError: continue
        key = phase(int(r[0])) if fname == 'engine.cuh' else fname
        a = agg.setdefault(key, dict(s=0, i=0, st={}))
        a['s'] += s; a['i'] += ins
        for k in d:
            if k.startswith('stall_') and '(' not in k and d[k].isdigit():
                a['st'][k[6:]] = a['st'].get(k[6:], 0) + int(d[k])
ts = sum(a['s'] for a in agg.values()); ti = sum(a['i'] for a in agg.values())
print(f'total samples {ts}  warp instructions {ti}  per eval {ti/nev:.0f}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['s']):
    top = sorted(a['st'].items(), key=lambda kv: -kv[1])[:5]
    print(f"{100*a['s']/ts:5.1f}% samples {100*a['i']/ti:5.1f}% inst ({a['i']/nev:7.0f}/eval) {a['s']/max(a['i'],1)*ti/ts:5.2f} rel.cost  {k:36s} " +
          ' '.join(f'{x}={100*y/max(a["s"],1):.0f}%' for x, y in top))
