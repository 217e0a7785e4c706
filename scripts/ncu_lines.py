"""Summarise an ncu report per CUDA source line: python scripts/ncu_lines.py report.ncu-rep [topN]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; data = []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if len(r) > 10 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-' and r[0].isdigit():   # a source-line row (address '-')
        d = dict(zip(hdr, r))
        try: s = int(r[hdr.index('# Samples')])
        except ValueError: continue
        data.append((s, fname, int(r[0]), r[1].strip(), d))
tot = sum(x[0] for x in data)
print('total samples', tot)
data.sort(key=lambda x: -x[0])
for s, f, ln, src, d in data[:top]:
    st = {k: int(d[k]) for k in d if k.startswith('stall_') and '(' not in k and d[k].isdigit() and int(d[k]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{s:7d} {100*s/tot:5.1f}% {f}:{ln:<4d} inst={d["Instructions Executed"]:>10s} {" ".join(f"{k[6:]}={v}" for k,v in st):40s} | {src[:90]}')
