"""ptxas -v statistics of bayes_drt_b200/csrc/build.log as a table: registers, stack, spill stores / loads per kernel (and per
non-inlined device function).    python scripts/ptxas_summary.py [filter]"""
import re
import subprocess
import sys

log = open('bayes_drt_b200/csrc/build.log').read().split('\n')
flt = sys.argv[1] if len(sys.argv) > 1 else ''
rows = []
name, kind = None, None
for i, l in enumerate(log):
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        name, kind = m.group(1), 'kernel'
    m2 = re.search(r"Function properties for (\S+)", l)
    if m2:
        fn = m2.group(1)
        sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", log[i + 1])
        reg = None
        for j in range(i + 1, min(i + 4, len(log))):
            r = re.search(r"Used (\d+) registers", log[j])
            if r:
                reg = int(r.group(1))
                break
        rows.append((fn, reg, *(int(x) for x in sp.groups())))
names = subprocess.run(['c++filt'] + [r[0] for r in rows], capture_output=True, text=True).stdout.split('\n')
for (fn, reg, st, ss, sl), dn in zip(rows, names):
    dn = re.sub(r'\(.*', '', dn).replace('void ', '')
    if flt in dn:
        print(f'{str(reg):>5} regs {st:5d} stack {ss:5d} spill-st {sl:5d} spill-ld  {dn}')
