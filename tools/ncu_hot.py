"""Hot SASS lines of one kernel in an .ncu-rep (source page): python tools/ncu_hot.py rep [min_frac]"""
import csv, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]
si, sa, ex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
body = rows[2:]
tot = sum(float(r[sa] or 0) for r in body if len(r) > sa)
print('total samples', tot)
for n, r in enumerate(body):
    if len(r) <= sa:
        continue
    v = float(r[sa] or 0)
    if v >= tot * frac:
        print('%5d %6.2f%% exec %10s  %s' % (n, 100 * v / tot, r[ex], r[si].strip()[:110]))
