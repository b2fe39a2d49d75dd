#!/usr/bin/env python
"""Per-source-line executed warp instructions of the first kernel in an .ncu-rep, in file/line order,
with region sums.  usage: python scripts/ncu_regions.py rep.ncu-rep [file:lo-hi ...]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Line No']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
g = float(rr[2][rr[0].index('launch__grid_size')])
def I(x):
    try: return int(x)
    except: return 0
per = collections.OrderedDict()
seen_files = []
for n, h0 in enumerate(hi):
    if rows[h0 - 2][0] != 'File Path': continue
    fname = rows[h0 - 2][1].split('/')[-1]
    if fname in seen_files: break          # second kernel
    seen_files.append(fname)
    h = rows[h0]; ix = {nm: i for i, nm in enumerate(h)}
    end = hi[n + 1] - 2 if n + 1 < len(hi) else len(rows)
    for r in rows[h0 + 1:end]:
        if len(r) < 10 or r[0] == '': continue
        c = I(r[ix['Instructions Executed']])
        if c: per[(fname, int(r[0]))] = (c / g, I(r[ix['# Samples']]), r[1][:110])
tot = sum(v[0] for v in per.values())
print('total per CTA', round(tot, 1))
for (f, l), (c, s, t) in per.items():
    print(f'{f.replace("marlnav_","")}:{l:5d} {c:7.1f} smp={s:5d} {t}')
