#!/usr/bin/env python
"""Dynamic instruction mix / hot instructions of one kernel from `ncu -i rep --page source --csv --kernel-id :::N`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
print(rows[0][1][:120])
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("Instructions Executed")].isdigit() and r[0].startswith("0x")]
iS, iE, iSm = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
tot = sum(int(r[iE]) for r in data)
tots = sum(int(r[iSm]) for r in data)
print('total warp instr', tot, 'static', len(data), 'samples', tots)
mix = collections.Counter()
for r in data:
    toks = r[iS].strip().split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    mix[op.split('.')[0]] += int(r[iE])
for k, v in mix.most_common(28):
    print(f"  {k:10s} {v:12d} {100 * v / tot:5.1f}%")
if thr:
    print('---- instructions with >= %.1f%% of samples' % thr)
    for r in data:
        if tots and 100 * int(r[iSm]) / tots >= thr:
            print(f"  {100 * int(r[iSm]) / tots:5.1f}%  exec={r[iE]:>10s}  {r[iS].strip()[:100]}")
