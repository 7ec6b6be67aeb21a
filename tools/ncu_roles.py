#!/usr/bin/env python
"""Stall-sample digest of `ncu --page source --csv`: the hottest instructions and samples per bucket of the SASS listing
(for a warp-specialised kernel the buckets are the roles).  Usage: ncu_roles.py <source.csv> [top] [bucket]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 200
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
iS, iN, iE, iA = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Address')
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[iA] in seen:
        continue
    seen.add(r[iA])
    data.append(r)
tot = sum(int(r[iN]) for r in data)
print('total samples', tot, ' instructions', len(data), ' warp-instr executed', sum(int(r[iE]) for r in data))
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iN]))[:top]):
    r = data[i]
    print(f"{i:5d} {int(r[iN]):6d} {100 * int(r[iN]) / tot:5.1f}% {r[iE]:>9s}  {r[iS].strip()[:100]}")
print('---- markers')
for i, r in enumerate(data):
    if any(k in r[iS] for k in ('UBLKCP', 'UTMALDG', 'UTCIMMA', 'UTCBAR', 'LDTM', 'BAR.SYNC', 'CREDUX')):
        print(f"{i:5d} {int(r[iN]):6d} {r[iE]:>9s}  {r[iS].strip()[:80]}")
print('---- samples per bucket')
for b in range(0, len(data), bucket):
    n = sum(int(r[iN]) for r in data[b:b + bucket])
    if n:
        print(f"{b:5d} {n:6d} {100 * n / tot:5.1f}%")
