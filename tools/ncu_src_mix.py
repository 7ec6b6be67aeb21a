#!/usr/bin/env python
"""Opcode mix of one kernel from `ncu --page source --csv` output: per-opcode dynamic warp-instruction counts and
instructions per processed value.   Usage: ncu_src_mix.py <source.csv> <values> [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
values = float(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
hdr_i = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hdr_i]
iS, iE = hdr.index('Source'), hdr.index('Instructions Executed')
iA = hdr.index('Address') if 'Address' in hdr else 0
seen, data = set(), []
for r in rows[hdr_i + 1:]:
    if len(r) != len(hdr) or not r[iE].isdigit():
        continue
    if r[iA] in seen:      # the CSV repeats the listing per view
        continue
    seen.add(r[iA])
    data.append(r)
tot = sum(int(r[iE]) for r in data)
mix = collections.Counter()
for r in data:
    t = r[iS].split()
    op = t[1] if t[0].startswith('@') else t[0]
    mix[op.rstrip(';')] += int(r[iE])
print(f"total warp instr {tot}  static {len(data)}" + (f"  thread-instr/value {tot * 32 / values:.2f}" if values else ""))
for k, v in mix.most_common(top):
    print(f"  {k:28s} {v:12d} {100 * v / tot:5.1f}%" + (f"  {v * 32 / values:6.3f}/value" if values else ""))
