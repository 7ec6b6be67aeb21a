#!/usr/bin/env python
"""Per-region (role) instruction counts by issue pipe from `ncu --page source --csv`, in thread-instructions per value.
Usage: ncu_pipes.py <source.csv> <values> <boundary,boundary,...>   (boundaries = SASS instruction indices)"""
import collections
import csv
import sys

ALU = ('LOP3', 'PRMT', 'SHF', 'IADD3', 'ISETP', 'FSETP', 'SEL', 'FSEL', 'FMNMX', 'VIADD', 'LEA', 'PLOP3', 'MOV', 'I2FP', 'IABS', 'VIMNMX', 'FLO', 'BREV', 'POPC', 'CS2R', 'S2R', 'IADD', 'LOP', 'R2P', 'P2R', 'FCHK', 'SGXT', 'BMSK', 'VABSDIFF', 'F2FP')
FMA = ('IMAD', 'FFMA', 'FMUL', 'FADD', 'HFMA2', 'HADD2', 'HMUL2', 'IDP', 'FHADD')
LSU = ('LDS', 'STS', 'LDG', 'STG', 'LD', 'ST', 'ATOM', 'RED', 'LDSM', 'LDC', 'LDL', 'STL')
XU = ('MUFU', 'F2I', 'I2F', 'F2F', 'CREDUX', 'REDUX')
rows = list(csv.reader(open(sys.argv[1])))
values = float(sys.argv[2])
bounds = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else []
hi = next(i for i, r in enumerate(rows) if 'Instructions Executed' in r)
hdr = rows[hi]
iS, iE, iA = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Address')
seen, data = set(), []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[iA] in seen:
        continue
    seen.add(r[iA])
    data.append(r)
bounds = [0] + bounds + [len(data)]


def pipe(op):
    base = op.split('.')[0]
    for name, group in (('alu', ALU), ('fma', FMA), ('lsu', LSU), ('xu', XU)):
        if base in group:
            return name
    if base.startswith('BRA') or base in ('BSYNC', 'BSSY', 'EXIT', 'WARPSYNC', 'YIELD', 'NOP', 'BAR', 'CALL', 'RET', 'BPT'):
        return 'ctl'
    return 'other:' + base


tot_all = collections.Counter()
for a, b in zip(bounds[:-1], bounds[1:]):
    c = collections.Counter()
    others = collections.Counter()
    for r in data[a:b]:
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).rstrip(';')
        p = pipe(op)
        n = int(r[iE])
        if p.startswith('other:'):
            others[p[6:]] += n
            p = 'other'
        c[p] += n
    tot = sum(c.values())
    tot_all.update(c)
    print(f"[{a:5d},{b:5d}) total {tot * 32 / values:6.3f}/value  " + '  '.join(f"{k}={v * 32 / values:.3f}" for k, v in c.most_common()) +
          '   other: ' + ' '.join(f"{k}:{v * 32 / values:.3f}" for k, v in others.most_common(5)))
print('all: ' + '  '.join(f"{k}={v * 32 / values:.3f}" for k, v in tot_all.most_common()), ' total', f"{sum(tot_all.values()) * 32 / values:.3f}")
