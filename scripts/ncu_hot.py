#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: sample share per SASS chunk and per opcode.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; python scripts/ncu_hot.py src.csv [chunk]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hdr = next(r for r in rows if '# Samples' in r)
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[hdr.index('# Samples')].isdigit()]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
tot = sum(int(r[iS]) for r in data)
totE = sum(int(r[iE]) for r in data)
print('total samples', tot, 'total warp instr', totE, 'n sass', len(data))
for c in range(0, len(data), chunk):
    s = sum(int(r[iS]) for r in data[c:c + chunk])
    e = sum(int(r[iE]) for r in data[c:c + chunk])
    if s > tot * 0.01:
        print(f'sass[{c}:{c + chunk}] {100 * s / tot:.1f}% samples, {100 * e / totE:.1f}% instr')
ops = collections.Counter()
for r in data:
    t = r[iSrc].split()
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += int(r[iE])
print('opcode share of executed warp instr:', [(k, f'{100 * v / totE:.1f}') for k, v in ops.most_common(24)])
top = sorted(data, key=lambda r: -int(r[iS]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 0]
for r in top:
    print(r[iS], r[iE], r[iSrc])
