#!/usr/bin/env python
"""Per-region view of an ncu source page (ncu -i rep --page source --csv > file): consecutive SASS instructions with a
similar execution count form a region; prints each region's share of the stall samples and of the executed instructions,
and the instructions with the most samples.  Usage: python tools/ncu_regions.py source.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
ia, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
names = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_dispatch', 'stall_not_selected', 'stall_selected', 'stall_branch_resolving', 'stall_no_inst']
idx = {n: hdr.index(n) for n in names}
data = rows[2:]
tot = sum(int(r[isamp]) for r in data)
totex = sum(int(r[iex]) for r in data)
print('samples', tot, 'warp instructions', totex)
k = 0
while k < len(data):
    e = int(data[k][iex]); j = k; s = 0; n = 0; ex = 0
    while j < len(data) and (abs(int(data[j][iex]) - e) <= 0.3 * max(e, 1)):
        s += int(data[j][isamp]); ex += int(data[j][iex]); n += 1; j += 1
    if s > 0.004 * tot:
        print('lines %4d-%4d n=%4d exec/instr~%8d  samples %6d (%4.1f%%)  instr %4.1f%%  %s' % (k, j, n, e, s, 100 * s / tot, 100 * ex / totex, data[k][ia].strip()[:40]))
    k = j
print('--- totals by reason:', {n[6:]: sum(int(r[idx[n]]) for r in data) for n in names})
print('--- top instructions')
for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
    print(r[isamp], r[iex], ' '.join('%s=%s' % (n[6:9], r[idx[n]]) for n in names[:5]), r[ia].strip()[:90])
