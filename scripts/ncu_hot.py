#!/usr/bin/env python
"""Hot SASS instructions of one kernel from an ncu report: python scripts/ncu_hot.py rep kernel_regex [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
lines = out.splitlines()
# first kernel instance only
rows = []
hdr = None
n_k = 0
for r in csv.reader(lines):
    if r and r[0] == 'Kernel Name':
        n_k += 1
        if n_k > 1: break
        continue
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr and len(r) == len(hdr): rows.append(r)
ix = {k: hdr.index(k) for k in ('Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_branch_resolving', 'stall_math', 'stall_not_selected','stall_selected')}
tot_s = sum(int(r[ix['# Samples']]) for r in rows); tot_i = sum(int(r[ix['Instructions Executed']]) for r in rows)
print('instructions', len(rows), 'samples', tot_s, 'warp-instr', tot_i)
# cumulative position so we can see regions
for k, r in enumerate(rows): r.append(k)
rows2 = sorted(rows, key=lambda r: -int(r[ix['# Samples']]))[:top]
for r in sorted(rows2, key=lambda r: r[-1]):
    print('%5d %-60s smp %6d (%4.1f%%) inst %8d thr/inst %4.1f  long_sb %5s wait %5s short %5s br %5s math %5s' % (
        r[-1], r[ix['Source']].strip()[:60], int(r[ix['# Samples']]), 100.0 * int(r[ix['# Samples']]) / max(tot_s, 1), int(r[ix['Instructions Executed']]),
        int(r[ix['Thread Instructions Executed']]) / max(int(r[ix['Instructions Executed']]), 1),
        r[ix['stall_long_sb']], r[ix['stall_wait']], r[ix['stall_short_sb']], r[ix['stall_branch_resolving']], r[ix['stall_math']]))
