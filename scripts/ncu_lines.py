#!/usr/bin/env python
"""Per-CUDA-source-line totals of one kernel from an ncu report (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys, os
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = []
fname, func, hdr, first_func = None, None, None, None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = os.path.basename(r[1]); continue
    if r[0] == 'Function Name':
        func = r[1]
        if first_func is None:
            first_func = func
        continue
    if r[0] == 'Line No':
        hdr = r; continue
    if hdr is None or func != first_func or len(r) != len(hdr) or not r[0]:
        continue
    rows.append((fname, r))
i_s, i_i, i_t = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
tot_s = sum(int(r[i_s]) for _, r in rows); tot_i = sum(int(r[i_i]) for _, r in rows)
print(first_func, ': samples', tot_s, 'warp-instr', tot_i)
rows.sort(key=lambda fr: -int(fr[1][i_i]))
for f, r in rows[:top]:
    print('%-16s %4s  inst %9d (%4.1f%%)  thr/inst %4.1f  smp %5d (%4.1f%%)  %s' % (
        f, r[0], int(r[i_i]), 100.0 * int(r[i_i]) / max(tot_i, 1), int(r[i_t]) / max(int(r[i_i]), 1), int(r[i_s]),
        100.0 * int(r[i_s]) / max(tot_s, 1), r[1].strip()[:110]))
