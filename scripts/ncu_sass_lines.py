#!/usr/bin/env python
"""Per-source-line instruction totals of one kernel: joins the SASS page of an ncu report with the line
table of the matching cubin (nvdisasm -g).  Works when ncu could not import the .cu source itself.
   python scripts/ncu_sass_lines.py rep kernel_regex cubin mangled_substring [top]"""
import csv, re, subprocess, sys, collections, os
rep, rx, cubin, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx], capture_output=True, text=True).stdout
hdr, rows, nk = None, [], 0
for r in csv.reader(out.splitlines()):
    if r and r[0] == 'Kernel Name':
        nk += 1
        if nk > 1: break
        continue
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): rows.append(r)
i_i, i_t, i_s = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
# walk the function's text: remember the current "//## File "...", line N" marker; instructions are lines with /*xxxx*/
lines, cur, infn = [], ('?', 0), False
for ln in dis.splitlines():
    if ln.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', ln):
        infn = mangled in ln
        continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r'/\*[0-9a-f]{4}\*/', ln): lines.append(cur)
print('sass rows', len(rows), 'disasm instr', len(lines))
n = min(len(rows), len(lines))
agg = collections.OrderedDict()
for k in range(n):
    a = agg.setdefault(lines[k], [0, 0, 0])
    a[0] += int(rows[k][i_i]); a[1] += int(rows[k][i_t]); a[2] += int(rows[k][i_s])
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
src = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2 if os.environ.get("BY_SAMPLES") else 0])[:top]:
    if f not in src:
        for d in ('multi-view-registration_b200/csrc', '.'):
            pth = os.path.join(d, f)
            if os.path.exists(pth): src[f] = open(pth).read().splitlines(); break
        else: src[f] = []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ''
    print('%-14s %4d inst %9d (%4.1f%%) thr/inst %4.1f smp %5d (%4.1f%%)  %s' % (f, l, a[0], 100.0 * a[0] / ti, a[1] / max(a[0], 1), a[2], 100.0 * a[2] / max(ts, 1), text))
