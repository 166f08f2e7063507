#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: instructions executed + stall samples.
usage: ncu_lines.py report.ncu-rep [function-substring] [top] [work-items]
(per-item counts are printed when work-items, e.g. the number of envs per launch, is given)"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ''
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
items = float(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
agg = collections.OrderedDict()
fname, func = None, ''
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
    elif r[0] == 'Function Name':
        func = r[1]
    elif r[0].isdigit() and len(r) > 7 and want in func:
        try:
            samples, inst = int(r[4]), int(r[7])
        except ValueError:
            continue
        a = agg.setdefault((fname, int(r[0])), [0, 0, r[1].strip()[:105]])
        a[0] += samples
        a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f'functions matching {want!r}: total samples {tot_s}, total warp-instr {tot_i}'
      + (f' = {tot_i / items:.0f} per item' if items else ''))
def line(a, f, ln):
    per = f'{a[1] / items:6.1f}/item ' if items else ''
    return f'{per}{100 * a[1] / tot_i:5.1f}% ins {100 * a[0] / tot_s:5.1f}% smp  {f}:{ln:<4d} {a[2]}'
print('--- by stall samples')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(line(a, f, ln))
print('--- by instructions')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(line(a, f, ln))
