#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: instructions executed + stall samples.
usage: ncu_lines.py report.ncu-rep [kernel-index] [top]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
files = {}
cur = None
func_seen = 0
want = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kern = -1
agg = collections.OrderedDict()
fname = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Kernel Name':
        kern += 1
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if r[0] and r[0].isdigit() and len(r) > 7:
        if kern not in (want, -1) and kern != want:
            continue
        key = (fname, int(r[0]))
        try:
            samples = int(r[4]); inst = int(r[7])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += samples; a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print(f'total samples {tot_s} total warp-instr {tot_i}')
print('--- by stall samples')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f'{100*a[0]/tot_s:5.1f}% smp {100*a[1]/tot_i:5.1f}% ins  {f}:{ln:<4d} {a[2]}')
print('--- by instructions')
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f'{100*a[1]/tot_i:5.1f}% ins {100*a[0]/tot_s:5.1f}% smp  {f}:{ln:<4d} {a[2]}')
