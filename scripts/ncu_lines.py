#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by source line: share of executed warp
instructions and of stall samples. Usage: ncu_lines.py dump.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and cur and r[0]:
        try:
            ln = int(r[0]); ie = int(r[hdr.index('Instructions Executed')]); smp = int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        if ie or smp:
            a = agg.setdefault((cur, ln), [0, 0, r[1].strip()[:100]])
            a[0] += ie; a[1] += smp
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print('total warp instructions', tot, 'stall samples', tots)
byfile = {}
for (f, l), (ie, smp, s) in agg.items():
    b = byfile.setdefault(f, [0, 0]); b[0] += ie; b[1] += smp
for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print("%-16s inst %.3f samples %.3f" % (f, v[0] / tot, v[1] / tots))
for (f, l), (ie, smp, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%s:%-4d inst=%.3f samp=%.3f  %s" % (f, l, ie / tot, smp / tots, s))
