#!/usr/bin/env python3
"""Summarise the hot SASS regions of an ncu report (source page): python tools/ncu_hot.py rep.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; ins = rows[2:]
tot = sum(int(r[ci['Instructions Executed']]) for r in ins); sam = sum(int(r[ci['# Samples']]) for r in ins)
print(rows[0][1][:100]); print('total warp-inst', tot, 'samples', sam, 'sass lines', len(ins))
seg = []; cur = None
for i, r in enumerate(ins):
    n = int(r[ci['Instructions Executed']]); s = int(r[ci['# Samples']]); t = int(r[ci['Thread Instructions Executed']])
    if cur is None or abs(n - cur['n']) > 0.25 * max(n, cur['n'], 1):
        cur = {'start': i, 'n': n, 'inst': 0, 'sam': 0, 'thr': 0, 'ops': {}}; seg.append(cur)
    cur['inst'] += n; cur['sam'] += s; cur['thr'] += t; cur['end'] = i
    op = r[1].strip().split()[0] if not r[1].strip().startswith('@') else r[1].strip().split()[1]
    op = op.split('.')[0]; cur['ops'][op] = cur['ops'].get(op, 0) + 1
for s in seg:
    if s['inst'] > minpct / 100 * tot:
        ops = sorted(s['ops'].items(), key=lambda kv: -kv[1])[:6]
        print(f"{s['start']:5d}-{s['end']:5d} x{s['n']:9d} inst {100*s['inst']/tot:5.1f}% samp {100*s['sam']/sam:5.1f}% lanes {s['thr']/max(s['inst'],1):5.1f}  {ops}")
