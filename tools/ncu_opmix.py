#!/usr/bin/env python
"""Executed-instruction opcode mix per kernel from an .ncu-rep (SASS page; works without --import-source):
  python tools/ncu_opmix.py rep kernel_substring [top]"""
import csv, subprocess, sys
from collections import Counter

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
cur, hdr, per, smp = None, None, {}, {}
idx = 0
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] in ('Kernel Name', 'Function Name'):
        idx += 1
        cur = f"{idx}:{r[1]}"
        continue
    if r and 'Instructions Executed' in r and 'Source' in r:
        hdr = r
        continue
    if cur is None or hdr is None or len(r) < len(hdr) - 5:
        continue
    try:
        n = int(r[hdr.index('Instructions Executed')]); s = int(r[hdr.index('# Samples')] or 0)
    except ValueError:
        continue
    t = r[hdr.index('Source')].strip()
    op = (t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0].rstrip(';')
    per.setdefault(cur, Counter())[op] += n
    smp.setdefault(cur, Counter())[op] += s
want = sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
for k, c in per.items():
    if want in k:
        tot, ts = sum(c.values()), max(sum(smp[k].values()), 1)
        print(k[:100], 'warp-instr', tot)
        for op, n in c.most_common(top):
            print(f'  {op:10s} {100 * n / tot:5.1f}% instr  {100 * smp[k][op] / ts:5.1f}% samples')
