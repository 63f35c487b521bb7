#!/usr/bin/env python
"""Dynamic opcode mix of one kernel from an .ncu-rep source page: python tools/ncu_opmix.py rep kernel_regex"""
import csv, subprocess, sys, collections
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--kernel-name', 'regex:' + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and 'Instructions Executed' in r]
seg = rows[hi[0] + 1: hi[1] - 1] if len(hi) > 1 else rows[hi[0] + 1:]
h = rows[hi[0]]
ie, src, smp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
mix, tot, samp = collections.Counter(), 0, collections.Counter()
for r in seg:
    if len(r) <= ie: continue
    try: n = int(r[ie])
    except ValueError: continue
    toks = r[src].split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0].rstrip(';')
    mix[op] += n; tot += n
    try: samp[op] += int(r[smp])
    except ValueError: pass
print('total warp-instr', tot)
ts = sum(samp.values())
for op, n in mix.most_common(25):
    print(f'{op:10s} {n:12d} {100*n/tot:5.1f}%   samples {100*samp[op]/max(ts,1):5.1f}%')
