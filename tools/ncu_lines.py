#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line of one kernel, from an .ncu-rep captured with --import-source on
(read here, no GPU).  Optionally only one SASS opcode (e.g. where do the MOVs come from):

  python tools/ncu_lines.py rep kernel_substring [top] [opcode|-] [samples]     ('samples': rank lines by stall samples)
  python tools/ncu_lines.py gpurun_out/prof_v47_ds.ncu-rep "k_ds_stream<(int)1>" 25 MOV
"""
import collections
import csv
import subprocess
import sys

rep, fn_sub = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
opname = sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != '-' else None
by_samples = len(sys.argv) > 5 and sys.argv[5] == 'samples'
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
cur_file = hdr = cur_line = None
active = False
first_fn = None
sel, tot = collections.Counter(), collections.Counter()
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        # the report lists every profiled launch; keep the first launch of the first matching function
        if fn_sub in r[1] and first_fn in (None, r[1]):
            first_fn = r[1]
            active = True
        else:
            active = False
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if not active or hdr is None:
        continue
    if r[0] not in ('', '-'):
        cur_line = (cur_file, r[0], r[1].strip()[:100])
        continue
    sass = r[3].strip()
    if sass in ('...', '-', ''):
        continue
    try:
        n = int(r[hdr.index('# Samples' if by_samples else 'Instructions Executed')])
    except ValueError:
        continue
    t = sass.split()
    op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
    tot[cur_line] += n
    if opname is None or op == opname:
        sel[cur_line] += n
T, A = sum(tot.values()), sum(sel.values())
print(first_fn)
print(('stall samples ' if by_samples else 'warp-instructions ') + f'{T}' + (f'; {opname} {A} = {100 * A / max(T, 1):.1f} %' if opname else ''))
for k, v in sel.most_common(top):
    print(f'{100 * v / max(T, 1):5.2f}%  {k[0]}:{k[1]}  {k[2]}')
