#!/usr/bin/env python
"""Executed warp-instructions per CUDA source line for one kernel: python tools/ncu_lines.py rep kernel_regex [top]"""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name', 'regex:' + sys.argv[2]],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Line No' in r and 'Instructions Executed' in r]
lines = []
for t, start in enumerate(hi):
    h = rows[start]
    ie, ln, src, smp = h.index('Instructions Executed'), h.index('Line No'), h.index('Source'), h.index('# Samples')
    end = hi[t + 1] if t + 1 < len(hi) else len(rows)
    for r in rows[start + 1:end]:
        if len(r) > ie and r[ln] not in ('', '-'):
            try: lines.append((int(r[ie]), int(r[smp]) if r[smp].isdigit() else 0, f"{t}:{r[ln]}", r[src].strip()))
            except ValueError: pass
tot = sum(x[0] for x in lines); ts = sum(x[1] for x in lines)
print('total warp-instr (this launch)', tot)
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for n, s, l, t in sorted(lines, reverse=True)[:top]:
    print(f'{100*n/tot:5.1f}% instr {100*s/max(ts,1):5.1f}% samples  L{l:>6s}  {t[:120]}')
