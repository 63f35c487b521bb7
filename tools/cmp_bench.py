#!/usr/bin/env python
"""Per-kernel ms/step of several bench lines side by side:  python tools/cmp_bench.py a_bench.json b_bench.json ..."""
import json, sys
def load(f):
    return json.loads(open(f).read().strip().splitlines()[-1])
ds = [load(f) for f in sys.argv[1:]]
ks = sorted(set().union(*[set(d['kernels']) for d in ds]), key=lambda k: -ds[-1]['kernels'].get(k, {'ms_per_step': 0})['ms_per_step'])
print(f"{'kernel':28s}" + "".join(f"{f.split('/')[-1].replace('_bench.json','')[:12]:>13s}" for f in sys.argv[1:]))
for k in ks:
    print(f"{k:28s}" + "".join(f"{d['kernels'].get(k, {}).get('ms_per_step', float('nan')):13.1f}" for d in ds))
print(f"{'step':28s}" + "".join(f"{d['ms_per_step']:13.1f}" for d in ds))
print(f"{'sm_mhz':28s}" + "".join(f"{d['clocks']['sm_mhz']:13.0f}" for d in ds))
