#!/usr/bin/env python
"""Repeat one evaluation many times and report any run whose results differ bitwise from the first (race detector).
  python tools/stress_determinism.py [w h pairs reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from codec_eval_b200.metrics import GpuMetrics, MetricConfig
from codec_eval_b200.synth import G, cheap_distort

w, h, npairs, reps = (int(v) for v in (sys.argv[1:5] + ["3840", "2160", "2", "12"][len(sys.argv) - 1:]))
refs = [G(9 + i, w, h) for i in range(max(1, npairs // 2))]
pairs = [(refs[i % len(refs)], cheap_distort(refs[i % len(refs)], 50 + (7 * i) % 45, seed=i), w, h) for i in range(npairs)]
cfg = MetricConfig(butteraugli=True) if os.environ.get('CE_STRESS_BA') else MetricConfig.all()
ctx = GpuMetrics(0, workspace_bytes=6 << 30)
first, bad = None, 0
for r in range(reps):
    out = ctx.evaluate_batch_raw(pairs, cfg)
    cur = [(o.status, o.sse, o.dssim, o.ssimulacra2, o.butteraugli, o.butteraugli_pnorm3) for o in out[:npairs]]
    if first is None:
        first = cur
    elif cur != first:
        bad += 1
        for i, (a, b) in enumerate(zip(first, cur)):
            if a != b:
                print(f"rep {r} pair {i}: {a} != {b}")
print("reps", reps, "mismatching runs", bad)
sys.exit(1 if bad else 0)
