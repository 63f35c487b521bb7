#!/usr/bin/env python
"""Small end-to-end invocation for compute-sanitizer: odd and even sizes, grouped references, every metric."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codec_eval_b200.metrics import GpuMetrics, MetricConfig
from codec_eval_b200.synth import G, cheap_distort

with GpuMetrics(0, workspace_bytes=1 << 30) as m:
    for (w, h) in [(64, 48), (77, 35), (160, 96), (9, 33)]:
        ref = G(1, w, h)
        pairs = [(ref, cheap_distort(ref, q, seed=q), w, h) for q in (40, 80)] + [(G(2, w, h), cheap_distort(G(2, w, h), 60), w, h)]
        r = m.evaluate_batch(pairs, MetricConfig.all().with_xyb_roundtrip())
        print(w, h, [round(x.ssimulacra2, 3) for x in r], [round(x.butteraugli, 4) for x in r])
    a = np.random.default_rng(0).random((24, 40, 4), dtype=np.float32)
    print(m.calculate_dssim(a, np.clip(a + 0.01, 0, 1).astype(np.float32)))
print("done")
