#!/usr/bin/env python
"""Run one metric selection over a small synthetic device-resident batch (the command ncu wraps).
  python tools/prof_run.py --metrics butteraugli --pairs 48 --w 768 --h 512 --reps 2"""
import argparse, ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from codec_eval_b200 import _lib
from codec_eval_b200.metrics import GpuMetrics, MetricConfig
from codec_eval_b200.synth import G, cheap_distort

ap = argparse.ArgumentParser()
ap.add_argument("--metrics", default="psnr,dssim,ssimulacra2,butteraugli")
ap.add_argument("--pairs", type=int, default=48)
ap.add_argument("--w", type=int, default=768)
ap.add_argument("--h", type=int, default=512)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--per-ref", type=int, default=0, help="distortions per reference (0 = every pair has its own reference)")
a = ap.parse_args()
m = set(a.metrics.split(","))
cfg = MetricConfig(dssim="dssim" in m, ssimulacra2="ssimulacra2" in m, butteraugli="butteraugli" in m, psnr="psnr" in m)
nuniq = min(a.pairs, 6)
refs = [G(i, a.w, a.h) for i in range(nuniq)]
dists = [cheap_distort(refs[i], 55 + 7 * i, seed=i) for i in range(nuniq)]
R = np.stack([refs[i % nuniq] for i in range(a.pairs)])
D = np.stack([dists[i % nuniq] for i in range(a.pairs)])
d_ref, d_dist = torch.from_numpy(R).cuda(), torch.from_numpy(D).cuda()
ctx = GpuMetrics(0)
if a.profile:
    ctx.profile(True)
if a.per_ref:
    n_ref = (a.pairs + a.per_ref - 1) // a.per_ref
    U = np.stack([refs[i % nuniq] for i in range(n_ref)])
    ri = np.arange(a.pairs, dtype=np.uint32) // a.per_ref
    D = np.stack([dists[(i // a.per_ref) % nuniq] for i in range(a.pairs)])
    d_ref, d_dist = torch.from_numpy(U).cuda(), torch.from_numpy(D).cuda()
for _ in range(a.reps):
    if a.per_ref:
        out = ctx.evaluate_batch_device_grouped(d_ref.data_ptr(), n_ref, d_dist.data_ptr(), a.pairs, ri, a.w, a.h, cfg)
    else:
        out = ctx.evaluate_batch_device(d_ref.data_ptr(), d_dist.data_ptr(), a.pairs, a.w, a.h, cfg)
torch.cuda.synchronize()
print("ok", out[0].status, out[0].ssimulacra2, out[0].dssim, out[0].butteraugli, out[0].psnr, "launches", ctx.launch_count())
if a.profile:
    for k, v in sorted(ctx.profile_report().items(), key=lambda kv: -kv[1]["ms"]):
        print(f"{k:28s} {v['launches']:4d} {v['ms']/a.reps:8.3f} ms/rep {v['bytes']/v['ms']/1e6:8.1f} GB/s")
ctx.close()
