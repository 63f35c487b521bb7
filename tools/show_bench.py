import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ["value", "ms_per_step", "gpu_launches", "clocks", "e2e", "cpu_baseline", "sanity"]})
for m, v in d.get("per_metric", {}).items():
    print(f"{m:12s} {v['mpix_pairs_per_sec']:10.1f} MPix-pairs/s  {v['ms_per_step']:8.3f} ms/step  staged {v['staged_model_gbs_per_gpu']:8.1f} GB/s = {v['frac_of_hbm_staged_model']:.3f} of HBM")
print(json.dumps(d["roofline"]))
for k, v in d["kernels"].items():
    print(f"{k:26s} {v['launches']:5d} {v['ms_per_step']:9.3f} ms/step {v['gbs'] or 0:9.1f} GB/s {v['frac_of_peak'] or 0:.3f}")
