#!/usr/bin/env python
"""ncu launch list -> profiles/<tag>_launches.csv + profiles/ncu_traffic.json.

Input: the CSV log of
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_raw.csv python bench.py --groups 131 --ncu
(1048 pairs = 4 sub-batches of 262, the sub-batch size of the full 10,000-pair corpus, so per-launch figures carry
over).  The run holds PASSES identical passes (--ncu: 1 warm-up + 1; a plain `--steps 1 --warmup 3` run: 5); the LAST
pass is kept.

  python tools/launch_table.py gpurun_out/launches_raw.csv r2_final [pairs_in_run=1048] [pairs_full=10000] [passes=2]

Writes profiles/<tag>_launches.csv (one row per launch of one pass), prints the per-kernel share table, and records in
profiles/ncu_traffic.json: DRAM bytes per launch per kernel (keyed like bench.py's kernel table; `roofline.traffic`),
the DRAM bytes of one full-corpus step and the DRAM throughput averaged over the pass.  Per-launch times under ncu are
cold-cache and serialised: compare SHARES with bench.py's CUDA-event table, not absolutes."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw, tag = sys.argv[1], sys.argv[2]
pairs_run = int(sys.argv[3]) if len(sys.argv) > 3 else 1048
pairs_full = int(sys.argv[4]) if len(sys.argv) > 4 else 10000
PASSES = int(sys.argv[5]) if len(sys.argv) > 5 else 2

NAMES = {"k_ds_stream<0>": "k_ds_stats<ref>", "k_ds_stream<1>": "k_ds_stats<pair>",
         "k_ds_stream<0, 1>": "k_ds_stats<ref>", "k_ds_stream<0, 0>": "k_ds_stats<ref> edge", "k_ds_stream<1, 1>": "k_ds_stats<pair>",
         "k_ds_stream<1, 0>": "k_ds_stats<pair> edge", "k_ds_blur2<1>": "k_ds_blur2",
         "k_ds_blur2<0>": "k_ds_blur2", "k_ba_combine4": "k_ba_combine", "k_ba_mask4": "k_ba_mask", "k_ba_malta<1>": "k_ba_malta", "k_ba_malta<0>": "k_ba_malta",
         "k_ba_malta_diff<1>": "k_ba_malta_diff", "k_ba_malta_diff<0>": "k_ba_malta_diff", "k_ba_opsin<1, 1>": "k_ba_opsin",
         "k_ba_opsin<1, 0>": "k_ba_opsin"}
for t in ("0", "1"):
    for m, nm in (("0", ""), ("1", "<ref>"), ("2", "<pair>")):
        NAMES[f"k_s2_hpass<{m}, {t}>"] = "k_s2_hpass" + nm
        NAMES[f"k_s2_vpass<{m}, {t}>"] = "k_s2_vpass" + nm
    NAMES[f"k_ba_blur_h<0, 16, {t}>"] = "k_ba_blur_h<R16>"
    NAMES[f"k_ba_blur_v<0, 16, 3, 1, {t}>"] = "k_ba_blur_v<R16>+lf"
    NAMES[f"k_ba_blur2d<1, 7, 3, 2, {t}>"] = "k_ba_blur2d<R7>+hf_split"
    NAMES[f"k_ba_blur2d<2, 3, 2, 3, {t}>"] = "k_ba_blur2d<R3>+uhf_split"
    NAMES[f"k_ba_blur2d<3, 6, 1, 0, {t}>"] = "k_ba_blur2d<R6>"


def bench_name(k):
    k = re.sub(r"^void\s+", "", k).replace("ce::", "")
    base = k.split("(")[0].replace("(int)", "").replace("(bool)", "")
    m = re.match(r"k_ds_stream<(\d), \d, \d>", base)   # interior / edge / cp.async variants: one name, like bench.py
    if m:
        return "k_ds_stats<pair>" if m.group(1) == "1" else "k_ds_stats<ref>"
    return NAMES.get(base, base)


rows = [r for r in csv.reader(l for l in open(raw, errors="replace") if l.startswith('"'))]
hdr = rows[0]
col = {n: hdr.index(n) for n in ("ID", "Kernel Name", "Grid Size", "Block Size", "Metric Name", "Metric Unit", "Metric Value")}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}
launches = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= col["Metric Value"]:
        continue
    L = launches.setdefault(r[col["ID"]], {"kernel": bench_name(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
    v = float(r[col["Metric Value"]].replace(",", "")) * scale.get(r[col["Metric Unit"]], 1.0)
    L[r[col["Metric Name"]]] = v
ours = [L for L in launches.values() if L["kernel"].startswith("k_") and not L["kernel"].startswith("k_jpg") and "gpu__time_duration.sum" in L]
per_pass = len(ours) // PASSES
last = ours[-per_pass:]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
out_csv = os.path.join(ROOT, "profiles", f"{tag}_launches.csv")
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["#", "kernel", "grid", "block", "gpu__time_duration_us", "dram_read_MB", "dram_write_MB"])
    for i, L in enumerate(last):
        w.writerow([i, L["kernel"], L["grid"], L["block"], f"{L['gpu__time_duration.sum']:.2f}",
                    f"{L.get('dram__bytes_read.sum', 0) / 1e6:.2f}", f"{L.get('dram__bytes_write.sum', 0) / 1e6:.2f}"])
acc = collections.defaultdict(lambda: [0, 0.0, 0.0])
for L in last:
    a = acc[L["kernel"]]
    a[0] += 1
    a[1] += L["gpu__time_duration.sum"]
    a[2] += L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0)
tot_us = sum(a[1] for a in acc.values())
tot_b = sum(a[2] for a in acc.values())
print(f"{len(ours)} launches of this library in the run, {per_pass} per pass; last pass: {tot_us / 1e3:.2f} ms, {tot_b / 1e9:.2f} GB DRAM "
      f"-> {tot_b / tot_us / 1e3:.0f} GB/s averaged over the pass")
print(f"{'kernel':30s} {'launches':>8s} {'ms':>8s} {'share':>7s} {'DRAM MB/launch':>15s}")
for k, (n, us, b) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:30s} {n:8d} {us / 1e3:8.3f} {100 * us / tot_us:6.1f}% {b / n / 1e6:15.1f}")
peak = 6512.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
table = json.load(open(tpath)) if os.path.exists(tpath) else {}
table["cfg5"] = {k: {"bytes_per_launch": b / n, "launches_per_pass": n, "ms_per_pass_under_ncu": us / 1e3} for k, (n, us, b) in acc.items()}
table["_source"] = f"profiles/{tag}_launches.csv ({pairs_run} pairs = {pairs_run // 262} sub-batches of 262)"
table["_step_dram_gb_cfg5"] = tot_b / 1e9 * pairs_full / pairs_run
table["_step_dram_frac_cfg5"] = tot_b / tot_us / 1e3 / peak
table["_dram_bytes_per_pixel_pair"] = tot_b / (pairs_run * 1024 * 1024)
json.dump(table, open(tpath, "w"), indent=1, sort_keys=True)
print("wrote", out_csv, "and", tpath)
