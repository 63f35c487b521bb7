#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` report of bench.py: average DRAM bytes (read + write) per
launch of every profiled kernel, keyed by the names bench.py's per-kernel table uses.
  python tools/make_traffic.py gpurun_out/prof.ncu-rep cfg2 [profiles/ncu_traffic.json]"""
import csv, json, os, re, subprocess, sys, collections

rep, workload = sys.argv[1], sys.argv[2]
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
kn, rd, wr, tm = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

def bench_name(k):
    k = re.sub(r"^void\s+", "", k).replace("ce::", "")
    base = k.split("(")[0]
    base = base.replace("(int)", "").replace("(bool)", "")
    m = {"k_ds_stream<0>": "k_ds_stats<ref>", "k_ds_stream<1>": "k_ds_stats<pair>", "k_ds_blur2s": "k_ds_blur2",
         "k_s2_hpass<0>": "k_s2_hpass", "k_s2_hpass<1>": "k_s2_hpass<ref>", "k_s2_hpass<2>": "k_s2_hpass<pair>",
         "k_s2_vpass<0>": "k_s2_vpass", "k_s2_vpass<1>": "k_s2_vpass<ref>", "k_s2_vpass<2>": "k_s2_vpass<pair>",
         "k_ba_opsin<1, 1>": "k_ba_opsin", "k_ba_opsin<1, 0>": "k_ba_opsin", "k_ds_blur2<1>": "k_ds_blur2", "k_ds_blur2<0>": "k_ds_blur2",
         "k_ba_combine4": "k_ba_combine", "k_ba_malta<1>": "k_ba_malta", "k_ba_malta<0>": "k_ba_malta",
         "k_s2_hpass<0, 1>": "k_s2_hpass", "k_s2_hpass<1, 1>": "k_s2_hpass<ref>", "k_s2_hpass<2, 1>": "k_s2_hpass<pair>",
         "k_s2_hpass<0, 0>": "k_s2_hpass", "k_s2_hpass<1, 0>": "k_s2_hpass<ref>", "k_s2_hpass<2, 0>": "k_s2_hpass<pair>",
         "k_s2_vpass<0, 1>": "k_s2_vpass", "k_s2_vpass<1, 1>": "k_s2_vpass<ref>", "k_s2_vpass<2, 1>": "k_s2_vpass<pair>",
         "k_s2_vpass<0, 0>": "k_s2_vpass", "k_s2_vpass<1, 0>": "k_s2_vpass<ref>", "k_s2_vpass<2, 0>": "k_s2_vpass<pair>",
         "k_ba_blur_h<0, 16, 1>": "k_ba_blur_h<R16>", "k_ba_blur_h<0, 16, 0>": "k_ba_blur_h<R16>",
         "k_ba_blur_v<0, 16, 3, 1, 1>": "k_ba_blur_v<R16>+lf", "k_ba_blur_v<0, 16, 3, 1, 0>": "k_ba_blur_v<R16>+lf",
         "k_ba_blur2d<1, 7, 3, 2, 1>": "k_ba_blur2d<R7>+hf_split", "k_ba_blur2d<2, 3, 2, 3, 1>": "k_ba_blur2d<R3>+uhf_split",
         "k_ba_blur2d<3, 6, 1, 0, 1>": "k_ba_blur2d<R6>", "k_ba_blur2d<1, 7, 3, 2, 0>": "k_ba_blur2d<R7>+hf_split",
         "k_ba_blur2d<2, 3, 2, 3, 0>": "k_ba_blur2d<R3>+uhf_split", "k_ba_blur2d<3, 6, 1, 0, 0>": "k_ba_blur2d<R6>", "k_jpg_ycc<0>": "k_jpg_ycc", "k_jpg_ycc<2>": "k_jpg_ycc", "k_jpg_rgb<0>": "k_jpg_rgb", "k_jpg_rgb<2>": "k_jpg_rgb", "k_ba_blur_h<0, 16>": "k_ba_blur_h<R16>",
         "k_ba_blur_v<0, 16, 3, 1>": "k_ba_blur_v<R16>+lf", "k_ba_blur2d<1, 7, 3, 2>": "k_ba_blur2d<R7>+hf_split",
         "k_ba_blur2d<2, 3, 2, 3>": "k_ba_blur2d<R3>+uhf_split", "k_ba_blur2d<3, 6, 1, 0>": "k_ba_blur2d<R6>",
         "k_ba_opsin<1>": "k_ba_opsin", "k_ba_malta_diff<1>": "k_ba_malta_diff", "k_ba_malta_diff<0>": "k_ba_malta_diff"}
    return m.get(base, base)

acc = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[2:]:
    name = bench_name(r[kn])
    b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
    a = acc[name]
    a[0] += 1; a[1] += b; a[2] += float(r[tm])
table = {}
if os.path.exists(out_path):
    table = json.load(open(out_path))
table.setdefault(workload, {})
for name, (cnt, b, t) in acc.items():
    table[workload][name] = b / cnt
    print(f"{name:30s} launches {cnt:3d}  dram bytes/launch {b / cnt / 1e6:10.2f} MB")
json.dump(table, open(out_path, "w"), indent=1, sort_keys=True)
print("wrote", out_path)
