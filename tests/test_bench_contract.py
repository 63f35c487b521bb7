"""CPU test of bench.py's reference arm (`--impl reference`): the one bench leg that runs without a GPU.  It times the
CPU oracle port on a bounded sample of the same workload and must print exactly one JSON line with the contract's keys.
Under torchrun only rank 0 runs it; the other ranks exit 0 silently."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                          capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_contract_line():
    p = _run()
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["metric"] == "mpix_pairs_per_sec_all_metrics" and d["unit"] == "MPix-pairs/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["steps"] == 1 and d["warmup"] >= 3          # the timing rules ask for at least 3 warm-up steps
    assert "workload" in d["config"] and "model" not in d["config"]
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.corpus_config(1) and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # value and ms_per_step describe the same measurement: sample MPix / time
    assert d["ms_per_step"] > 0


def test_reference_arm_on_a_non_zero_rank_is_silent():
    p = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_committed_gpu_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (what DESIGN.md and profiles/README.md quote) have the full contract:
    roofline of the dominant kernel against the measured peak, cpu_baseline, e2e with real copies, launches, clocks."""
    for name, n in (("r1_v47_bench.json", 1), ("r1_v45_bench.json", 1), ("r1_v47_bench_2gpu.json", 2), ("r1_v47_bench_4gpu.json", 4)):
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["metric"] == "mpix_pairs_per_sec_all_metrics"
        assert d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] > 0 and d["warmup"] >= 3
        r = d["roofline"]
        assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        e = d["e2e"]
        assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
        assert "workload" in d["config"]
        if n == 1:
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1


def test_single_pair_latency_leg_with_a_stand_in_context():
    """The configs[0] leg of bench.py is host logic over three already-tested single-pair entries: run it against a
    recording stand-in (no GPU here) so a slip in it cannot surface for the first time on the GPU box."""
    sys.path.insert(0, ROOT)
    import bench

    class Ctx:
        def __init__(self):
            self.calls = []

        def calculate_psnr(self, r, t, w, h):
            self.calls.append(("psnr", r.shape, t.shape, w, h))
            return 30.0

        def calculate_dssim_rgb8(self, r, t, w, h):
            self.calls.append(("dssim", r.shape, t.shape, w, h))
            return 0.001

        def calculate_ssimulacra2(self, r, t, w, h):
            self.calls.append(("ssimulacra2", r.shape, t.shape, w, h))
            return 70.0

        def evaluate_batch(self, pairs, cfg):
            self.calls.append(("batch", len(pairs), cfg.psnr and cfg.dssim and cfg.ssimulacra2 and not cfg.butteraugli))
            return [None]

    c = Ctx()
    out = bench.single_pair_latency(c, reps=4)
    assert [k[0] for k in c.calls[:4]] == ["psnr", "dssim", "ssimulacra2", "batch"]      # the order of calculate_metrics
    assert len(c.calls) == 4 * (4 + 3) and c.calls[0][1:] == ((512, 512, 3), (512, 512, 3), 512, 512)
    assert c.calls[3] == ("batch", 1, True)
    assert out["reps"] == 4 and out["ms_psnr"] >= 0 and out["ms_dssim"] >= 0 and out["ms_ssimulacra2"] >= 0
    assert out["ms_per_pair"] >= 0 and out["ms_one_batch_call"] >= 0 and out["mpix_pairs_per_sec"] > 0 and "cfg1" in out["what"]


def test_both_arms_describe_the_same_config():
    """The driver compares the two arms' `config` dicts: they come from one function, and the corpus layout the GPU arm
    shards is the one the CPU arm samples (group-major, quality-minor; ragged shards at 8 ranks)."""
    sys.path.insert(0, ROOT)
    import bench

    c1, c8 = bench.corpus_config(1), bench.corpus_config(8)
    assert c1["pairs"] == 10000 and c1["width"] == c1["height"] == 1024 and "cfg5" in c1["workload"]
    assert {k: v for k, v in c1.items() if k != "parallelism"} == {k: v for k, v in c8.items() if k != "parallelism"}
    assert all(len(str(v)) <= 120 for v in c1.values())      # the driver's record truncates longer strings
    for world in (1, 2, 8):
        ref_ids, shards = bench.corpus_layout(world)
        assert sorted(i for s in shards for i in s) == list(range(10000))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 8
        for s in shards:                                       # a reference group never straddles two ranks
            assert len(s) % 8 == 0 and len({int(ref_ids[i]) for i in s}) == len(s) // 8
    assert len({len(s) for s in bench.corpus_layout(8)[1]}) == 2       # 1250 groups over 8 ranks: 157 / 156 -> ragged


def test_traffic_table_matches_the_benched_kernels():
    """`roofline.traffic` is looked up by kernel name in profiles/ncu_traffic.json (made by tools/launch_table.py from the
    ncu launch list): every kernel of the newest committed bench line must be in it, with the launch count the launch
    list implies for the full corpus, and the step-level figures must be the ones the line carries."""
    import glob

    t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    newest = sorted(glob.glob(os.path.join(ROOT, "profiles", "r2d_bench.json")) or glob.glob(os.path.join(ROOT, "profiles", "r2*_bench.json")))[-1]
    d = json.loads(open(newest).read().strip().splitlines()[-1])
    assert d["roofline"]["traffic_source"] == t["_source"]
    assert abs(d["roofline"]["step_dram_frac"] - t["_step_dram_frac_cfg5"]) < 1e-12
    for k, v in d["kernels"].items():
        assert k in t["cfg5"], k
        # the bench table counts launches over its profiled steps (2): 39 sub-batches x launches per sub-batch
        per_sub = t["cfg5"][k]["launches_per_pass"] / 4
        assert v["launches"] == 2 * 39 * per_sub, (k, v["launches"], per_sub)
    top = d["roofline"]["kernel"]
    assert d["roofline"]["traffic"] == t["cfg5"][top]["bytes_per_launch"]
    # DRAM traffic of the dominant kernel within 1.5x of its algorithmic bytes (no wasted re-reads)
    assert 0.8 < d["roofline"]["traffic"] / d["roofline"]["algorithmic_bytes_per_launch"] < 1.5


def test_committed_round2_lines_carry_the_corpus_contract():
    """The round-2 lines: BASELINE's 1/2/4/8 config strong-scaled, per-metric / cfg1-cfg4 / parity nested inside `roofline`
    (so the driver's record keeps them), both byte accountings, the library stamp, the reference arm on the same config."""
    lines = (("r2d_bench.json", 1), ("r2c_bench.json", 1), ("r2b_bench.json", 1), ("r2b_bench_2gpu.json", 2), ("r2d_bench_4gpu.json", 4))
    for name, n in lines:
        d = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        assert d["n_gpus"] == n and d["scaling"] == "strong" and d["metric"] == "mpix_pairs_per_sec_all_metrics"
        assert d["config"]["pairs"] == 10000 and "cfg5" in d["config"]["workload"] and d["dtype"] == "f32"
        assert d["steps"] >= 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and "src:" in d["lib"] and d["status_ok"]
        assert abs(d["value"] - 10000 * 1024 * 1024 / 1e6 / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-6
        e = d["e2e"]
        assert 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] == 10000 * (1 + 1 / 8) * 3 * 1024 * 1024
        assert e["results_identical_to_resident"] is True and e["pageable"]["value"] > 0
        r = d["roofline"]
        assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["frac_per_pair"] >= r["frac"]
        assert set(r["per_metric"]) == {"psnr", "ssimulacra2", "dssim", "butteraugli"}
        assert 0.3 < r["step_dram_frac"] < 1.0 and r["step_dram_gb"] > 0
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
        if n == 1:
            p = r["parity"]
            assert p["ok"] and p["checked"] >= 8 and p["sse_exact"] and p["max_abs_ssim2"] < p["tol"]["ssimulacra2_abs"]
            assert p["max_rel_dssim"] < p["tol"]["dssim_rel"] and p["max_rel_ba"] < p["tol"]["butteraugli_rel"]
            assert all(k in r for k in ("cfg1", "cfg2", "cfg3", "cfg4")) and r["cfg4"]["mpix_pairs_per_sec"] > 0
            assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    for name in ("r2d_reference.json", "r2c_reference.json"):
        ref = json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])
        ours = json.loads(open(os.path.join(ROOT, "profiles", name.replace("_reference", "_bench"))).read().strip().splitlines()[-1])
        assert ref["impl"] == "reference" and ref["config"] == ours["config"] and ref["metric"] == ours["metric"]
        assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["gpu_launches"] == 0 and ref["value"] == ref["cpu_baseline"]["value"]
