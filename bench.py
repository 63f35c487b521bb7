#!/usr/bin/env python
"""bench.py -- megapixel-pairs/second of the metric hot path (BASELINE.json `metric`).

Headline workload = BASELINE.json configs[4], the config the 1/2/4/8-GPU metric is quoted on: the corpus sweep,
10,000 synthetic 1024x1024 pairs (1250 references x 8 JPEG quality levels; the references tile a pool of 64
distinct synthetic images = 512 distinct pairs), all four metrics, STRONG-scaled: the corpus is
partitioned over the ranks by reference group (codec_eval_b200.shard.partition_pairs), every rank evaluates its
shard on its own GPU, and one NCCL all_gather of the 56-byte result rows gives every rank the full table -- inside the
timed region.  A step = one pass over the whole corpus.  At N = 1 the same corpus runs on one GPU.

  value : shards resident in HBM (shard.evaluate_sharded_resident -> ce_evaluate_batch_device_grouped)
  e2e   : the same corpus from HOST buffers through ce_evaluate_batch (shard.evaluate_sharded_table), host->device
          copies and the result read-back inside the timed region

Secondary measurements (nested under `roofline` / `config` so the driver's record keeps them): per-metric throughput
of the same corpus at every N; on rank 0 at N = 1 also cfg2 (Kodak-shaped 192-pair batch, with the per-kernel table),
cfg1 (single 512x512 pair latency), cfg3 (SSIMULACRA2 only through the reference handle), cfg4 (256 x 4K pairs,
Butteraugli + DSSIM), end to end from pageable memory, and a parity check of pairs of the TIMED batch against the
oracle at the contract tolerances.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA kernels via the C ABI)
  python bench.py --impl reference ...                           the CPU arm: the oracle (port of the reference's CPU
        metric path; the Rust crates cannot be built here) on all host cores, bounded sample of the same corpus.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

QUALS = [50, 60, 70, 75, 80, 85, 90, 95]
SUBSAMPLING = [2, 2, 2, 0, 0, 0, 2, 2]            # per quality level: 4:2:0 / 4:4:4 (the reference sweeps both)
CORPUS = {"width": 1024, "height": 1024, "pool_refs": 64, "groups": 1250}    # 1250 x 8 = 10,000 pairs
STAGED_BYTES_PER_PX = {"psnr": 6.0, "ssimulacra2": 326.0, "dssim": 225.0, "butteraugli": 970.0}
TOL = {"ssimulacra2_abs": 0.01, "dssim_rel": 1e-4, "butteraugli_rel": 1e-3}                     # north_star contract
METRICS_DESC = "psnr+dssim+ssimulacra2+butteraugli (max + 3-norm)"


def corpus_config(world: int) -> dict:
    """The `config` dict: identical in both arms (the driver compares them)."""
    c = CORPUS
    n = c["groups"] * len(QUALS)
    return {
        "workload": f"cfg5 corpus sweep: {n} synthetic {c['width']}x{c['height']} pairs, all four metrics, strong-scaled by pair",
        "pairs": n, "width": c["width"], "height": c["height"], "metrics": METRICS_DESC,
        "references": f"{c['groups']} reference groups x {len(QUALS)} JPEG qualities (q50-95, 4:2:0 and 4:4:4)",
        "pool": f"{c['pool_refs']} distinct references tiled = {c['pool_refs'] * len(QUALS)} distinct pairs (SURVEY 8d)",
        "l2": "no flush: every step streams 35 GB of distinct input addresses, far beyond the 126 MB L2",
        "parallelism": f"pairs sharded over {world} rank(s) by reference group; NCCL all_gather of 56 B/pair result rows",
    }


# ----------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """start of the timed region: samples taken before it (warm-up, also under load) are only a fallback"""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        t_mark = getattr(self, "t_mark", 0.0)
        inside = [r for t, r in self.rows if t >= t_mark]
        rows = inside if len(inside) >= 2 else [r for _, r in self.rows]   # short timed regions: include the warm-up samples
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: everything else (NCCL's version banner, library chatter) is sent to
    stderr by pointing fd 1 at fd 2; emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def pool_references(n: int, w: int, h: int) -> np.ndarray:
    """The pool of distinct synthetic references [n, h, w, 3] (seeds 0..n-1), cached under /tmp."""
    from codec_eval_b200.synth import G_many

    cache = f"/tmp/ce_bench3_pool_{n}_{w}x{h}.npy"
    if os.path.exists(cache):
        try:
            a = np.load(cache)
            if a.shape == (n, h, w, 3):
                return a
        except Exception:
            pass
    a = G_many(list(range(n)), w, h)
    try:
        tmp = f"{cache}.{os.getpid()}.tmp.npy"
        np.save(tmp, a)
        os.replace(tmp, cache)
    except Exception:
        pass
    return a


def device_distortions(ctx, d_refs, quals=QUALS, subs=SUBSAMPLING):
    """d_refs: uint8 cuda tensor [R, h, w, 3] -> [R, len(quals), h, w, 3]: every reference at every quality level,
    generated by the on-device baseline-JPEG round trip (bit-exact with libjpeg-turbo, tests/test_jpeg_source.py)."""
    import torch

    R, h, w, _ = d_refs.shape
    out = torch.empty((R, len(quals), h, w, 3), dtype=torch.uint8, device=d_refs.device)
    for ss in sorted(set(subs)):
        ks = [k for k, s in enumerate(subs) if s == ss]
        tmp = torch.empty((R, len(ks), h, w, 3), dtype=torch.uint8, device=d_refs.device)
        ctx.jpeg_roundtrip_device(d_refs.data_ptr(), R, w, h, [quals[k] for k in ks], ss, tmp.data_ptr())
        for j, k in enumerate(ks):
            out[:, k] = tmp[:, j]
        del tmp
    return out


def corpus_layout(world: int):
    """-> (ref_ids [n], shards): pair p = (group p // 8, quality p % 8); group g uses pool reference g % pool_refs."""
    from codec_eval_b200.shard import partition_pairs

    nq = len(QUALS)
    n = CORPUS["groups"] * nq
    ref_ids = np.arange(n) // nq
    shards = partition_pairs(ref_ids.tolist(), [CORPUS["width"] * CORPUS["height"]] * n, world)
    return ref_ids, shards


# ----------------------------------------------------------------------------- the CPU arm
def cpu_sample(pool, n_pairs: int):
    """The first n_pairs pairs of the corpus (group-major, quality-minor) as host arrays: refs, dists [n, h, w, 3].
    Distortions from the C oracle's JPEG source (bit-exact with libjpeg-turbo and with the CUDA source)."""
    from oracle import oracle as O

    w, h = CORPUS["width"], CORPUS["height"]
    refs, dists = [], []
    for p in range(n_pairs):
        g, k = divmod(p, len(QUALS))
        ref = pool[g % pool.shape[0]]
        refs.append(ref)
        dists.append(O.jpeg_roundtrip(ref, w, h, QUALS[k], SUBSAMPLING[k]))
    return np.stack(refs), np.stack(dists)


def run_reference(args):
    """--impl reference: the reference's CPU metric path = the oracle port (the Rust crates cannot be built here), all
    host threads, each step a bounded sample of the same corpus.  Like calculate_metrics (src/eval/session.rs:437-497)
    it recomputes the reference side for every pair: the reference has no reuse on this path."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    w, h = CORPUS["width"], CORPUS["height"]
    cores = host_cores()   # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every core it is allowed to run on
    sample = max(8, min(cores, 32))
    pool = pool_references(min(CORPUS["pool_refs"], (sample + len(QUALS) - 1) // len(QUALS)), w, h)
    refs, dists = cpu_sample(pool, sample)
    for _ in range(args.warmup):
        O.evaluate_batch(refs, dists, w, h, 15, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.evaluate_batch(refs, dists, w, h, 15, threads=cores)
    dt = time.perf_counter() - t0
    val = sample * w * h / 1e6 * args.steps / dt
    what = f"{sample} of {CORPUS['groups'] * len(QUALS)} pairs per step, OpenMP over pairs, no reference-side reuse"
    line = {
        "impl": "reference", "metric": "mpix_pairs_per_sec_all_metrics", "value": val, "unit": "MPix-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": corpus_config(world),
        "cpu_baseline": {"value": val, "unit": "MPix-pairs/s", "cores": cores, "kind": "port", "sample": what},
        "e2e": {"value": val, "unit": "MPix-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------- secondary workloads (rank 0, N = 1)
def single_pair_latency(ctx, reps: int = 20):
    """BASELINE.json configs[0]: ONE 512x512 pair (G(0), JPEG q80 4:2:0) through the single-pair entries, in the order
    EvalSession::calculate_metrics calls them (src/eval/session.rs:437-497: PSNR, DSSIM, SSIMULACRA2), from pageable
    host buffers, synchronous -- wall-clock latency per call, median of `reps`."""
    from codec_eval_b200.metrics import MetricConfig
    from codec_eval_b200.synth import G, J

    w = h = 512
    ref = G(0, w, h)
    dist = J(ref, 80, 2)
    calls = [("psnr", lambda: ctx.calculate_psnr(ref, dist, w, h)), ("dssim", lambda: ctx.calculate_dssim_rgb8(ref, dist, w, h)),
             ("ssimulacra2", lambda: ctx.calculate_ssimulacra2(ref, dist, w, h))]
    times = {k: [] for k, _ in calls}
    total, fused = [], []
    cfg = MetricConfig(psnr=True, dssim=True, ssimulacra2=True)
    for it in range(reps + 3):
        t_all = 0.0
        for k, fn in calls:
            t0 = time.perf_counter()
            fn()
            dt = (time.perf_counter() - t0) * 1e3
            t_all += dt
            if it >= 3:
                times[k].append(dt)
        t0 = time.perf_counter()
        ctx.evaluate_batch([(ref, dist, w, h)], cfg)      # the three metrics in ONE call (metrics forked over streams)
        dt = (time.perf_counter() - t0) * 1e3
        if it >= 3:
            total.append(t_all)
            fused.append(dt)
    med = lambda v: float(sorted(v)[len(v) // 2])
    return {"what": "cfg1: 1 pair 512x512, PSNR+DSSIM+SSIMULACRA2, single-pair C-ABI entries, pageable host buffers",
            "ms_per_pair": med(total), "ms_psnr": med(times["psnr"]), "ms_dssim": med(times["dssim"]),
            "ms_ssimulacra2": med(times["ssimulacra2"]), "ms_one_batch_call": med(fused), "reps": reps,
            "mpix_pairs_per_sec": w * h / 1e6 / (med(total) / 1e3)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=0, help="reference groups of the corpus (default 1250 = 10,000 pairs); profiling runs shrink it")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip cfg1-cfg4, per-metric and pageable legs")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (profiling runs only: the line then has e2e = null)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel tables (JSON) here")
    ap.add_argument("--ncu", action="store_true", help="the command ncu wraps for the launch list: 1 warm-up pass + 1 pass, "
                    "nothing else (no timing claim: the line says profiling_run)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.ncu:
        args.warmup, args.steps = 1, 1
        args.no_cpu_baseline = args.no_secondary = args.no_e2e = True
    if args.groups > 0:
        CORPUS["groups"] = args.groups

    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from codec_eval_b200 import _lib, shard
    from codec_eval_b200.metrics import GpuMetrics, GpuReference, MetricConfig

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: codec_eval_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = shard.bind_to_gpu_numa(local_rank)          # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    w, h = CORPUS["width"], CORPUS["height"]
    nq, img_bytes = len(QUALS), CORPUS["width"] * CORPUS["height"] * 3
    n_total = CORPUS["groups"] * nq
    t_setup = time.perf_counter()
    ctx = GpuMetrics(local_rank, workspace_bytes=48 << 30)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    L = ctx._L

    # ---- the corpus: pool on the host, distortions generated on the device, this rank's shard materialised in HBM
    pool = pool_references(CORPUS["pool_refs"], w, h)                       # [P, h, w, 3]
    P = pool.shape[0]
    d_pool_ref = torch.from_numpy(pool).to(dev)
    d_pool_dist = device_distortions(ctx, d_pool_ref)                       # [P, nq, h, w, 3]
    ref_ids, shards = corpus_layout(world)
    mine = np.asarray(shards[rank], np.int64)                               # global pair indices of this rank, ascending
    n_local = int(mine.size)
    my_groups = np.unique(ref_ids[mine])                                    # ascending
    local_ref_of = np.searchsorted(my_groups, ref_ids[mine]).astype(np.uint32)
    d_ref = d_pool_ref.index_select(0, torch.from_numpy(my_groups % P).to(dev))                     # [G_local, h, w, 3]
    flat_dist = d_pool_dist.view(P * nq, h, w, 3)
    d_dist = flat_dist.index_select(0, torch.from_numpy((ref_ids[mine] % P) * nq + mine % nq).to(dev))   # [n_local, ...]
    n_ref_local = int(my_groups.size)
    torch.cuda.synchronize()
    log(f"[rank {rank}] corpus ready in {time.perf_counter() - t_setup:.1f} s: {n_local} pairs / {n_ref_local} references resident "
        f"({(d_ref.numel() + d_dist.numel()) / 1e9:.1f} GB), numa {numa}")

    mpix_total = n_total * w * h / 1e6
    sub_batch = ctx.sub_batch_capacity(MetricConfig.all(), w, h)

    def resident_step(cfg: MetricConfig):
        def step():
            return shard.evaluate_sharded_resident(ctx, shards, d_ref.data_ptr(), n_ref_local, d_dist.data_ptr(), local_ref_of, w, h,
                                                   cfg, device=dev)
        return step

    def timed(step, steps, warmup):
        for _ in range(warmup):
            step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            last = step()
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, last

    # ---- headline: all four metrics, shards resident in HBM, NCCL gather inside the timed region
    all_cfg = MetricConfig.all()
    step_all = resident_step(all_cfg)
    sampler = ClockSampler(local_rank)
    sampler.start()                       # nvidia-smi needs ~0.2 s to deliver its first sample: start it under the warm-up
    for _ in range(args.warmup):
        step_all()
    l0 = ctx.launch_count()
    sampler.mark()
    ms, table = timed(step_all, args.steps, 0)
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    value = mpix_total * args.steps / (ms / 1e3)
    log(f"[rank {rank}] resident: {value:.0f} MPix-pairs/s, {ms / args.steps:.1f} ms/step")
    results = shard.bytes_to_results(table)
    status_ok = all(results[i].status == 0 and results[i].valid == 15 for i in range(n_total))

    # ---- parity of the TIMED batch: k pairs of this rank's shard, chosen by a seeded draw, against the oracle (rank 0)
    parity = None
    if rank == 0 and not args.ncu:
        from oracle import oracle as O

        k = min(8, n_local)
        pick = np.sort(np.random.default_rng(1234).choice(n_local, size=k, replace=False))
        refs_h = d_ref[torch.from_numpy(local_ref_of[pick].astype(np.int64)).to(dev)].cpu().numpy()
        dists_h = d_dist[torch.from_numpy(pick).to(dev)].cpu().numpy()
        exp = O.evaluate_batch(refs_h, dists_h, w, h, 15, threads=host_cores())
        d_s2 = d_ds = d_ba = d_pn = 0.0
        sse_ok = True
        for j, p in enumerate(pick):
            got, e = results[int(mine[p])], exp[j]
            sse_ok = sse_ok and got.sse == e.sse and (got.psnr == e.psnr or abs(got.psnr - e.psnr) < 1e-12)
            d_s2 = max(d_s2, abs(got.ssimulacra2 - e.ssimulacra2))
            d_ds = max(d_ds, abs(got.dssim - e.dssim) / max(abs(e.dssim), 1e-300))
            d_ba = max(d_ba, abs(got.butteraugli - e.butteraugli) / max(abs(e.butteraugli), 1e-300))
            d_pn = max(d_pn, abs(got.butteraugli_pnorm3 - e.butteraugli_pnorm3) / max(abs(e.butteraugli_pnorm3), 1e-300))
        parity = {"checked": int(k), "of": "pairs of the timed batch (JPEG q50-95) vs the CPU oracle", "sse_exact": bool(sse_ok),
                  "max_abs_ssim2": d_s2, "max_rel_dssim": d_ds, "max_rel_ba": d_ba, "max_rel_ba_pnorm3": d_pn, "tol": TOL,
                  "ok": bool(sse_ok and d_s2 < TOL["ssimulacra2_abs"] and d_ds < TOL["dssim_rel"] and
                             d_ba < TOL["butteraugli_rel"] and d_pn < TOL["butteraugli_rel"] and status_ok)}
        log(f"parity: {parity}")

    # ---- per-kernel CUDA-event pass over the same corpus (profiler on; a few steps: every launch is bracketed)
    peak, peak_src = measured_peak_gbs()

    def kernel_table(prof, steps):
        return {k: {"launches": v["launches"], "ms_per_step": v["ms"] / steps,
                    "gbs": (v["bytes"] / (v["ms"] / 1e3) / 1e9) if v["ms"] > 0 else None,
                    "frac_distinct": (v["bytes"] / (v["ms"] / 1e3) / 1e9 / peak) if v["ms"] > 0 else None,
                    "frac_per_pair": (v["bytes_per_pair"] / (v["ms"] / 1e3) / 1e9 / peak) if v["ms"] > 0 else None}
                for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}

    prof_steps = max(1, min(args.steps, 2))
    prof = {}
    if not args.ncu:
        ctx.profile(True, reset=True)
        prof_ms, _ = timed(step_all, prof_steps, 0)
        prof = ctx.profile_report()
        ctx.profile(False, reset=False)
    kernels = kernel_table(prof, prof_steps)
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    total_bytes = sum(v["bytes"] for v in prof.values())
    traffic_table = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic_table = json.load(open(tpath))
        except Exception:
            traffic_table = {}
    roofline = None
    if prof:
        name, v = max(prof.items(), key=lambda kv: kv[1]["ms"])
        ach = v["bytes"] / (v["ms"] / 1e3) / 1e9
        # what ncu says binds the kernels that are not HBM bound (profiles/r2b_*_summary.txt)
        binding = {"k_ba_malta": "FP32 pipe, not HBM: ncu FMA pipe 68 % busy, 79 % issue-active (16 line sums x 3 bands per pixel)",
                   "k_ds_stats<pair>": "FP32 issue, not HBM: ncu 85 % issue-active, FMA pipe 66 % (un-fused 3x3 chains, dssim-core order)",
                   "k_s2_vpass<pair>": "FP32 issue, not HBM: ncu 80 % issue-active (3 recurrences + SSIM / edge terms per pixel)",
                   "k_s2_hpass<pair>": "its own DRAM traffic: ncu 5.2 TB/s = 80 % of the measured peak (row-pass planes out), 67 % issue-active"}
        tr = traffic_table.get("cfg5", {}).get(name)
        roofline = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "frac_per_pair": v["bytes_per_pair"] / (v["ms"] / 1e3) / 1e9 / peak,
                    "binding_resource": binding.get(name, "HBM"),
                    "traffic": tr.get("bytes_per_launch") if isinstance(tr, dict) else tr,
                    "traffic_source": traffic_table.get("_source"),
                    "peak_source": peak_src, "launches": v["launches"],
                    "avg_launch_ms": v["ms"] / v["launches"], "algorithmic_bytes_per_launch": v["bytes"] / v["launches"],
                    "share_of_kernel_time": v["ms"] / total_kernel_ms if total_kernel_ms else None,
                    # the whole step: algorithmic bytes of every launch / summed kernel time, and the ncu DRAM bytes
                    "step_algorithmic_gb": total_bytes / prof_steps / 1e9,
                    "step_frac": (total_bytes / (total_kernel_ms / 1e3) / 1e9 / peak) if total_kernel_ms else None,
                    "step_dram_gb": traffic_table.get("_step_dram_gb_cfg5"),
                    "step_dram_frac": traffic_table.get("_step_dram_frac_cfg5")}
        roofline["kernels"] = {k: {"ms": round(t["ms_per_step"], 3), "frac": round(t["frac_distinct"] or 0, 3),
                                   "frac_pp": round(t["frac_per_pair"] or 0, 3)} for k, t in list(kernels.items())[:16]}

    # ---- per-metric throughput of the same sharded corpus (BASELINE's metric is per metric)
    per_metric = {}
    if not args.no_secondary:
        for mname, cfg in [("psnr", MetricConfig(psnr=True)), ("ssimulacra2", MetricConfig(ssimulacra2=True)),
                           ("dssim", MetricConfig(dssim=True)), ("butteraugli", MetricConfig(butteraugli=True))]:
            m_ms, _ = timed(resident_step(cfg), 3, 1)
            v = mpix_total * 3 / (m_ms / 1e3)
            gbs = v * 1e6 * STAGED_BYTES_PER_PX[mname] / 1e9 / world
            per_metric[mname] = {"mpix_pairs_per_sec": round(v, 1), "ms_per_step": round(m_ms / 3, 3),
                                 "frac_of_hbm_staged_model": round(gbs / peak, 4)}
        log(f"[rank {rank}] per metric: {per_metric}")

    # ---- end to end: the same corpus from HOST buffers through ce_evaluate_batch.  The host side holds the pool
    # (pinned); every group hands over its own ref_id, so each of the 1250 references and each of the 10,000 distorted
    # images crosses PCIe every step exactly as a full-size host corpus would (host memory is read more than once, the
    # link is not spared).
    e2e = None
    pageable = None
    if not args.no_e2e:
        h_pool_ref = torch.from_numpy(pool).pin_memory()
        h_pool_dist = torch.empty((P * nq, h, w, 3), dtype=torch.uint8).pin_memory()
        h_pool_dist.copy_(flat_dist)
        torch.cuda.synchronize()

        def pair_table(ref_base: int, dist_base: int):
            tab = (_lib.CePair * max(n_local, 1))()
            for i, p in enumerate(mine):
                g, q = int(ref_ids[p]), int(p % nq)
                tab[i] = _lib.CePair(ref_base + (g % P) * img_bytes, dist_base + ((g % P) * nq + q) * img_bytes, img_bytes, img_bytes,
                                     w, h, g, 0)
            return tab

        tab = pair_table(h_pool_ref.data_ptr(), h_pool_dist.data_ptr())

        def step_e2e():
            return shard.evaluate_sharded_table(ctx, shards, tab, all_cfg, device=dev)

        # what the platform gives every rank when all ranks copy at once (pinned -> device, 256 MB x 8): the end-to-end
        # leg cannot beat resident unless this exceeds the rate the shard needs
        probe_h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
        probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        probe_d.copy_(probe_h, non_blocking=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for _ in range(8):
            probe_d.copy_(probe_h, non_blocking=True)
        p1.record(stream)
        torch.cuda.synchronize()
        h2d_gbs = 8 * (256 << 20) / (p0.elapsed_time(p1) / 1e3) / 1e9
        if world > 1:
            t = torch.tensor([h2d_gbs], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            h2d_gbs = float(t.item())
        del probe_h, probe_d
        e2e_ms, e2e_table = timed(step_e2e, args.steps, 2)
        e2e_val = mpix_total * args.steps / (e2e_ms / 1e3)
        same = bool(np.array_equal(e2e_table, table))
        e2e = {"value": e2e_val, "unit": "MPix-pairs/s", "h2d_bytes_per_step": (n_total + CORPUS["groups"]) * img_bytes,
               "d2h_bytes_per_step": n_total * (1 + 108 + 10 + 4) * 8, "ms_per_step": e2e_ms / args.steps,
               "host_memory": "pinned", "results_identical_to_resident": same,
               "h2d_gbs_per_gpu_all_ranks_copying": round(h2d_gbs, 2),
               "h2d_gbs_per_gpu_needed_to_match_resident": round((n_local + n_ref_local) * img_bytes / (ms / args.steps / 1e3) / 1e9, 2)}
        log(f"[rank {rank}] e2e pinned: {e2e_val:.0f} MPix-pairs/s")
        if not args.no_secondary:
            # what the Rust caller holds today: pageable Vec<u8> (src/eval/session.rs:394)
            pg_ref, pg_dist = pool.copy(), h_pool_dist.numpy().copy()
            tab_pg = pair_table(pg_ref.ctypes.data, pg_dist.ctypes.data)
            pg_ms, _ = timed(lambda: shard.evaluate_sharded_table(ctx, shards, tab_pg, all_cfg, device=dev), 2, 1)
            pageable = {"value": round(mpix_total * 2 / (pg_ms / 1e3), 1), "ms_per_step": round(pg_ms / 2, 2),
                        "vs_pinned": round((mpix_total * 2 / (pg_ms / 1e3)) / e2e_val, 4)}
            e2e["pageable"] = pageable
            del pg_ref, pg_dist, tab_pg
        del h_pool_ref, h_pool_dist

    # ---- the other BASELINE configs, rank 0 at N = 1 (each: resident inputs unless it says otherwise)
    other = {}
    if rank == 0 and world == 1 and not args.no_secondary:
        del d_dist, d_ref
        torch.cuda.empty_cache()

        def quick(step, mpix, steps=5, warm=3):
            m, last = timed(step, steps, warm)
            return mpix * steps / (m / 1e3), m / steps, last

        try:    # cfg2: Kodak-shaped batch, 24 refs 768x512 x 8 qualities, all four metrics (round 1's headline)
            w2, h2 = 768, 512
            r2 = torch.from_numpy(pool_references(24, w2, h2)).to(dev)
            x2 = device_distortions(ctx, r2).view(24 * nq, h2, w2, 3)
            ri2 = (np.arange(24 * nq) // nq).astype(np.uint32)
            st2 = lambda cfg: (lambda: ctx.evaluate_batch_device_grouped(r2.data_ptr(), 24, x2.data_ptr(), 24 * nq, ri2, w2, h2, cfg))
            mp2 = 24 * nq * w2 * h2 / 1e6
            v2, ms2, _ = quick(st2(all_cfg), mp2, steps=10)
            ctx.profile(True, reset=True)
            timed(st2(all_cfg), 5, 0)
            k2 = kernel_table(ctx.profile_report(), 5)
            ctx.profile(False, reset=False)
            c2 = {"what": "cfg2: 24 refs 768x512 x 8 JPEG qualities = 192 pairs, all four metrics, resident",
                  "mpix_pairs_per_sec": round(v2, 1), "ms_per_step": round(ms2, 3)}
            for mname, cfg in [("ssimulacra2", MetricConfig(ssimulacra2=True)), ("dssim", MetricConfig(dssim=True)),
                               ("butteraugli", MetricConfig(butteraugli=True)), ("psnr", MetricConfig(psnr=True))]:
                c2[mname] = round(quick(st2(cfg), mp2)[0], 1)
            c2["kernels"] = {k: {"ms": round(t["ms_per_step"], 3), "frac": round(t["frac_distinct"] or 0, 3)} for k, t in list(k2.items())[:12]}
            other["cfg2"] = c2
            other["_cfg2_kernels_full"] = k2
            del r2, x2
        except Exception as e:
            other["cfg2"] = {"error": repr(e)[:200]}
        try:    # cfg3: codec-iter quick eval: 512x512, qualities 75/85/95 x 4 sweep cells, SSIMULACRA2 only, through the
            #       reference handle (Ssimulacra2Reference::new / .compare, crates/codec-iter/src/eval.rs:138-149): host buffers
            w3 = h3 = 512
            refs3 = pool_references(15, w3, h3)
            d3 = device_distortions(ctx, torch.from_numpy(refs3).to(dev), [75, 85, 95] * 4, [2, 2, 2, 0, 0, 0] * 2).cpu().numpy()
            s2cfg = MetricConfig.ssimulacra2_only()
            handles = [GpuReference(ctx, refs3[i], w3, h3, s2cfg) for i in range(15)]

            def step3():
                return [hd.compare_many(list(d3[i])) for i, hd in enumerate(handles)]

            v3, ms3, _ = quick(step3, 15 * 12 * w3 * h3 / 1e6)
            for hd in handles:
                hd.close()
            # the same 180 pairs handed over in ONE ce_evaluate_batch call (what the batched run_eval twin does)
            pairs3 = [(refs3[i], d3[i, k], w3, h3) for i in range(15) for k in range(12)]
            v3b, ms3b, _ = quick(lambda: ctx.evaluate_batch_raw(pairs3, s2cfg), 15 * 12 * w3 * h3 / 1e6)
            other["cfg3"] = {"what": "cfg3: 15 refs 512x512 x 12 distortions, SSIMULACRA2 only via ce_reference_compare_many, host dists",
                             "mpix_pairs_per_sec": round(v3, 1), "ms_per_step": round(ms3, 3),
                             "one_batch_call_mpix_pairs_per_sec": round(v3b, 1), "one_batch_call_ms": round(ms3b, 3)}
        except Exception as e:
            other["cfg3"] = {"error": repr(e)[:200]}
        try:    # cfg4: 256 pairs 3840x2160, Butteraugli + DSSIM only, every pair its own reference (no reuse)
            w4, h4 = 3840, 2160
            r4p = torch.from_numpy(pool_references(8, w4, h4)).to(dev)
            x4p = device_distortions(ctx, r4p, [85], [2]).view(8, h4, w4, 3)
            sel = torch.arange(256, device=dev) % 8
            r4, x4 = r4p.index_select(0, sel), x4p.index_select(0, sel)
            del r4p, x4p
            cfg4 = MetricConfig(butteraugli=True, dssim=True)
            mp4 = 256 * w4 * h4 / 1e6
            st4 = lambda cfg: (lambda: ctx.evaluate_batch_device(r4.data_ptr(), x4.data_ptr(), 256, w4, h4, cfg))
            v4, ms4, _ = quick(st4(cfg4), mp4, steps=3, warm=1)
            c4 = {"what": "cfg4: 256 pairs 3840x2160, Butteraugli + DSSIM, one reference per pair (no reuse), resident",
                  "mpix_pairs_per_sec": round(v4, 1), "ms_per_step": round(ms4, 2)}
            c4["dssim"] = round(quick(st4(MetricConfig(dssim=True)), mp4, steps=3, warm=1)[0], 1)
            c4["butteraugli"] = round(quick(st4(MetricConfig(butteraugli=True)), mp4, steps=3, warm=1)[0], 1)
            other["cfg4"] = c4
            del r4, x4
        except Exception as e:
            other["cfg4"] = {"error": repr(e)[:200]}
        try:
            other["cfg1"] = single_pair_latency(ctx)
        except Exception as e:
            other["cfg1"] = {"error": repr(e)[:200]}
        log(f"other configs: { {k: v for k, v in other.items() if not k.startswith('_')} }")

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on a bounded sample of the same corpus
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O

        cores = host_cores()
        sample = max(8, min(cores, 32))
        refs_c, dists_c = cpu_sample(pool, sample)
        t0 = time.perf_counter()
        reps = 0
        while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 8):
            O.evaluate_batch(refs_c, dists_c, w, h, 15, threads=cores)
            reps += 1
        dt = time.perf_counter() - t0
        cpu = {"value": sample * reps * w * h / 1e6 / dt, "unit": "MPix-pairs/s", "cores": cores, "kind": "port",
               "sample": f"{sample} pairs of the corpus x {reps} passes, all four metrics, OpenMP over pairs, no reference reuse, {dt:.1f} s"}

    if rank == 0:
        cfg_out = corpus_config(world)
        line = {
            "metric": "mpix_pairs_per_sec_all_metrics", "value": value, "unit": "MPix-pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg_out, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        if roofline is not None:
            roofline["sub_batch_pairs"] = sub_batch
            roofline["per_metric"] = per_metric
            roofline["parity"] = parity
            for k in ("cfg1", "cfg2", "cfg3", "cfg4"):
                if k in other:
                    roofline[k] = other[k]
        line["per_metric"] = per_metric
        line["parity"] = parity
        line["lib"] = L.ce_version().decode()
        line["kernels"] = kernels
        line["status_ok"] = status_ok
        if args.ncu:
            line["profiling_run"] = True
        emit(line)
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"config": cfg_out, "steps": args.steps, "value": value, "e2e": e2e, "kernels": kernels, "roofline": roofline,
                           "per_metric": per_metric, "parity": parity, "other": other}, f, indent=1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
