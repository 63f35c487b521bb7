"""Sharding of an image-pair batch over ranks (one process per GPU) and the final score gather.

The path shards by pair with no exchange during compute (every (reference, distorted) pair is independent,
src/eval/session.rs:375-431; codec-compare runs pairs-parallel on CPU, crates/codec-compare/src/full_comparison.rs:319-328).
All distortions of one reference stay on one rank so reference-side reuse survives; ranks are balanced by pixels.
The only collective is one all_gather of the fixed-width result rows, sent as raw bytes so the u64 SSE and the
f64 scores arrive bit-identical.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib

RESULT_BYTES = C.sizeof(_lib.CeResult)  # 56


def partition_pairs(ref_ids: Sequence[int], pixels: Sequence[int], world: int) -> List[List[int]]:
    """Assign pair indices to `world` ranks.  Pairs sharing a ref_id land on the same rank; groups are placed
    largest-first on the least-loaded rank (ties -> lowest rank), so the result is deterministic on every rank."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if len(ref_ids) != len(pixels):
        raise ValueError("ref_ids and pixels differ in length")
    groups = {}
    for i, (g, p) in enumerate(zip(ref_ids, pixels)):
        groups.setdefault(int(g), [0, []])
        groups[int(g)][0] += int(p)
        groups[int(g)][1].append(i)
    order = sorted(groups.items(), key=lambda kv: (-kv[1][0], kv[0]))
    load = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for _, (px, idx) in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += px
        shards[r].extend(idx)
    for s in shards:
        s.sort()
    return shards


def results_to_bytes(results, n: int) -> np.ndarray:
    """ce_result[n] (ctypes array) -> uint8 [n, 56]"""
    if n == 0:
        return np.zeros((0, RESULT_BYTES), np.uint8)
    return np.frombuffer(results, dtype=np.uint8, count=n * RESULT_BYTES).reshape(n, RESULT_BYTES).copy()


def bytes_to_results(rows: np.ndarray):
    """uint8 [n, 56] -> ce_result[n]"""
    n = rows.shape[0]
    out = (_lib.CeResult * max(n, 1))()
    if n:
        C.memmove(out, np.ascontiguousarray(rows).ctypes.data, n * RESULT_BYTES)
    return out


def gather_results(local_rows: np.ndarray, shards: List[List[int]], device=None) -> np.ndarray:
    """All ranks contribute the rows of their shard; every rank gets the [n_total, 56] table in pair order.

    Uses torch.distributed (NCCL on GPUs, gloo on CPU).  Shards may be ragged: rows are padded to the longest
    shard for the fixed-size all_gather and scattered back by the (rank-independent) shard index lists."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    n_total = sum(len(s) for s in shards)
    table = np.zeros((n_total, RESULT_BYTES), np.uint8)
    if world == 1:
        table[shards[0]] = local_rows
        return table
    rank = dist.get_rank()
    assert local_rows.shape[0] == len(shards[rank])
    m = max(len(s) for s in shards)
    send = torch.zeros(m * RESULT_BYTES, dtype=torch.uint8, device=device)
    if local_rows.size:
        send[: local_rows.size] = torch.from_numpy(local_rows.reshape(-1)).to(send.device)
    recv = torch.empty(world * m * RESULT_BYTES, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, m, RESULT_BYTES)
    for r in range(world):
        k = len(shards[r])
        if k:
            table[shards[r]] = recv[r, :k]
    return table


def _world_rank():
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def evaluate_sharded(ctx, pairs: Sequence[Tuple[np.ndarray, np.ndarray, int, int]], ref_ids: Sequence[int], config,
                     intensity_target: float = 80.0, device=None):
    """Each rank evaluates its shard on its own GPU context `ctx`; returns ce_result[n_total] on every rank.
    Host pairs of any mix of sizes (ce_evaluate_batch groups them by size)."""
    world, rank = _world_rank()
    shards = partition_pairs(ref_ids, [p[2] * p[3] for p in pairs], world)
    mine = [pairs[i] for i in shards[rank]]
    out = ctx.evaluate_batch_raw(mine, config, intensity_target)
    rows = results_to_bytes(out, len(mine))
    return bytes_to_results(gather_results(rows, shards, device=device))


def evaluate_sharded_table(ctx, shards: List[List[int]], table, config, intensity_target: float = 80.0, device=None):
    """Same, for a caller that already holds its shard as a ce_pair table (host pointers it keeps alive): `table` is
    ce_pair[len(shards[rank])] in the order of shards[rank].  Returns the uint8 [n_total, 56] result table."""
    world, rank = _world_rank()
    n = len(shards[rank])
    out = ctx.evaluate_pair_table(table, n, config, intensity_target)
    return gather_results(results_to_bytes(out, n), shards, device=device)


def evaluate_sharded_resident(ctx, shards: List[List[int]], d_ref: int, n_ref: int, d_dist: int, ref_index, width: int,
                              height: int, config, intensity_target: float = 80.0, device=None):
    """Same, for a shard that is already resident in this rank's HBM (what an on-device decoder produces): n_ref
    reference images at d_ref, len(shards[rank]) distorted images at d_dist (raw device pointers, tight RGB8,
    uniform size), pair i compares reference ref_index[i].  Returns the uint8 [n_total, 56] result table."""
    world, rank = _world_rank()
    n = len(shards[rank])
    out = ctx.evaluate_batch_device_grouped(d_ref, n_ref, d_dist, n, ref_index, width, height, config, intensity_target)
    return gather_results(results_to_bytes(out, n), shards, device=device)


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Before a rank allocates its pinned staging memory: restrict the process to the CPUs of its GPU's NUMA node
    (intersected with the CPUs it is allowed to use), so page-locked buffers are first-touched next to the PCIe root
    the copies go through.  Best effort -- containers often expose one node only; returns what it did."""
    import os

    info = {"bound": False}
    try:
        import torch

        props = torch.cuda.get_device_properties(local_rank)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bus}"
        with open(f"{base}/numa_node") as f:
            info["numa_node"] = int(f.read().strip())
        with open(f"{base}/local_cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                if not part:
                    continue
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["local_cpus"] = len(cpus)
        info["allowed_cpus"] = len(allowed)
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus"] = len(use)
    except Exception as e:   # no sysfs entry, no permission: run unbound
        info["error"] = repr(e)[:120]
    return info
