"""CPU tests of the N>1 path: pair sharding and the final score gather, world_size 2 over gloo.
No compute here (no GPU): each rank fabricates the ce_result rows of its shard from the pair index."""
import ctypes as C
import os
import socket

import numpy as np
import pytest


def test_partition_keeps_reference_groups_and_balances():
    from codec_eval_b200.shard import partition_pairs

    ref_ids = [i // 8 for i in range(192)]           # 24 refs x 8 qualities (cfg2)
    px = [768 * 512] * 192
    for world in (1, 2, 4, 8):
        shards = partition_pairs(ref_ids, px, world)
        assert sorted(sum(shards, [])) == list(range(192))
        for s in shards:
            assert len(s) == 192 // world
            groups = {}
            for i in s:
                groups.setdefault(ref_ids[i], 0)
                groups[ref_ids[i]] += 1
            assert all(v == 8 for v in groups.values())   # a reference never straddles ranks
    # ragged, mixed sizes: balanced by pixels, deterministic
    ref_ids = [0, 0, 0, 1, 2, 2, 3]
    px = [100, 100, 100, 1000, 10, 10, 500]
    a = partition_pairs(ref_ids, px, 2)
    assert a == partition_pairs(ref_ids, px, 2)
    assert a == [[3], [0, 1, 2, 4, 5, 6]]
    assert partition_pairs([], [], 3) == [[], [], []]
    with pytest.raises(ValueError):
        partition_pairs([0], [1, 2], 2)


def _fake_rows(idx):
    from codec_eval_b200 import _lib

    arr = (_lib.CeResult * max(len(idx), 1))()
    for k, i in enumerate(idx):
        arr[k].status = 0
        arr[k].valid = 15
        arr[k].sse = (1 << 40) + i            # needs all 64 bits to survive the gather
        arr[k].ssimulacra2 = 100.0 - i / 7.0
        arr[k].dssim = i * 1e-5
        arr[k].butteraugli = float(np.nextafter(1.0 + i, 2.0 + i))
        arr[k].psnr = np.inf if i == 0 else 30.0 + i
        arr[k].butteraugli_pnorm3 = i / 3.0
    return arr


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from codec_eval_b200.shard import bytes_to_results, gather_results, partition_pairs, results_to_bytes

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ref_ids = [0, 0, 0, 1, 2, 2, 3, 3, 3, 3, 4]
        px = [64 * 64] * 3 + [256 * 256] + [96 * 80] * 2 + [128 * 128] * 4 + [32 * 32]
        shards = partition_pairs(ref_ids, px, world)
        mine = shards[rank]
        rows = results_to_bytes(_fake_rows(mine), len(mine))
        table = gather_results(rows, shards)
        res = bytes_to_results(table)
        ok = True
        exp = _fake_rows(list(range(len(ref_ids))))
        for i in range(len(ref_ids)):
            ok &= res[i].sse == exp[i].sse and res[i].ssimulacra2 == exp[i].ssimulacra2
            ok &= res[i].dssim == exp[i].dssim and res[i].butteraugli == exp[i].butteraugli
            ok &= res[i].psnr == exp[i].psnr and res[i].valid == 15
        # the three sharded entries, over a stand-in context that answers for whatever shard it is handed
        from codec_eval_b200.shard import evaluate_sharded, evaluate_sharded_resident, evaluate_sharded_table

        class Ctx:
            def __init__(self):
                self.seen = []

            def evaluate_batch_raw(self, pairs, config, intensity_target=80.0):
                self.seen.append(("raw", [int(p[0][0]) for p in pairs]))
                return _fake_rows([int(p[0][0]) for p in pairs])      # pair i carries its index in its first byte

            def evaluate_pair_table(self, table, n, config, intensity_target=80.0):
                self.seen.append(("table", n))
                return _fake_rows([int(table[k]) for k in range(n)])

            def evaluate_batch_device_grouped(self, d_ref, n_ref, d_dist, n, ref_index, w, h, config, intensity_target=80.0):
                self.seen.append(("resident", n, n_ref, list(ref_index)))
                return _fake_rows(list(range(d_dist, d_dist + n)))

        fake = Ctx()
        pairs = [(np.full(3, i, np.uint8), np.full(3, i, np.uint8), int(np.sqrt(px[i])), int(np.sqrt(px[i]))) for i in range(len(ref_ids))]
        for r3 in (bytes_to_results(np.asarray(t)) if not hasattr(t, "_length_") else t for t in (
                evaluate_sharded(fake, pairs, ref_ids, None),
                evaluate_sharded_table(fake, shards, mine, None))):
            for i in range(len(ref_ids)):
                ok &= r3[i].sse == exp[i].sse and r3[i].butteraugli == exp[i].butteraugli
        ok &= fake.seen[0] == ("raw", mine) and fake.seen[1] == ("table", len(mine))
        # resident: a uniform batch of 9 pairs / 3 references; rank r's shard starts at "device pointer" shards2[r][0]
        rid2 = [0, 0, 0, 1, 1, 1, 2, 2, 2]
        shards2 = partition_pairs(rid2, [100] * 9, world)
        mine2 = shards2[rank]
        local_refs = sorted({rid2[i] for i in mine2})
        t2 = evaluate_sharded_resident(fake, shards2, 0, len(local_refs), mine2[0] if mine2 else 0,
                                       [local_refs.index(rid2[i]) for i in mine2], 10, 10, None)
        r2 = bytes_to_results(t2)
        exp2 = _fake_rows(list(range(9)))
        contiguous = all(s2 == list(range(s2[0], s2[0] + len(s2))) for s2 in shards2 if s2)
        if contiguous:
            for i in range(9):
                ok &= r2[i].sse == exp2[i].sse
        q.put((rank, bool(ok), [len(s) for s in shards]))
    finally:
        dist.destroy_process_group()


def test_gather_results_gloo_world2():
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _, _ in out) == [0, 1]
    assert all(ok for _, ok, _ in out)
    assert out[0][2] == out[1][2] and sum(out[0][2]) == 11 and min(out[0][2]) > 0   # ragged shards


def test_result_bytes_roundtrip():
    from codec_eval_b200.shard import RESULT_BYTES, bytes_to_results, results_to_bytes

    assert RESULT_BYTES == 56
    a = _fake_rows([0, 5, 9])
    b = bytes_to_results(results_to_bytes(a, 3))
    assert bytes(C.string_at(C.addressof(a), 3 * 56)) == bytes(C.string_at(C.addressof(b), 3 * 56))
    assert results_to_bytes(a, 0).shape == (0, 56)


def test_partition_properties_on_random_batches():
    """Properties that hold for any batch: every pair is assigned exactly once, a reference's pairs stay together,
    the assignment does not depend on who computes it, and the load obeys the greedy (largest-first) bound
    max_load <= mean_load + largest_group."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from codec_eval_b200.shard import partition_pairs

    batches = st.lists(st.tuples(st.integers(0, 12), st.sampled_from([64 * 64, 512 * 512, 768 * 512, 3840 * 2160])),
                       min_size=0, max_size=60)

    @settings(max_examples=150, deadline=None)
    @given(batches, st.integers(1, 8))
    def check(batch, world):
        ref_ids = [b[0] for b in batch]
        px = [b[1] for b in batch]
        shards = partition_pairs(ref_ids, px, world)
        assert len(shards) == world
        assert sorted(i for s in shards for i in s) == list(range(len(batch)))
        owner = {}
        for r, s in enumerate(shards):
            assert s == sorted(s)
            for i in s:
                assert owner.setdefault(ref_ids[i], r) == r
        assert shards == partition_pairs(list(ref_ids), list(px), world)
        if batch:
            group = {}
            for g, p in batch:
                group[g] = group.get(g, 0) + p
            loads = [sum(px[i] for i in s) for s in shards]
            assert max(loads) <= sum(px) / world + max(group.values())

    check()
