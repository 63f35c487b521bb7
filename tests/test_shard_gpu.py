"""The product sharding path under NCCL on real GPUs (needs >= 2 of them: `gpurun --gpus 2`): two ranks, ragged
shards of mixed image sizes, reference groups kept together, results gathered as raw bytes -- bit-identical to one
rank evaluating the whole batch (the reference analogue runs pairs-parallel on the CPU,
crates/codec-compare/src/full_comparison.rs:319-328; src/eval/session.rs:375-431 is the loop being sharded)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch():
    from codec_eval_b200.synth import G, J

    pairs, ref_ids = [], []
    sizes = [(160, 96), (64, 48), (100, 100), (160, 96), (33, 9), (256, 128), (64, 48)]     # mixed sizes, 7 groups
    for g, (w, h) in enumerate(sizes):
        ref = G(100 + g, w, h)
        for q in ([85, 60, 30] if g % 2 == 0 else [75, 50]):                                 # ragged groups: 3, 2, 3, ...
            pairs.append((ref, J(ref, q, 2 if q > 40 else 0), w, h))
            ref_ids.append(g)
    return pairs, ref_ids


def _key(r):
    return (r.status, r.valid, r.sse, r.dssim, r.ssimulacra2, r.butteraugli, r.psnr, r.butteraugli_pnorm3)


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from codec_eval_b200.metrics import GpuMetrics, MetricConfig
    from codec_eval_b200.shard import evaluate_sharded, evaluate_sharded_resident, partition_pairs

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        pairs, ref_ids = _batch()
        cfg = MetricConfig.all()
        dev = torch.device("cuda", rank)
        with GpuMetrics(rank, workspace_bytes=2 << 30) as ctx:
            got = evaluate_sharded(ctx, pairs, ref_ids, cfg, device=dev)
            whole = ctx.evaluate_batch_raw(pairs, cfg)            # the same batch on this rank alone
            ok = all(_key(got[i]) == _key(whole[i]) for i in range(len(pairs)))
            shards = partition_pairs(ref_ids, [p[2] * p[3] for p in pairs], world)
            # the resident entry: a uniform batch (3 references x 4 distortions of 128x96), shard materialised in HBM
            from codec_eval_b200.synth import G, cheap_distort

            w, h = 128, 96
            refs = [G(7 + g, w, h) for g in range(3)]
            rid = [g for g in range(3) for _ in range(4)]
            dists = [cheap_distort(refs[g], 40 + 15 * k, seed=g) for g in range(3) for k in range(4)]
            sh = partition_pairs(rid, [w * h] * 12, world)
            mine = sh[rank]
            groups = sorted({rid[i] for i in mine})
            d_ref = torch.from_numpy(np.stack([refs[g] for g in groups])).to(dev)
            d_dist = torch.from_numpy(np.stack([dists[i] for i in mine])).to(dev)
            table = evaluate_sharded_resident(ctx, sh, d_ref.data_ptr(), len(groups), d_dist.data_ptr(),
                                              [groups.index(rid[i]) for i in mine], w, h, cfg, device=dev)
            from codec_eval_b200.shard import bytes_to_results

            res = bytes_to_results(table)
            alone = ctx.evaluate_batch_raw([(refs[rid[i]], dists[i], w, h) for i in range(12)], cfg)
            ok2 = all(_key(res[i]) == _key(alone[i]) for i in range(12))
            scores = [(got[i].sse, got[i].ssimulacra2, got[i].dssim, got[i].butteraugli) for i in range(len(pairs))]
        q.put((rank, bool(ok), bool(ok2), [len(s) for s in shards], scores))
    finally:
        dist.destroy_process_group()


def test_evaluate_sharded_nccl_two_ranks(O):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2); the host logic is covered by tests/test_shard.py over gloo")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(o[0] for o in out) == [0, 1]
    assert all(o[1] and o[2] for o in out)
    assert out[0][3] == out[1][3] and sum(out[0][3]) == 18 and min(out[0][3]) > 0 and out[0][3][0] != out[0][3][1]   # ragged
    assert out[0][4] == out[1][4]                               # both ranks hold the same table
    pairs, _ = _batch()
    for i in (0, 4, 9, 17):                                     # and it is the right table
        ref, dist_, w, h = pairs[i]
        sse, s2, ds, ba = out[0][4][i]
        assert sse == O.sse(ref, dist_)
        assert abs(ds - O.dssim(ref, dist_, w, h)) <= 1e-4 * O.dssim(ref, dist_, w, h)
        if w >= 8 and h >= 8:
            assert abs(s2 - O.ssimulacra2(ref, dist_, w, h)) < 0.01
            assert abs(ba - O.butteraugli(ref, dist_, w, h)[0]) <= 1e-3 * O.butteraugli(ref, dist_, w, h)[0]
