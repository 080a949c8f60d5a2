"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: tracklet sharding and the final gather."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from objectcentricocccompletion_b200 import dist as occ_dist
        from objectcentricocccompletion_b200 import synth
        from oracle import oracle
        from oracle.make_golden import edge_batch

        batch = edge_batch()                      # contains short / no-point tracklets as well
        extra = synth.make_batch(3, 10, 0.2, seed=2, small=True)
        nseg = len(batch.segments)
        batch.segments += extra.segments
        for t in extra.tracklets:
            t.segment += nseg
        batch.tracklets += extra.tracklets
        res = occ_dist.annotate_distributed(batch, annotate_fn=oracle.annotate_batch)
        sub, mine = occ_dist.shard_batch(batch, rank, world)
        if rank == 0:
            exp = oracle.annotate_batch(batch)
            ok = len(res) == len(exp)
            for r, e in zip(res, exp):
                ok &= r is not None and r["status"] == e["status"]
                if e["occ"] is not None:
                    ok &= bool((r["occ"] == e["occ"]).all()) and r["n_unknown"] == e["n_unknown"]
            q.put(("result", bool(ok), mine))
        else:
            q.put(("shard", res is None, mine))
    finally:
        dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    by = {g[0]: g for g in got}
    assert by["result"][1] is True and by["shard"][1] is True
    a, b = sorted(by["result"][2]), sorted(by["shard"][2])
    assert sorted(a + b) == list(range(8)) and not set(a) & set(b)          # a partition of the tracklets


def test_shard_indices_balance():
    sys.path.insert(0, ROOT)
    from objectcentricocccompletion_b200.dist import shard_indices

    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 100, 1000).tolist()
    for w in (1, 2, 4, 8):
        sh = shard_indices(costs, w)
        assert sorted(i for s in sh for i in s) == list(range(1000))
        loads = [sum(costs[i] for i in s) for s in sh]
        assert max(loads) - min(loads) <= max(costs)                        # LPT bound
    assert shard_indices([], 4) == [[], [], [], []]


def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from objectcentricocccompletion_b200 import dist as occ_dist

        n = 1000 + 37 * rank                                     # uneven payloads
        local = torch.full((n + 5,), rank + 1, dtype=torch.uint8)  # the buffer may be larger than the payload
        sizes = occ_dist.exchange_sizes(n)
        out = occ_dist.gather_labels(local, sizes, dst=0)
        again = occ_dist.gather_labels(local, sizes, dst=0, out=out)      # reusing the destination buffer
        if rank == 0:
            want = torch.cat([torch.full((1000 + 37 * r,), r + 1, dtype=torch.uint8) for r in range(world)])
            q.put(("dst", sizes == [1000 + 37 * r for r in range(world)], bool((out == want).all()), again is out))
        else:
            q.put(("src", out is None and again is None, True, True))
    finally:
        dist.destroy_process_group()


def test_gather_labels_exact_sizes_world3():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gather_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    got = [q.get(timeout=180) for _ in range(3)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(g[1] and g[2] and g[3] for g in got), got


def test_shard_by_segment_keeps_only_referenced_segments():
    sys.path.insert(0, ROOT)
    from objectcentricocccompletion_b200 import dist as occ_dist
    from objectcentricocccompletion_b200 import synth

    batch = synth.make_batch(12, 10, 0.2, seed=4, small=True, tracklets_per_segment=3)      # 4 segments
    seen = []
    for rank in range(2):
        sub, mine = occ_dist.shard_batch(batch, rank, 2, by="segment")
        assert len(sub.segments) == 2 and len(sub.tracklets) == 6                          # whole segments
        for t, gi in zip(sub.tracklets, mine):
            src = batch.tracklets[gi]
            assert sub.segments[t.segment] is batch.segments[src.segment] and t.boxes is src.boxes
        seen += mine
    assert sorted(seen) == list(range(12))
    sub, mine = occ_dist.shard_batch(batch, 0, 8, by="tracklet")                            # more ranks than segments
    assert len(sub.segments) == len({batch.tracklets[i].segment for i in mine})
