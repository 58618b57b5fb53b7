"""world_size-2 gloo test of the only multi-rank logic on the inference path: batch sharding with no data-path
collective, timing as max over ranks (bench.py's helpers)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    lo, hi = bench.shard_range(10, rank, world)
    ms = bench.max_over_ranks(10.0 * (rank + 1), device="cpu")
    tot = bench.sum_over_ranks(float(hi - lo), device="cpu")
    q.put((rank, lo, hi, ms, tot))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_max_reduce_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]
    assert all(r[3] == 20.0 for r in res)  # max over ranks
    assert all(r[4] == 10.0 for r in res)  # units summed over ranks


def test_shard_range_covers_everything():
    import bench
    for n in (1, 7, 64, 129):
        for w in (1, 2, 3, 8):
            spans = [bench.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
