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


# --------------------------------------------------------------------------------- training: gradient exchange
def _dp_worker(rank, world, port, q):
    """Each rank holds the same replicated parameters (same seed), its own shard of a global batch, and a
    rank-specific gradient; after allreduce_gradients + the 1/world factor every rank holds the global mean, and a
    plain SGD-like update with it keeps the replicas bit-identical (the property DP training relies on)."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch.nn as nn
    from hifigan_b200.train import FlatParams, allreduce_gradients, shard_batch
    torch.manual_seed(7)
    net = nn.Sequential(nn.Conv1d(3, 5, 3), nn.Conv1d(5, 2, 1))
    flat = FlatParams(net, "cpu")
    # parameters are views into the flat buffer, gradients views into flat.g
    assert all(p.data_ptr() >= flat.p.data_ptr() for p in net.parameters())
    lo, hi = shard_batch(8, rank, world)
    data = torch.arange(8, dtype=torch.float32)
    local = data[lo:hi].sum()
    for p in net.parameters():
        p.grad.fill_(float(local))          # stands in for this rank's summed per-sample gradients
    # exchanged slice by slice, as the sub-discriminator lanes do (DiscriminatorTrainer.run_phases): the two convs'
    # spans tile the buffer, so the result must equal one all-reduce of the whole buffer
    spans = [flat.span_of(m) for m in net]
    assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == flat.numel
    for sp in spans:
        scale = allreduce_gradients(flat, span=sp)
    mean = (flat.g * scale)
    flat.p.sub_(0.1 * mean)
    q.put((rank, lo, hi, float(mean[0]), float(scale), flat.p.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_keeps_replicas_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 8)]
    assert all(r[4] == 0.5 for r in res)
    assert all(abs(r[3] - 28.0 / 2) < 1e-6 for r in res)          # (0+1+2+3 + 4+5+6+7) / world
    assert torch.equal(res[0][5], res[1][5])                       # replicas stay bit-identical


def test_shard_batch_rejects_ragged_split():
    import pytest
    from hifigan_b200.train import shard_batch
    assert shard_batch(128, 3, 8) == (48, 64)
    with pytest.raises(ValueError):
        shard_batch(130, 0, 8)
