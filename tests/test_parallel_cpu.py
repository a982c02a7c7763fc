"""CPU (gloo, world_size 2): the host-side data-parallel logic -- sharding, gradient
bucketing in backward order, sum all-reduce + 1/world scaling == mean over the global
batch (the semantics DDP would give the reference modules)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from avdn_b200 import parallel


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 64, 4096):
        for w in (1, 2, 3, 8):
            ranges = [parallel.shard_range(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_bucket_edges_cover_arena_in_backward_order():
    first = [i * 10 for i in range(57)]
    edges = parallel.bucket_edges(first, 600)
    assert edges[0][2] == 600 and edges[-1][1] == 0 and edges[-1][0] == 0
    for (c0, lo0, hi0), (c1, lo1, hi1) in zip(edges, edges[1:]):
        assert lo0 == hi1 and c0 > c1
    assert parallel.bucket_edges([0], 12) == [(0, 0, 12)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        n = 600
        per_sample = torch.randn(8, n, dtype=torch.float64)          # "gradient" of each of 8 episodes
        lo, hi = parallel.shard_range(8, rank, world)
        # each rank: loss normalised by its LOCAL batch (agent.py:884) -> local mean gradient
        flat = per_sample[lo:hi].mean(0).clone()
        p = torch.full((n,), float(rank))                              # replicas must start identical
        parallel.broadcast_([p], 0)
        first = [i * 10 for i in range(57)]
        done = torch.zeros(n, dtype=torch.bool)
        for c, blo, bhi in parallel.bucket_edges(first, n):
            parallel.allreduce_sum_(flat, blo, bhi)
            assert not done[blo:bhi].any()
            done[blo:bhi] = True
        assert done.all()
        flat *= 1.0 / world                                            # FusedAdamW grad_scale
        ok = torch.allclose(flat, per_sample.mean(0), atol=1e-12) and bool((p == 0).all())
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_bucketed_allreduce_is_global_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
