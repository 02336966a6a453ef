"""World-size-2 gloo test of the multi-rank path on CPU: contiguous guide partition, per-rank
scoring, max-over-ranks timing and in-order gather.  The per-rank scorer here is the oracle (this
is a test; on the GPU box every rank calls libissl_cuda on its own device instead)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from crackling_b200.sharding import shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100_000, 10_000_019):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
    sys.path.insert(0, os.path.dirname(__file__))
    import torch.distributed as dist
    from conftest import golden_case
    from crackling_b200.sharding import gather_in_order, max_over_ranks, shard_bounds
    from oracle import oracle
    import issl_testdata as td

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    case = golden_case("w8_families")
    guides = td.pack_guides(case.guides)
    lo, hi = shard_bounds(guides.size, world, rank)
    r = oracle.score(case.issl, guides[lo:hi], 4, 75, "and", threads=1)
    t = max_over_ranks(1.0 + rank, dist)
    mit = gather_in_order(r["mit"], guides.size, dist)
    cfd = gather_in_order(r["cfd"], guides.size, dist)
    if rank == 0:
        full = oracle.score(case.issl, guides, 4, 75, "and", threads=1)
        q.put((t, bool(np.array_equal(mit, full["mit"])), bool(np.array_equal(cfd, full["cfd"]))))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    t, mit_ok, cfd_ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == 2.0 and mit_ok and cfd_ok
