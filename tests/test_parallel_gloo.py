"""CPU: the N>1 plumbing with two gloo ranks on 127.0.0.1 (shards, max-over-ranks timing, bucketed gradient all-reduce)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from emlight_b200 import parallel as P
    r, w = P.init("gloo")
    assert (r, w) == (rank, world)
    lo, hi = P.shard_range(257, rank, world)
    t = P.max_over_ranks(10.0 + rank)
    g = torch.Generator().manual_seed(5)
    shapes = [(300, 7), (11,), (1024, 33), (3,)]
    grads = [torch.randn(*s, generator=g) * (rank + 1) for s in shapes]        # rank r holds (r+1) * base
    nb = P.allreduce_mean_(grads, bucket_bytes=64 << 10)
    g2 = torch.Generator().manual_seed(5)
    base = [torch.randn(*s, generator=g2) for s in shapes]
    err = max(float((a - b * (sum(range(1, world + 1)) / world)).abs().max()) for a, b in zip(grads, base))
    out.put((rank, lo, hi, t, nb, err))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, lo0, hi0, t0, nb0, e0), (r1, lo1, hi1, t1, nb1, e1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 129, 129, 257)          # disjoint, contiguous, cover [0, n)
    assert t0 == t1 == 11.0                                     # max over ranks
    assert nb0 == nb1 == 3                                      # 8.4 KB + 44 B | 135 KB | 12 B -> 3 buckets of <= 64 KB (one oversize)
    assert e0 < 1e-6 and e1 < 1e-6                              # mean of r*base over ranks


def test_shard_range_properties():
    from emlight_b200.parallel import shard_range
    for n in (0, 1, 7, 256, 1000):
        for w in (1, 2, 3, 8):
            pieces = [shard_range(n, r, w) for r in range(w)]
            assert pieces[0][0] == 0 and pieces[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            sizes = [hi - lo for lo, hi in pieces]
            assert max(sizes) - min(sizes) <= 1
