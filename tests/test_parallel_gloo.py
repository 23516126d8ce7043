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


def _mlp():
    torch.manual_seed(3)
    return torch.nn.Sequential(torch.nn.Linear(7, 33), torch.nn.Tanh(), torch.nn.Linear(33, 5), torch.nn.Tanh(), torch.nn.Linear(5, 2))


def _flat_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from emlight_b200 import parallel as P
    P.init("gloo")
    net = _mlp()
    opt = P.FlatAdam(net.named_parameters(), lr=1e-2, betas=(0.9, 0.999), bucket_bytes=512)
    g = torch.Generator().manual_seed(11)
    x, y = torch.randn(8, 7, generator=g), torch.randn(8, 2, generator=g)
    lo, hi = P.shard_range(8, rank, world)
    sunk = []
    for step in range(4):
        opt.zero_grad()
        loss = ((net(x[lo:hi]) - y[lo:hi]) ** 2).sum() / (hi - lo)              # per-rank mean; the all-reduce averages the ranks
        if step % 2 == 0:
            loss.backward()                                                     # gradients arrive through autograd (+= into the views)
        else:                                                                   # ... or through the sink, bucket by bucket, like DenseNet._backward
            names = [n for n, _ in net.named_parameters()]
            grads = torch.autograd.grad(loss, list(net.parameters()))
            named = dict(zip(names, grads))
            for n in reversed(names):
                opt.sink({n: named[n]})
            sunk.append(opt.early_buckets)
        opt.step()
    out.put((rank, [p.detach().clone() for p in net.parameters()], len(opt.buckets), sunk,
             all(p.data_ptr() >= opt.flat_p.data_ptr() for p in net.parameters()), [p._version for p in net.parameters()]))
    dist.destroy_process_group()


def test_flat_adam_two_ranks_equals_torch_adam_on_the_whole_batch():
    """FlatAdam (flat views, in-place bucketed all-reduce launched from the sink, fused update) on two gloo ranks with half batches
    == torch.optim.Adam in one process on the whole batch."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_flat_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    net = _mlp()
    ref = torch.optim.Adam(net.parameters(), lr=1e-2, betas=(0.9, 0.999))
    g = torch.Generator().manual_seed(11)
    x, y = torch.randn(8, 7, generator=g), torch.randn(8, 2, generator=g)
    for _ in range(4):
        ref.zero_grad()
        (((net(x) - y) ** 2).sum() / 8).backward()
        ref.step()
    for rank, params, nb, sunk, in_flat, versions in res:
        assert nb >= 2 and in_flat and sunk == [nb, nb]                       # every bucket was launched from inside the "backward"
        assert all(v >= 4 for v in versions)                                    # version counters bumped by every step (cache invalidation)
        for a, b in zip(params, net.parameters()):
            assert float((a - b.detach()).abs().max()) < 1e-6
    for a, b in zip(res[0][1], res[1][1]):
        assert torch.equal(a, b)                                                # ranks stay bit-identical
