"""CPU, world size 2 (gloo): the SynchronizedBatchNorm semantics of SPADE in the training tape.  The reference shares batch statistics
across GPUs inside SPADE (normalization.py:80, sync_batchnorm/batchnorm.py:74-83); the process-per-GPU form is one all-reduce of the
per-channel sums in the forward and one in the backward (`gp_train.spade`).  Two ranks each run a SPADEResnetBlock on their half of a
batch; outputs, input gradients and (summed) parameter gradients must equal one process running the whole batch."""
import argparse
import os
import socket
import sys
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

B, H, W, FIN, FOUT = 4, 4, 8, 8, 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_block(x, seg, gout, seed=3, setter=setattr):
    """One learned-shortcut SPADEResnetBlock (train mode) on the tape: (out, dx, {param name: grad}, running_mean of norm_0)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_gp_train_cpu as T
    from emlight_b200 import gp_ops, gp_train as gt
    from emlight_b200.genprojector import SPADEResnetBlock
    T.install_sims(gp_ops, T.build_emu(tempfile.mkdtemp()), setter)
    torch.manual_seed(seed)
    blk = SPADEResnetBlock(FIN, FOUT, argparse.Namespace(norm_G="spectralspadesyncbatch3x3", semantic_nc=3)).train()
    with torch.no_grad():
        for p in blk.parameters():
            if p.dim() == 1:
                p.uniform_(-0.5, 0.5)                       # non-zero biases
    b = x.shape[0]
    tape = gt.Tape()
    box = {}
    tape.record(lambda: box.setdefault("dx", tape.take(x)))
    out = gt.resnet_block(tape, blk, x, b, H, W, seg, "bf16x3", True)
    tape.add(out, gout)
    tape.backward()
    grads = {n: tape.param_grads[p] for n, p in blk.named_parameters() if p in tape.param_grads}
    return out, box["dx"], grads, blk.norm_0.param_free_norm.running_mean.clone(), blk.norm_0.param_free_norm.running_var.clone()


def _inputs():
    g = torch.Generator().manual_seed(11)
    return (torch.randn(B, H, W, FIN, generator=g) * 2 + 0.3, torch.nn.functional.pad(torch.rand(B, H, W, 3, generator=g), (0, 1)),
            torch.randn(B, H, W, FOUT, generator=g))


def _worker(rank, world, port, q, split):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo")
    x, seg, gout = _inputs()
    lo, hi = (0, split) if rank == 0 else (split, B)
    out, dx, grads, rm, rv = _run_block(x[lo:hi].contiguous(), seg[lo:hi].contiguous(), gout[lo:hi].contiguous())
    q.put((rank, out.numpy(), dx.numpy(), {k: v.numpy() for k, v in grads.items()}, rm.numpy(), rv.numpy()))
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("split", [2, 3], ids=["even_2+2", "uneven_3+1"])
def test_sharded_spade_block_equals_full_batch(monkeypatch, split):
    """split = 3: shards of 3 and 1 samples (what parallel.shard_range yields when the batch does not divide): the synchronised
    statistics must divide by the TRUE global count (parallel.global_count), not by local count x world size."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, split)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    x, seg, gout = _inputs()
    out, dx, grads, rm, rv = _run_block(x, seg, gout, setter=monkeypatch.setattr)
    got_out = torch.cat([torch.from_numpy(r[1]) for r in res])
    got_dx = torch.cat([torch.from_numpy(r[2]) for r in res])
    assert float((got_out - out).abs().max()) <= 1e-4 * float(out.abs().max())
    assert float((got_dx - dx).abs().max()) <= 1e-3 * float(dx.abs().max())
    assert set(res[0][3]) == set(grads) and len(grads) >= 20
    top = max(float(v.abs().max()) for v in grads.values())
    for name, want in grads.items():
        total = sum(torch.from_numpy(r[3][name]) for r in res)         # gradients of a SUM over samples add up across ranks
        assert float((total - want).abs().max()) <= 1e-3 * max(float(want.abs().max()), 1e-3 * top), name
    for r in res:                                                        # every rank holds the global running statistics
        assert float((torch.from_numpy(r[4]) - rm).abs().max()) <= 1e-5 and float((torch.from_numpy(r[5]) - rv).abs().max()) <= 1e-5
