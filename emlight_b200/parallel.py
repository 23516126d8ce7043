"""Process-per-GPU plumbing: the hot path is per-sample independent, so ranks take disjoint
shards of the batch and there is no data-path collective; the only exchanges are the max-over-ranks device time of
a measurement and -- for training -- one bucketed all-reduce of the gradients per step (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Replaces the reference's nn.DataParallel scatter/replicate/gather
(GenProjector/model_trainer.py:20-24) with torch.distributed."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for a single process)."""
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n items for `rank`; shards differ by at most one item and cover [0, n) exactly."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (used for device-time measurements)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_mean_(tensors, bucket_bytes=32 << 20):
    """In-place mean over ranks of a list of (gradient) tensors, flattened into buckets of ~bucket_bytes so that the
    collective count is set by launch latency, not by the parameter count (DenseNet: 9.3 M params = 37 MB fp32 -> 2 buckets)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    world = dist.get_world_size()
    buckets, cur, cur_bytes = [], [], 0
    for t in tensors:
        if t is None:
            continue
        nb = t.numel() * t.element_size()
        if cur and (cur_bytes + nb > bucket_bytes or t.dtype != cur[0].dtype):
            buckets.append(cur); cur, cur_bytes = [], 0
        cur.append(t); cur_bytes += nb
    if cur:
        buckets.append(cur)
    for b in buckets:
        flat = torch.cat([t.reshape(-1) for t in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        o = 0
        for t in b:
            t.copy_(flat[o:o + t.numel()].view_as(t)); o += t.numel()
    return len(buckets)
