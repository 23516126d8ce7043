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


_COUNT_CACHE = {}


def global_count(local_n, device=None):
    """Sum over ranks of a per-rank count (samples in this rank's shard).  Shards may differ by one sample when the batch does not
    divide by the world size (`shard_range`), so synchronised batch statistics must divide by the TRUE global count, not by
    local_n * world_size.  One tiny all-reduce + host read the first time a (local_n, world) pair is seen, cached afterwards: ranks
    keep their shard sizes from step to step (drop_last loaders, `shard_range`)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(local_n)
    key = (int(local_n), dist.get_world_size(), dist.get_rank())
    if key not in _COUNT_CACHE:
        t = torch.tensor([float(local_n)], dtype=torch.float64, device=device if device is not None else "cpu")
        dist.all_reduce(t)
        _COUNT_CACHE[key] = int(round(float(t.item())))
    return _COUNT_CACHE[key]


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


def _up4(n):
    return (n + 3) & ~3


class FlatAdam:
    """Adam over flat buffers with the gradient exchange built in -- the training-step epilogue of SURVEY 2.2 / 8(e).

    Replaces `torch.optim.Adam` + `allreduce_mean_` (and, in the reference, nn.DataParallel's gradient gather,
    GenProjector/trainers/model_trainer.py:20-24, with Adam from RegressionNetwork/train.py:55-57 / pix2pix_model.py:56-70):

    * every parameter becomes a view of ONE flat fp32 buffer, every `.grad` a view of a second one (moments: two more), laid out in
      REVERSE registration order -- the order in which a backward pass finishes gradients -- so a bucket is a contiguous slice and
      the all-reduce runs in place: no `torch.cat`, no copy-back;
    * buckets (>= `bucket_bytes` each) are all-reduced asynchronously AS SOON AS their last gradient exists: a module whose backward
      is one node (emlight_b200.DenseNet) calls `sink(named_grads)` at its block boundaries (`module._grad_sink = opt.sink`), so the
      34 MB fc bucket -- 90 % of the DenseNet's gradient bytes, finished first -- travels over NVLink while the convolutional
      backward still runs; anything not reduced by then is reduced in `step()`;
    * `step()` waits for the collectives and applies ONE fused kernel (`eml_adam_step`: 1/world scaling + Adam, same update as
      torch.optim.Adam) over the flat buffers, then bumps the parameters' version counters (torch.autograd.graph.increment_version) so
      that the modules' packed-weight caches see the change.
    Contract: one backward per step between `zero_grad()` and `step()`; all parameters fp32 on one CUDA device (or CPU tensors with
    gloo for the host-side tests, where the update runs as the same formula in torch ops)."""

    def __init__(self, named_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, bucket_bytes=8 << 20):
        named = [(n, p) for n, p in named_params if p.requires_grad]
        if not named:
            raise ValueError("FlatAdam: no trainable parameters")
        named.reverse()
        dev = named[0][1].device
        if any(p.device != dev or p.dtype != torch.float32 for _, p in named):
            raise ValueError("FlatAdam: parameters must be fp32 tensors on one device")
        self.lr, self.betas, self.eps, self.t = float(lr), (float(betas[0]), float(betas[1])), float(eps), 0
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.offsets, total = {}, 0
        for n, p in named:
            self.offsets[n] = (total, p.numel())
            total += _up4(p.numel())                     # 16-byte aligned slices: the kernels take float4 pointers to the weights
        self.flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.m = torch.zeros_like(self.flat_p)
        self.v = torch.zeros_like(self.flat_p)
        self.grad_views = {}
        with torch.no_grad():
            for n, p in named:
                o, k = self.offsets[n]
                self.flat_p[o:o + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + k].view(p.shape)
                p.grad = self.flat_g[o:o + k].view(p.shape)
                self.grad_views[n] = p.grad
        # buckets: contiguous [lo, hi) slices of the flat buffer, closed once they hold >= bucket_bytes
        self.buckets, lo, members = [], 0, []
        for n in self.names:
            o, k = self.offsets[n]
            members.append(n)
            hi = o + _up4(k)
            if (hi - lo) * 4 >= bucket_bytes:
                self.buckets.append((lo, hi, members)); lo, members = hi, []
        if members:
            self.buckets.append((lo, total, members))
        self._bucket_of = {n: i for i, (_, _, ms) in enumerate(self.buckets) for n in ms}
        self._missing = [set(ms) for _, _, ms in self.buckets]
        self._handles = [None] * len(self.buckets)
        self._reduced = [False] * len(self.buckets)
        self.comm_bytes = total * 4
        self.early_buckets = 0                          # buckets whose all-reduce was launched from inside the backward (last step)
        self.comm = True                                # False: skip the collective (measurement of what the exchange costs; ranks then diverge)
        self.measure = False                            # True: CUDA events around the wait for the collectives in step() -> exposed_ms()
        self._wait_events = []

    # ------------------------------------------------------------------ gradient side
    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        for p, n in zip(self.params, self.names):        # a foreign zero_grad(set_to_none=True) may have dropped the views
            if p.grad is None or p.grad.data_ptr() != self.grad_views[n].data_ptr():
                p.grad = self.grad_views[n]
        self._missing = [set(ms) for _, _, ms in self.buckets]
        self._handles = [None] * len(self.buckets)
        self._reduced = [False] * len(self.buckets)
        self.early_buckets = 0

    def _launch(self, i):
        if self._reduced[i]:
            return
        self._reduced[i] = True
        if self.comm and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            lo, hi, _ = self.buckets[i]
            self._handles[i] = dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, async_op=True)

    def sink(self, named_grads):
        """Called by a module's backward with the gradients finished so far ({parameter name: tensor}); consumes them (the entries
        are removed, autograd then sees None for those parameters) and launches the all-reduce of every bucket that became complete."""
        for n in list(named_grads):
            v = self.grad_views.get(n)
            if v is None:
                continue
            v.add_(named_grads.pop(n).reshape(v.shape))
            i = self._bucket_of[n]
            self._missing[i].discard(n)
            if not self._missing[i]:
                self._launch(i)
                self.early_buckets += 1

    def exposed_ms(self):
        """Mean device time per step() that the compute stream spent WAITING for the gradient all-reduce (measure = True): the part of
        the exchange that the backward did not hide.  Synchronises."""
        if not self._wait_events:
            return 0.0
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self._wait_events) / len(self._wait_events)
        self._wait_events = []
        return ms

    # ------------------------------------------------------------------ update
    @torch.no_grad()
    def step(self):
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        for i in range(len(self.buckets)):
            self._launch(i)
        timed = self.measure and self.flat_p.is_cuda
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for h in self._handles:
            if h is not None:
                h.wait()                                 # the compute stream waits for the NCCL stream: time exposed = not overlapped
        if timed:
            e1.record()
            self._wait_events.append((e0, e1))
        self.t += 1
        b1, b2 = self.betas
        if self.flat_p.is_cuda:
            from . import _lib
            lib = _lib.load()
            with torch.cuda.device(self.flat_p.device):
                _lib.check(lib.eml_adam_step(_lib.ptr(self.flat_p), _lib.ptr(self.flat_g), _lib.ptr(self.m), _lib.ptr(self.v),
                                             self.flat_p.numel(), self.lr, b1, b2, self.eps, self.t, 1.0 / world, _lib.stream_ptr()),
                           "eml_adam_step")
        else:                                            # host-side tests (gloo): the same update in torch ops
            g = self.flat_g / world
            self.m.mul_(b1).add_(g, alpha=1 - b1)
            self.v.mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1, bc2 = 1 - b1 ** self.t, 1 - b2 ** self.t
            self.flat_p.addcdiv_(self.m, self.v.sqrt() / (bc2 ** 0.5) + self.eps, value=-self.lr / bc1)
        torch.autograd.graph.increment_version(self.params)
