"""`models.networks.sync_batchnorm` as the reference's trainer imports it (GenProjector/model_trainer.py:6,20-24).

The reference spreads ONE process over several GPUs with nn.DataParallel and synchronises SPADE's batch statistics through a host-side
master/slave rendezvous (sync_batchnorm/comm.py).  Here every GPU has its own process (torchrun) and the synchronisation is one
all-reduce of the per-channel sums inside SPADE (emlight_b200/genprojector.py, gp_train.py), so `DataParallelWithCallback` has nothing
left to do: it wraps the module, exposes `.module` like nn.DataParallel, and calls it directly."""
import torch.nn as nn


class DataParallelWithCallback(nn.Module):
    def __init__(self, module, device_ids=None, output_device=None, dim=0):
        super().__init__()
        self.module = module
        self.device_ids = list(device_ids) if device_ids is not None else []

    def forward(self, *inputs, **kwargs):
        return self.module(*inputs, **kwargs)


def patch_replication_callback(data_parallel):
    return data_parallel


SynchronizedBatchNorm1d = nn.BatchNorm1d          # the SPADE layers of this package carry their own synchronised statistics
SynchronizedBatchNorm2d = nn.BatchNorm2d
SynchronizedBatchNorm3d = nn.BatchNorm3d


def convert_model(module):
    return module


def patch_sync_batchnorm():
    import contextlib
    return contextlib.nullcontext()
