"""Sinkhorn earth-mover loss over sphere anchors -- drop-in for RegressionNetwork/geomloss and gmloss.

``SamplesLoss(loss="sinkhorn", p=2, blur=.05, reach=None, diameter=None, scaling=.5, batchsize=None)``
mirrors geomloss/samples_loss.py:22-46: called as ``loss(x, y)`` with x,y of shape (B,N,1) it returns the (B,)
debiased Sinkhorn divergences, differentiable w.r.t. ``x``.  The anchor count is taken from the input
(the reference hard-wires 96, geomloss/utils.py:66) and ``batchsize`` is accepted but not baked in.
``GMSamplesLoss`` is the gmloss variant: ``loss(x, y, geometry)`` (gmloss/samples_loss.py:34-45).
"""
import numpy as np
import torch
from torch.nn import Module

from . import _lib
from .panorama import sphere_points


def anchor_distance_matrix(n, geometry=None):
    """(n,n) fp32 chord lengths between the fp32 anchors (geomloss/utils.py:64-77; gmloss/utils.py:63-93)."""
    if geometry is None:
        a = sphere_points(n)
    else:
        g = np.asarray(geometry, dtype=np.float64).reshape(-1)
        if g.shape[0] != n:
            raise ValueError("geometry must hold one radius per anchor")
        k = np.arange(n)
        theta = (np.pi * (3 - np.sqrt(5))) * k
        z = np.linspace(1 - 1.0 / n, 1.0 / n - 1, n)
        a = np.stack((g * np.cos(theta), g * np.sin(theta), z), axis=1)
    a = torch.from_numpy(a).float()
    return (a[:, None, :] - a[None, :, :]).norm(dim=2).contiguous()


class _Sinkhorn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, M, blur, scaling, diameter):
        _lib.require_cuda(x, y, M)
        lib = _lib.load()
        B, N = x.shape[0], x.shape[1]
        xd = x.detach().reshape(B, N).contiguous().float()
        yd = y.detach().reshape(B, N).contiguous().float()
        loss = torch.empty(B, device=x.device, dtype=torch.float32)
        grad = torch.empty(B, N, device=x.device, dtype=torch.float32)
        ws_bytes = lib.eml_sinkhorn_workspace_bytes(B, N)
        ws = torch.empty(max(ws_bytes, 16), device=x.device, dtype=torch.uint8)
        _lib.check(lib.eml_sinkhorn_fwdbwd(_lib.ptr(xd), _lib.ptr(yd), _lib.ptr(M), _lib.ptr(loss), _lib.ptr(grad), B, N,
                                           float(blur), float(scaling), float(diameter if diameter else 0.0),
                                           _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "eml_sinkhorn_fwdbwd")
        ctx.save_for_backward(grad)
        ctx.xshape = x.shape
        return loss

    @staticmethod
    def backward(ctx, go):
        (grad,) = ctx.saved_tensors
        gx = (grad * go.reshape(-1, 1)).reshape(ctx.xshape)
        return gx, None, None, None, None, None


class SamplesLoss(Module):
    def __init__(self, loss="sinkhorn", p=2, blur=.05, reach=None, diameter=None, scaling=.5, batchsize=None):
        super().__init__()
        if loss != "sinkhorn" or p != 2 or reach is not None:
            raise ValueError("only the configuration EMLight uses is implemented: loss='sinkhorn', p=2, reach=None")
        self.loss, self.p, self.blur, self.reach = loss, p, blur, reach
        self.diameter, self.scaling, self.batchsize = diameter, scaling, batchsize
        self._M = {}

    def _matrix(self, n, device, geometry=None):
        if geometry is not None:
            return anchor_distance_matrix(n, geometry).to(device)
        key = (n, str(device))
        if key not in self._M:
            self._M[key] = anchor_distance_matrix(n).to(device)
        return self._M[key]

    @_lib.on_tensor_device
    def forward(self, *args):
        if len(args) != 2:
            raise ValueError("A SamplesLoss accepts two (x, y) arguments here (uniform weights, samples_loss.py:62-70).")
        x, y = args
        if x.dim() != 3 or x.shape[2] != 1 or x.shape != y.shape:
            raise ValueError("Input samples 'x' and 'y' should be encoded as (B,N,1) tensors of equal shape.")
        return _Sinkhorn.apply(x, y, self._matrix(x.shape[1], x.device), self.blur, self.scaling, self.diameter)


class GMSamplesLoss(SamplesLoss):
    """gmloss variant: anchors scaled by a per-anchor depth ``geometry`` (rebuilt per call like the reference)."""

    @_lib.on_tensor_device
    def forward(self, x, y, geometry):
        if x.dim() != 3 or x.shape[2] != 1 or x.shape != y.shape:
            raise ValueError("Input samples 'x' and 'y' should be encoded as (B,N,1) tensors of equal shape.")
        if torch.is_tensor(geometry):
            geometry = geometry.detach().cpu().numpy()
        return _Sinkhorn.apply(x, y, self._matrix(x.shape[1], x.device, geometry), self.blur, self.scaling, self.diameter)
