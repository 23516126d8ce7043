"""Spherical-Gaussian panorama reconstruction -- drop-in for RegressionNetwork/util.py.

``sphere_points(n)``                      util.py:286-299   (host-side numpy, float64, like the reference)
``convert_to_panorama(dirs,sizes,colors)`` util.py:222-245   (B,3N),(B,N),(B,3N) -> (B,3,128,256), differentiable
``render_from_params(...)``               train.py:115-122 / GenProjector/data.py:86-102 fused: heads -> panorama
"""
import numpy as np
import torch

from . import _lib


def sphere_points(n=128):
    """Fibonacci lattice on the unit sphere, (n,3) float64 (util.py:286-299)."""
    k = np.arange(n)
    theta = (np.pi * (3 - np.sqrt(5))) * k
    z = np.linspace(1 - 1.0 / n, 1.0 / n - 1, n)
    rad = np.sqrt(1 - z * z)
    return np.stack((rad * np.cos(theta), rad * np.sin(theta), z), axis=1)


class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, sizes, colors):
        _lib.require_cuda(dirs, sizes, colors)
        lib = _lib.load()
        B = colors.shape[0]
        if colors.shape[1] % 3:
            raise ValueError("colors must be (B, 3N)")
        N = colors.shape[1] // 3
        if dirs.shape != (B, 3 * N) or sizes.shape != (B, N):
            raise ValueError("expected dirs (B,3N), sizes (B,N), colors (B,3N); got %s %s %s"
                             % (tuple(dirs.shape), tuple(sizes.shape), tuple(colors.shape)))
        d, s, c = (t.detach().contiguous().float() for t in (dirs, sizes, colors))
        out = torch.empty(B, 3, 128, 256, device=colors.device, dtype=torch.float32)
        _lib.check(lib.eml_sg_render_fwd(_lib.ptr(d), 3 * N, _lib.ptr(s), N, _lib.ptr(c), None, _lib.ptr(out),
                                         B, N, _lib.stream_ptr()), "eml_sg_render_fwd")
        ctx.save_for_backward(d, s, c)
        return out

    @staticmethod
    def backward(ctx, go):
        d, s, c = ctx.saved_tensors
        lib = _lib.load()
        B, N = s.shape
        go = go.contiguous().float()
        need = ctx.needs_input_grad
        gd = torch.empty_like(d) if need[0] else None
        gs = torch.empty_like(s) if need[1] else None
        gc = torch.empty_like(c) if need[2] else None
        _lib.check(lib.eml_sg_render_bwd(_lib.ptr(d), 3 * N, _lib.ptr(s), N, _lib.ptr(c), _lib.ptr(go),
                                         _lib.ptr(gd), _lib.ptr(gs), _lib.ptr(gc), B, N, _lib.stream_ptr()),
                   "eml_sg_render_bwd")
        return gd, gs, gc


@_lib.on_tensor_device
def convert_to_panorama(dirs, sizes, colors):
    """Same signature and result as the reference's util.convert_to_panorama (CUDA tensors only)."""
    return _Render.apply(dirs, sizes, colors)


@_lib.on_tensor_device
@torch.no_grad()
def render_from_params(distribution, intensity, rgb_ratio, ambient=None, dirs=None, size=0.0025, gain=500.0):
    """Heads of the regression network -> (B,3,128,256) panorama in one launch (no (B,3N) colour tensor).

    colors[b,k,:] = distribution[b,k] * intensity[b] * gain * rgb_ratio[b,:]  (train.py:117-121); ``ambient`` (B,3)
    is added to every pixel when given.  ``dirs``: (N,3)/(3N,) shared anchors (default: sphere_points(N)) or (B,3N).
    """
    _lib.require_cuda(distribution, intensity, rgb_ratio)
    lib = _lib.load()
    B, N = distribution.shape
    dev = distribution.device
    dist = distribution.contiguous().float()
    inten = intensity.contiguous().float().view(B)
    rgb = rgb_ratio.contiguous().float()
    if dirs is None:
        dirs = torch.from_numpy(sphere_points(N)).float().to(dev)
    dirs = dirs.contiguous().float().to(dev)
    dirs_bs = 3 * N if dirs.numel() == B * 3 * N and dirs.dim() == 2 and dirs.shape[0] == B and B > 1 else 0
    if dirs.numel() not in (3 * N, B * 3 * N):
        raise ValueError("dirs must hold 3N or B*3N values")
    if torch.is_tensor(size):
        sizes = size.contiguous().float().to(dev)
        sizes_bs = N if sizes.numel() == B * N and B > 1 else 0
    else:
        sizes = torch.full((N,), float(size), device=dev)
        sizes_bs = 0
    amb = None if ambient is None else ambient.contiguous().float()
    out = torch.empty(B, 3, 128, 256, device=dev, dtype=torch.float32)
    _lib.check(lib.eml_sg_render_params_fwd(_lib.ptr(dirs), dirs_bs, _lib.ptr(sizes), sizes_bs, _lib.ptr(dist), N,
                                            _lib.ptr(inten), 1, _lib.ptr(rgb), 3, float(gain), _lib.ptr(amb), 3,
                                            _lib.ptr(out), B, N, _lib.stream_ptr()), "eml_sg_render_params_fwd")
    return out


@_lib.on_tensor_device
def genprojector_guide(distribution, intensity, rgb_ratio, ambient, alpha=1.0, dirs=None, size=0.0025):
    """The GenProjector's conditioning panorama from (predicted or ground-truth) light parameters -- GenProjector/data.py:86-102:
        env = (convert_to_panorama(dirs, 0.0025, dist * (intensity * 0.01) * rgb_ratio) + ambient / (128 * 256)) * alpha
    in ONE render launch (the render is linear in the colours, so alpha and the 0.01 gain fold into them and the ambient term rides
    in the kernel's per-image offset).  SURVEY 8f rank 3: this is what lets DenseNet -> render -> SPADE generator run in one process.
    distribution (B,N), intensity (B,) or (B,1), rgb_ratio (B,3), ambient (B,3), alpha scalar or (B,)  ->  (B,3,128,256)."""
    _lib.require_cuda(distribution, intensity, rgb_ratio, ambient)
    B = distribution.shape[0]
    a = torch.as_tensor(alpha, dtype=torch.float32, device=distribution.device).reshape(-1)
    a = a.expand(B) if a.numel() == 1 else a
    if a.numel() != B:
        raise ValueError("alpha must be a scalar or hold one value per image")
    inten = intensity.float().reshape(B) * a
    amb = ambient.float().reshape(B, 3) * (a / float(128 * 256)).unsqueeze(1)
    return render_from_params(distribution, inten, rgb_ratio, ambient=amb, dirs=dirs, size=size, gain=0.01)
