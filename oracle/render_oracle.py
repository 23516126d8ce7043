"""CPU oracle for the spherical-Gaussian -> equirect panorama render (test infrastructure).

Restates, in numpy, these reference functions (identical copies exist in
representation/util.py:190-228, GenProjector/util.py:346-420, Needlets/utils.py:10-33,191-203):

* ``sphere_points``        RegressionNetwork/util.py:286-299   Fibonacci lattice, float64
* ``convert_to_panorama``  RegressionNetwork/util.py:222-245   out[b,ch,r,c] = sum_k colors[b,3k+ch] *
                                                               exp((dirs[b,3k:3k+3] . p(r,c) - 1) / sizes[b,k])
  pixel directions p(r,c) (util.py:223-233): lat=(r+.5)*pi/128, lon=(c+.5)*pi/128,
  p = (sin lat cos lon, sin lat sin lon, cos lat), fp32.
* ``compose_colors``       RegressionNetwork/train.py:117-121  colors[b,k,ch] = dist[b,k]*intensity[b]*gain*rgb[b,ch]
                           (gain 500 in train.py:117; GenProjector/data.py:87 uses 0.01)
"""
import numpy as np

PANO_H, PANO_W = 128, 256


def sphere_points(n=128):
    golden_angle = np.pi * (3 - np.sqrt(5))
    theta = golden_angle * np.arange(n)
    z = np.linspace(1 - 1.0 / n, 1.0 / n - 1, n)
    radius = np.sqrt(1 - z * z)
    pts = np.zeros((n, 3))
    pts[:, 0] = radius * np.cos(theta)
    pts[:, 1] = radius * np.sin(theta)
    pts[:, 2] = z
    return pts


def pixel_dirs(dtype=np.float32):
    """(3, 128, 256) unit vectors, computed in fp32 like the reference's torch ops."""
    lat = ((np.arange(PANO_H, dtype=np.float32) + np.float32(0.5)) * np.float32(np.pi / 128))[:, None]
    lon = ((np.arange(PANO_W, dtype=np.float32) + np.float32(0.5)) * np.float32(np.pi / 128))[None, :]
    lat = np.broadcast_to(lat, (PANO_H, PANO_W)).astype(np.float32)
    lon = np.broadcast_to(lon, (PANO_H, PANO_W)).astype(np.float32)
    x = np.sin(lat) * np.cos(lon)
    y = np.sin(lat) * np.sin(lon)
    z = np.cos(lat)
    return np.stack((x, y, z)).astype(dtype)


def convert_to_panorama(dirs, sizes, colors, dtype=np.float32):
    """dirs (B,3N), sizes (B,N), colors (B,3N) -> (B,3,128,256). Accumulates light by light like the reference."""
    dirs = np.asarray(dirs, dtype=dtype)
    sizes = np.asarray(sizes, dtype=dtype)
    colors = np.asarray(colors, dtype=dtype)
    xyz = pixel_dirs(dtype).reshape(3, -1)
    B = colors.shape[0]
    n = colors.shape[1] // 3
    out = np.zeros((B, 3, PANO_H * PANO_W), dtype=dtype)
    for k in range(n):
        d = dirs[:, 3 * k:3 * k + 3] @ xyz                       # (B, P)
        g = np.exp((d - dtype(1)) / sizes[:, k:k + 1])           # (B, P)
        out = out + colors[:, 3 * k:3 * k + 3][:, :, None] * g[:, None, :]
    return out.reshape(B, 3, PANO_H, PANO_W)


def convert_to_panorama_grad(dirs, sizes, colors, grad_out):
    """Analytic gradients (float64) of sum(out*grad_out) w.r.t. dirs, sizes, colors."""
    dirs = np.asarray(dirs, np.float64); sizes = np.asarray(sizes, np.float64)
    colors = np.asarray(colors, np.float64); go = np.asarray(grad_out, np.float64).reshape(colors.shape[0], 3, -1)
    xyz = pixel_dirs(np.float64).reshape(3, -1)
    B = colors.shape[0]; n = colors.shape[1] // 3
    gd = np.zeros_like(dirs); gs = np.zeros_like(sizes); gc = np.zeros_like(colors)
    for k in range(n):
        d = dirs[:, 3 * k:3 * k + 3] @ xyz
        g = np.exp((d - 1.0) / sizes[:, k:k + 1])
        gc[:, 3 * k:3 * k + 3] = np.einsum('bcp,bp->bc', go, g)
        w = np.einsum('bcp,bc->bp', go, colors[:, 3 * k:3 * k + 3]) * g   # dL/d(exponent)
        gd[:, 3 * k:3 * k + 3] = (w / sizes[:, k:k + 1]) @ xyz.T
        gs[:, k] = -(w * (d - 1.0)).sum(1) / sizes[:, k] ** 2
    return gd, gs, gc


def compose_colors(dist, intensity, rgb_ratio, gain=500.0):
    """(B,N),(B,1),(B,3) -> (B,3N) k-major (train.py:117-121)."""
    c = dist[:, :, None] * (intensity[:, :, None] * np.float32(gain)) * rgb_ratio[:, None, :]
    return c.reshape(dist.shape[0], -1).astype(np.float32)


def convert_to_panorama_torch(dirs, sizes, colors):
    """The same light-by-light accumulation written with torch CPU ops (what the reference executes, util.py:222-245);
    used as the timed CPU baseline.  dirs (B,3N), sizes (B,N), colors (B,3N) torch fp32 tensors."""
    import torch
    xyz = torch.from_numpy(pixel_dirs(np.float32)).reshape(3, -1).to(dirs.device)      # the reference's `.cuda()` at util.py:233
    B = colors.shape[0]
    n = colors.shape[1] // 3
    out = torch.zeros(B, 3, PANO_H, PANO_W, dtype=dirs.dtype, device=dirs.device)
    for k in range(n):
        d = torch.matmul(dirs[:, 3 * k:3 * k + 3], xyz).view(-1, PANO_H, PANO_W)
        out = out + colors[:, 3 * k:3 * k + 3][:, :, None, None] * torch.exp((d - 1) / sizes[:, k].view(-1, 1, 1))[:, None, :, :]
    return out
