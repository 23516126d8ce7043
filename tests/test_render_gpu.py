"""GPU: SG panorama render through the C ABI vs the CPU oracle and the reference-generated golden vectors."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import render_oracle as RO
from oracle.make_golden import render_go

pytestmark = pytest.mark.gpu
RTOL = 1e-3      # north-star tolerance: 1e-3 relative (to the panorama's max), fp32


def test_render_matches_golden_and_oracle(cuda):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "render.npz"))
    for N in (96, 128):
        d, s, c = (torch.from_numpy(g["%s_%d" % (k, N)]).to(cuda) for k in ("dirs", "sizes", "colors"))
        pano = E.convert_to_panorama(d, s, c)
        assert pano.shape == (d.shape[0], 3, 128, 256) and pano.dtype == torch.float32
        ref = g["pano_%d" % N]
        assert np.abs(pano.cpu().numpy()[:, :, ::2, ::2] - ref).max() <= RTOL * np.abs(ref).max()
        full = RO.convert_to_panorama(g["dirs_%d" % N], g["sizes_%d" % N], g["colors_%d" % N])
        assert np.abs(pano.cpu().numpy() - full).max() <= RTOL * np.abs(full).max()


def test_render_backward_matches_golden(cuda):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "render.npz"))
    for N in (96, 128):
        d, s, c = (torch.from_numpy(g["%s_%d" % (k, N)]).to(cuda).requires_grad_() for k in ("dirs", "sizes", "colors"))
        pano = E.convert_to_panorama(d, s, c)
        go = torch.from_numpy(render_go(d.shape[0])).to(cuda)
        (pano * go).sum().backward()
        for t, key in ((d, "gdirs"), (s, "gsizes"), (c, "gcolors")):
            r = g["%s_%d" % (key, N)]
            assert np.abs(t.grad.cpu().numpy() - r).max() <= 2e-3 * np.abs(r).max(), key


def test_render_sharp_lobes_train_config(cuda):
    """The configuration train.py:115-122 uses: shared Fibonacci anchors, size 0.0025, colours from the heads."""
    import emlight_b200 as E
    B, N = 3, 96
    rng = np.random.default_rng(0)
    dist = rng.random((B, N), dtype=np.float32); dist /= dist.sum(1, keepdims=True)
    inten = rng.random((B, 1), dtype=np.float32); rgb = rng.random((B, 3), dtype=np.float32) + 0.2
    amb = rng.random((B, 3), dtype=np.float32) * 0.1
    dirs = RO.sphere_points(N).astype(np.float32).reshape(1, -1).repeat(B, 0)
    sizes = np.full((B, N), 0.0025, np.float32)
    cols = RO.compose_colors(dist, inten, rgb)
    ref = RO.convert_to_panorama(dirs, sizes, cols)
    t = lambda a: torch.from_numpy(a).to(cuda)
    got = E.convert_to_panorama(t(dirs), t(sizes), t(cols)).cpu().numpy()
    assert np.abs(got - ref).max() <= RTOL * np.abs(ref).max()
    fused = E.render_from_params(t(dist), t(inten), t(rgb)).cpu().numpy()
    assert np.abs(fused - ref).max() <= RTOL * np.abs(ref).max()
    fused_amb = E.render_from_params(t(dist), t(inten), t(rgb), ambient=t(amb)).cpu().numpy()
    assert np.abs(fused_amb - (ref + amb[:, :, None, None])).max() <= RTOL * np.abs(ref).max()


def test_render_linearity_and_edge_cases(cuda):
    """Size-independent properties at the BASELINE batch (256): linear in colours, zero colours -> zero map, N=1."""
    import emlight_b200 as E
    B, N = 256, 128
    gen = torch.Generator(device="cpu").manual_seed(5)
    dirs = torch.from_numpy(RO.sphere_points(N)).float().view(1, -1).repeat(B, 1).to(cuda)
    sizes = torch.full((B, N), 0.0025, device=cuda)
    c1 = torch.rand(B, 3 * N, generator=gen).to(cuda); c2 = torch.rand(B, 3 * N, generator=gen).to(cuda)
    p1, p2 = E.convert_to_panorama(dirs, sizes, c1), E.convert_to_panorama(dirs, sizes, c2)
    p12 = E.convert_to_panorama(dirs, sizes, 2 * c1 + 3 * c2)
    assert (p12 - (2 * p1 + 3 * p2)).abs().max() <= 1e-5 * p12.abs().max()
    assert E.convert_to_panorama(dirs, sizes, torch.zeros_like(c1)).abs().max() == 0
    one = E.convert_to_panorama(torch.tensor([[0., 0., 1.]], device=cuda), torch.tensor([[0.5]], device=cuda),
                                torch.tensor([[1., 2., 3.]], device=cuda))
    z = RO.pixel_dirs()[2]
    assert np.abs(one.cpu().numpy()[0, 1] - 2 * np.exp((z - 1) / 0.5)).max() < 1e-4
    assert E.convert_to_panorama(dirs[:0], sizes[:0], c1[:0]).shape == (0, 3, 128, 256)


def test_genprojector_guide_matches_data_py(cuda):
    """GenProjector/data.py:86-102 restated with the oracle render: env = (pano(dist * intensity * 0.01 * rgb) + ambient / 32768) * alpha."""
    import emlight_b200 as E
    from oracle import render_oracle as RO
    B, N = 3, 128
    g = torch.Generator().manual_seed(8)
    dist = torch.softmax(3 * torch.randn(B, N, generator=g), 1)
    inten = torch.rand(B, generator=g) * 400 + 50
    rgb = torch.nn.functional.normalize(torch.rand(B, 3, generator=g) + 0.2, dim=1)
    amb = torch.rand(B, 3, generator=g) * 2000
    alpha = torch.rand(B, generator=g) + 0.2
    out = E.genprojector_guide(dist.to(cuda), inten.to(cuda), rgb.to(cuda), amb.to(cuda), alpha.to(cuda)).cpu().numpy()
    dirs = np.tile(RO.sphere_points(N).astype(np.float32).reshape(1, -1), (B, 1))
    cols = RO.compose_colors(dist.numpy(), inten.numpy().reshape(B, 1), rgb.numpy(), gain=0.01)
    ref = RO.convert_to_panorama(dirs, np.full((B, N), 0.0025, np.float32), cols)
    ref = (ref + (amb.numpy() / (128 * 256))[:, :, None, None]) * alpha.numpy()[:, None, None, None]
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 2e-3 * np.abs(ref).max()
    one = E.genprojector_guide(dist[:1].to(cuda), inten[:1].to(cuda), rgb[:1].to(cuda), amb[:1].to(cuda), float(alpha[0])).cpu().numpy()
    assert np.abs(one - out[:1]).max() <= 1e-6 * np.abs(out).max()
