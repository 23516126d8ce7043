"""GPU: fused Sinkhorn forward+backward through the C ABI vs golden (reference geomloss/gmloss) and the oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import sinkhorn_oracle as SO

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _run(E, cuda, x, y, geometry=None, **kw):
    xs = torch.from_numpy(x).to(cuda).requires_grad_()
    ys = torch.from_numpy(y).to(cuda)
    if geometry is None:
        L = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=x.shape[0], **kw)(xs, ys)
    else:
        L = E.GMSamplesLoss("sinkhorn", p=2, blur=.025, **kw)(xs, ys, geometry)
    L.sum().backward()
    return L.detach().cpu().numpy(), xs.grad.cpu().numpy()


def test_sinkhorn_matches_reference_golden(cuda):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "sinkhorn.npz"))
    loss, grad = _run(E, cuda, g["geomloss_x"], g["geomloss_y"])
    assert loss.shape == (4,)
    assert np.abs(loss - g["geomloss_loss"]).max() <= RTOL * np.abs(g["geomloss_loss"]).max()
    assert np.abs(grad - g["geomloss_grad"]).max() <= RTOL * np.abs(g["geomloss_grad"]).max()
    loss, grad = _run(E, cuda, g["gmloss_x"], g["gmloss_y"], geometry=g["gmloss_geometry"])
    assert np.abs(loss - g["gmloss_loss"]).max() <= RTOL * np.abs(g["gmloss_loss"]).max()
    assert np.abs(grad - g["gmloss_grad"]).max() <= RTOL * np.abs(g["gmloss_grad"]).max()


@pytest.mark.parametrize("B,N", [(1, 96), (5, 128), (64, 128), (3, 8), (2, 160)])
def test_sinkhorn_matches_oracle(cuda, B, N):
    import emlight_b200 as E
    rng = np.random.default_rng(B * 1000 + N)
    x = (0.3 * rng.standard_normal((B, N, 1))).astype(np.float32)
    y = rng.random((B, N, 1)).astype(np.float32); y /= y.sum(1, keepdims=True)
    loss, grad = _run(E, cuda, x, y)
    lref, gref = SO.sinkhorn_loss(x, y, dtype=np.float64)
    assert np.abs(loss - lref).max() <= RTOL * np.abs(lref).max()
    assert np.abs(grad[:, :, 0] - gref).max() <= RTOL * np.abs(gref).max()


def test_sinkhorn_properties_full_batch(cuda):
    """BASELINE sizes (B=256, N=128): S(a,a)=0 with zero gradient; upstream gradient scaling; fixed diameter."""
    import emlight_b200 as E
    B, N = 256, 128
    gen = torch.Generator().manual_seed(11)
    y = torch.softmax(3 * torch.randn(B, N, generator=gen), 1).view(B, N, 1)
    ys = y.to(cuda)
    xs = y.clone().to(cuda).requires_grad_()
    L = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
    v = L(xs, ys)
    v.sum().backward()
    assert v.abs().max() < 1e-6 and xs.grad.abs().max() < 1e-6
    x2 = (y + 0.05 * torch.randn(B, N, 1, generator=gen)).to(cuda).requires_grad_()
    v2 = L(x2, ys)
    assert (v2 > 0).all()
    w = torch.rand(B, generator=gen).to(cuda)
    (v2 * w).sum().backward()
    g_w = x2.grad.clone(); x2.grad = None
    L(x2, ys).sum().backward()
    assert (g_w - x2.grad * w.view(B, 1, 1)).abs().max() <= 1e-6 * g_w.abs().max() + 1e-12
    # the eps schedule depends on the batch-wide diameter (sinkhorn_divergence.py:28-31): pinning it makes samples independent
    Lf = E.SamplesLoss("sinkhorn", p=2, blur=.025, diameter=1.0)
    a = Lf(x2.detach()[:7], ys[:7]); b = Lf(x2.detach(), ys)[:7]
    assert torch.equal(a, b)
