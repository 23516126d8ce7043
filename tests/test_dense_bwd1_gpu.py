"""GPU: the fused conv1-backward kernel (csrc/dense_bwd1.cu: tcgen05 dA = dN W1 with A in tensor memory, ReLU mask, in-place dS update through
TMA load / store, per-channel BatchNorm sums) and its helper entry points, through the C ABI, against the same algebra in float64 torch."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(dN, x, dS, w1, vec, ci):
    sc, sh, e, f, k1 = [v[:ci].double() for v in vec]
    xx = x[:, :ci].double()
    dA = dN.double() @ w1.double()                                   # (M, ci)
    z = torch.addcmul(sh.float(), x[:, :ci], sc.float())            # the kernel's fp32 fmaf decides the mask
    g = dA * (z > 0)
    out = dS.clone().double()
    out[:, :ci] += k1 * g
    return out, g.sum(0), (g * (e * xx + f)).sum(0)


@pytest.mark.parametrize("ci,tiles,precision,tol", [(24, 3, "bf16x3", 2e-5), (36, 400, "bf16x3", 2e-5), (150, 300, "bf16x3", 2e-5),
                                                    (204, 700, "bf16x3", 2e-5), (342, 310, "bf16x3", 2e-5), (108, 200, "bf16", 2e-2)])
def test_dense_bwd1_matches_float64(cuda, lib, ci, tiles, precision, tol):
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    M = 128 * tiles
    pitch = (ci + 12 + 7) & ~7
    gen = torch.Generator().manual_seed(ci)
    dN = torch.randn(M, 48, generator=gen).to(cuda)
    x = torch.randn(M, pitch, generator=gen).to(cuda)
    dS0 = torch.randn(M, pitch, generator=gen).to(cuda)
    w1 = (torch.randn(48, ci, generator=gen) / 7).to(cuda)
    sc, sh = torch.randn(ci, generator=gen).to(cuda), (0.3 * torch.randn(ci, generator=gen)).to(cuda)
    pa, pb = (0.5 + torch.rand(ci, generator=gen)).to(cuda), torch.randn(ci, generator=gen).to(cuda)
    mean, inv, gamma = torch.randn(ci, generator=gen).to(cuda), (0.5 + torch.rand(ci, generator=gen)).to(cuda), torch.randn(ci, generator=gen).to(cuda)
    vec = torch.empty(5, pitch, device=cuda)
    _lib.check(lib.eml_dense_bwd1_prep(P(sc), P(sh), P(pa), P(pb), P(mean), P(inv), P(gamma), ci, pitch, P(vec), st), "prep")
    assert torch.equal(vec[0, :ci], sc) and torch.equal(vec[1, :ci], sh)
    assert torch.allclose(vec[2, :ci], pa * inv) and torch.allclose(vec[3, :ci], (pb - mean) * inv) and torch.allclose(vec[4, :ci], gamma * inv)
    prec = _lib.PRECISIONS[precision]
    assert lib.eml_dense_bwd1_supported(ci, M, prec) == 1 and lib.eml_dense_bwd1_supported(ci, M + 1, prec) == 0
    wpack = torch.empty(lib.eml_dense_bwd1_wpack_bytes(ci), dtype=torch.uint8, device=cuda)
    _lib.check(lib.eml_dense_bwd1_pack(P(w1), P(wpack), ci, st), "pack")
    sums = torch.zeros(2, 352, dtype=torch.float64, device=cuda)
    dS = dS0.clone()
    _lib.check(lib.eml_dense_bwd1(P(dN), P(x), pitch, P(dS), pitch, P(wpack), P(vec), pitch, ci, M, P(sums), 352, prec, st), "eml_dense_bwd1")
    want, s1, s2 = _ref(dN, x, dS0, w1, vec, ci)
    scale = float((want[:, :ci] - dS0[:, :ci].double()).abs().max())
    assert float((dS[:, :ci].double() - want[:, :ci]).abs().max()) <= tol * scale + 1e-6 * float(want.abs().max())
    assert torch.equal(dS[:, ci:], dS0[:, ci:])                                  # channels past C_in are not touched (TMA store clips)
    assert float((sums[0, :ci] - s1).abs().max()) <= max(tol, 1e-5) * float(s1.abs().max()) + 1e-3
    assert float((sums[1, :ci] - s2).abs().max()) <= max(tol, 1e-5) * float(s2.abs().max()) + 1e-3
    # accumulate the deferred affine terms, then read a channel range through them
    coef = torch.zeros(2, pitch, device=cuda)
    dgb = torch.empty(2, ci, device=cuda)
    _lib.check(lib.eml_dense_bwd1_accum(P(sums), 352, P(vec), pitch, float(M), ci, P(coef[0]), P(coef[1]), P(dgb[0]), P(dgb[1]), st), "accum")
    assert torch.allclose(dgb[0].double(), sums[1, :ci], rtol=1e-6) and torch.allclose(dgb[1].double(), sums[0, :ci], rtol=1e-6)
    k1, e, f = vec[4, :ci].double(), vec[2, :ci].double(), vec[3, :ci].double()
    assert torch.allclose(coef[0, :ci].double(), -k1 / M * (sums[0, :ci] + f * sums[1, :ci]), rtol=1e-5, atol=1e-7)
    assert torch.allclose(coef[1, :ci].double(), -k1 / M * e * sums[1, :ci], rtol=1e-5, atol=1e-7)
    c0 = max(0, ci - 12)
    dy = torch.full((M, 16), 7.0, device=cuda)
    _lib.check(lib.eml_dense_bwd1_gather(P(dS), pitch, P(x), pitch, P(coef[0]), P(coef[1]), c0, 12, P(dy), 16, 16, M, st), "gather")
    ref = dS[:, c0:c0 + 12] + coef[0, c0:c0 + 12] + coef[1, c0:c0 + 12] * x[:, c0:c0 + 12]
    assert torch.allclose(dy[:, :12], ref, rtol=1e-6, atol=1e-6) and float(dy[:, 12:].abs().max()) == 0.0
    # full BatchNorm backward = masked term (kernel) + deferred affine terms: against autograd of relu(bn(pa x + pb)) . dA
    if tiles <= 3:
        xr = x[:, :ci].double().clone().requires_grad_(True)
        u = pa.double() * xr + pb.double()
        mu, var = u.mean(0), u.var(0, unbiased=False)
        eps = 1e-5
        zz = (u - mu) * torch.rsqrt(var + eps)
        beta = torch.randn(ci, generator=gen).to(cuda).double()
        a = torch.relu(gamma.double() * zz + beta)
        dA = (dN.double() @ w1.double()).detach()
        (a * dA).sum().backward()
        # the same through the kernel with this BN's own statistics
        inv2 = torch.rsqrt(var + eps).float()
        sc2, sh2 = (gamma * inv2 * pa), (gamma * inv2 * (pb - mu.float()) + beta.float())
        _lib.check(lib.eml_dense_bwd1_prep(P(sc2.contiguous()), P(sh2.contiguous()), P(pa), P(pb), P(mu.float().contiguous()), P(inv2.contiguous()),
                                           P(gamma), ci, pitch, P(vec), st), "prep")
        dS2 = torch.zeros(M, pitch, device=cuda)
        sums.zero_()
        _lib.check(lib.eml_dense_bwd1(P(dN), P(x), pitch, P(dS2), pitch, P(wpack), P(vec), pitch, ci, M, P(sums), 352, prec, st), "eml_dense_bwd1")
        coef.zero_()
        _lib.check(lib.eml_dense_bwd1_accum(P(sums), 352, P(vec), pitch, float(M), ci, P(coef[0]), P(coef[1]), P(dgb[0]), P(dgb[1]), st), "accum")
        _lib.check(lib.eml_dense_bwd1_gather(P(dS2), pitch, P(x), pitch, P(coef[0]), P(coef[1]), 0, ci, P(dS2), pitch, ci, M, st), "gather in place")
        du = dS2[:, :ci].double() * pa.double()                                    # dS holds d/du; autograd gave d/dx = pa * d/du
        assert float((du - xr.grad).abs().max()) <= 1e-3 * float(xr.grad.abs().max())


@pytest.mark.parametrize("C,M", [(204, 128 * 150), (330, 128 * 96), (342, 6400)])
def test_wgrad_1x1_tensor_core_ranges(cuda, lib, C, M):
    """eml_wgrad_1x1 (conv1's weight gradient, K = pixels on MN-major tcgen05 operands): C > 256 runs as two channel ranges."""
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(C)
    pitch = (C + 12 + 7) & ~7
    G = torch.randn(M, 48, generator=gen).to(cuda)
    x = torch.randn(M, pitch, generator=gen).to(cuda)
    sc, sh = torch.randn(C, generator=gen).to(cuda), (0.2 * torch.randn(C, generator=gen)).to(cuda)
    dW = torch.zeros(48, C, device=cuda)
    _lib.check(lib.eml_wgrad_1x1(P(G), 48, 48, P(x), pitch, C, P(sc), P(sh), 1, 0, 1, M, P(dW), M, _lib.PRECISIONS["bf16x3"], st), "eml_wgrad_1x1")
    a = torch.relu(torch.addcmul(sh, x[:, :C], sc)).double()
    want = G.double().t() @ a
    assert float((dW.double() - want).abs().max()) <= 1e-4 * float(want.abs().max())


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-4), ("bf16", 2e-2), ("fp32", 1e-5)])
@pytest.mark.parametrize("B,H,W,N", [(2, 5, 64, 12), (1, 37, 256, 12), (3, 4, 130, 12), (2, 3, 8, 12), (1, 70, 32, 16), (2, 9, 256, 7), (5, 1, 16, 12)])
def test_wgrad_3x3(cuda, lib, B, H, W, N, precision, tol):
    """eml_wgrad_3x3 (conv2's weight gradient, DenseNet.py:33-34 backward): dW[n,c,ky,kx] = sum_p dY[p,n] * N2[p+(ky-1,kx-1), c] with
    N2 = scale*b + shift inside the image and 0 outside.  Shapes cover every ring-slot case (H >= 4 rows), ragged widths (W = 130:
    a second 2-pixel tile), N < 12 with an unaligned tail, one-row images and bands (H = 70 > 32 rows)."""
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(B * 1000 + H * 10 + W)
    dY = torch.randn(B, H, W, 16, generator=gen).to(cuda)
    b = torch.randn(B, H, W, 48, generator=gen).to(cuda)
    sc, sh = torch.randn(48, generator=gen).to(cuda), (0.3 * torch.randn(48, generator=gen)).to(cuda)
    dW = torch.zeros(N, 48, 3, 3, device=cuda)
    _lib.check(lib.eml_wgrad_3x3(P(dY), 16, N, P(b), 48, 48, P(sc), P(sh), P(dW), B, H, W, _lib.PRECISIONS[precision], st), "eml_wgrad_3x3")
    n2 = torch.addcmul(sh, b, sc).double().permute(0, 3, 1, 2)                                   # (B, 48, H, W)
    w = torch.zeros(N, 48, 3, 3, dtype=torch.float64, device=cuda, requires_grad=True)
    y = torch.nn.functional.conv2d(n2, w, padding=1)
    y.backward(dY[..., :N].double().permute(0, 3, 1, 2))
    assert float((dW.double() - w.grad).abs().max()) <= tol * float(w.grad.abs().max())
    # accumulates into dW (the caller zeroes): a second call doubles it
    _lib.check(lib.eml_wgrad_3x3(P(dY), 16, N, P(b), 48, 48, P(sc), P(sh), P(dW), B, H, W, _lib.PRECISIONS[precision], st), "eml_wgrad_3x3")
    assert float((dW.double() - 2 * w.grad).abs().max()) <= 2 * tol * float(w.grad.abs().max())


@pytest.mark.parametrize("M,C,relu,pre,accumulate", [(1000, 48, 1, False, 0), (4099, 48, 0, True, 1), (777, 12, 1, True, 0), (2048, 128, 1, False, 1),
                                                      (1500, 132, 1, True, 0), (900, 342, 1, False, 0), (64, 216, 0, True, 1)])
def test_bn_backward_reduce_and_apply(cuda, lib, M, C, relu, pre, accumulate):
    """eml_bn_bwd_reduce / eml_bn_bwd_apply (BatchNorm2d under batch statistics + ReLU, DenseNet.py:32-41 backward) against fp64 autograd:
    narrow tensors (compact row-lane x quad mapping: C = 12, 48, 128), wide ones (block columns: 132, 216, 342 with a ragged last quad),
    with and without the composed pre-affine u = a*x + b, writing and accumulating."""
    from emlight_b200 import _lib
    P, st = _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(M + C)
    pitch = (C + 7) // 8 * 8
    x = torch.randn(M, pitch, generator=gen).to(cuda)
    g = torch.randn(M, pitch, generator=gen).to(cuda)
    pa = (torch.rand(C, generator=gen) + 0.5).to(cuda) if pre else None
    pb = torch.randn(C, generator=gen).to(cuda) if pre else None
    gamma, beta = (torch.rand(C, generator=gen) + 0.5).to(cuda), (0.3 * torch.randn(C, generator=gen)).to(cuda)
    xr = x[:, :C].double().clone().requires_grad_()
    u = xr * pa.double() + pb.double() if pre else xr
    mean, var = u.mean(0), u.var(0, unbiased=False)
    inv = torch.rsqrt(var + 1e-5)
    y = (u - mean) * inv * gamma.double() + beta.double()
    if relu:
        y = torch.relu(y)
    (y * g[:, :C].double()).sum().backward()
    mean32, inv32 = mean.detach().float().contiguous(), inv.detach().float().contiguous()
    sums = torch.zeros(2 * pitch, dtype=torch.float64, device=cuda)
    _lib.check(lib.eml_bn_bwd_reduce(P(g), pitch, P(x), pitch, P(pa), P(pb), P(mean32), P(inv32), P(gamma), P(beta), relu, 0, 0, 0, M, C,
                                     P(sums), pitch, st), "eml_bn_bwd_reduce")
    base = torch.randn(M, pitch, generator=gen).to(cuda)
    out = base.clone()
    _lib.check(lib.eml_bn_bwd_apply(P(g), pitch, P(x), pitch, P(pa), P(pb), P(mean32), P(inv32), P(gamma), P(beta), relu, 0, 0, 0, M, C,
                                    P(sums), pitch, P(out), pitch, accumulate, 1, st), "eml_bn_bwd_apply")
    want = xr.grad + (base[:, :C].double() if accumulate else 0)
    assert float((out[:, :C].double() - want).abs().max()) <= 2e-4 * float(xr.grad.abs().max())
    assert torch.equal(out[:, C:], base[:, C:])                                       # pitch padding untouched
    # the two sums are what the affine parameters' gradients are built from: d beta = S1, d gamma = S2
    yhat = ((u - mean) * inv).detach()
    gm = g[:, :C].double() * ((yhat * gamma.double() + beta.double() > 0) if relu else 1.0)
    assert float((sums[:C] - gm.sum(0)).abs().max()) <= 1e-4 * float(gm.sum(0).abs().max() + 1)
    assert float((sums[pitch:pitch + C] - (gm * yhat).sum(0)).abs().max()) <= 1e-4 * float((gm * yhat).sum(0).abs().max() + 1)
