"""The adjoint kernels of csrc/gp_bwd.cu executed on the CPU: the file is written in a restricted CUDA subset (1-D launches, no shared
memory / warp intrinsics, atomics only) so that `g++ -DEML_EMULATE` compiles the SAME source -- kernels and C-ABI wrappers -- into a
host library whose launches are loops over (block, thread).  That checks the index arithmetic, strides, padding handling and the
reduction pattern of every kernel against the torch formulas `gp_train.py` was validated with (tests/test_gp_train_cpu.py), on a box
without a GPU.  What emulation cannot show (launch limits, atomics under real concurrency) is covered by tests/test_gp_train_gpu.py."""
import ctypes
import os
import subprocess
from ctypes import c_double, c_float, c_int, c_long, c_void_p

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "emlight_b200", "csrc", "gp_bwd.cu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libgp_bwd_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DEML_EMULATE", "-x", "c++", SRC, "-o", out])
    lib = ctypes.CDLL(out)
    from emlight_b200 import _lib
    for name in ("eml_col2im_lut", "eml_act_bwd", "eml_bias_act_bwd", "eml_spade_bwd", "eml_bn_free_bwd", "eml_instance_norm_bwd",
                 "eml_upsample2_bwd", "eml_tanh_nchw_bwd", "eml_pool2d_bwd", "eml_loss_seed", "eml_im2col_lut_bf16_t", "eml_col2im_csr"):
        fn = getattr(lib, name + "_emu")
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]          # the emulated entry points have the product's signatures
    return lib


def P(t):
    return None if t is None else c_void_p(t.data_ptr())


def _slope(u, act):
    if act == 1:
        return (u > 0).float()
    if act == 2:
        return torch.where(u > 0, 1.0, 0.2)
    return torch.ones_like(u)


@pytest.mark.parametrize("kind,h,w,stride,C", [("sphere", 8, 16, 1, 5), ("sphere", 8, 16, 2, 8), ("conv", 9, 7, 2, 3), ("conv", 6, 6, 1, 4)])
def test_col2im_is_the_adjoint_of_the_gather(emu, kind, h, w, stride, C):
    from emlight_b200.genprojector import _conv_lut, _sphere_lut
    idx, wgt, ho, wo = (_sphere_lut if kind == "sphere" else _conv_lut)(h, w, stride)
    idx, wgt = torch.from_numpy(idx).contiguous(), torch.from_numpy(wgt).contiguous()
    B, Cp = 2, (C + 3) & ~3
    gen = torch.Generator().manual_seed(h * w + C)
    dA = torch.randn(B * ho * wo, 9 * Cp, generator=gen)
    dx = torch.zeros(B, h * w, Cp)
    assert emu.eml_col2im_lut_emu(P(dA), Cp, P(idx), P(wgt), P(dx), Cp, B, ho * wo, h * w, None) == 0
    want = torch.zeros(B, h * w, Cp)
    d3 = dA.reshape(B, ho * wo * 9, Cp)
    for t in range(4):
        want.index_add_(1, idx[:, :, t].reshape(-1).clamp_min(0).long(), d3 * wgt[:, :, t].reshape(1, -1, 1))
    assert float((dx - want).abs().max()) <= 1e-5 * float(want.abs().max())
    # adjoint identity <gather(x), dA> == <x, col2im(dA)>
    x = torch.randn(B, h * w, Cp, generator=gen)
    A = torch.zeros(B, ho * wo * 9, Cp)
    for t in range(4):
        A += x[:, idx[:, :, t].reshape(-1).clamp_min(0).long()] * wgt[:, :, t].reshape(1, -1, 1)
    assert abs(float((A * d3).sum()) - float((x * dx).sum())) <= 1e-4 * float((A * d3).abs().sum())
    # gather form over the inverted table: same result, every element written (start from garbage), deterministic
    from emlight_b200 import gp_ops
    offs, src, wv = gp_ops.lut_csr((idx, wgt, ho, wo), h * w)
    assert int(offs[-1]) == int(((idx >= 0) & (wgt != 0)).sum()) and offs.dtype == torch.int32
    dx2 = torch.full((B, h * w, Cp), 9.0)
    assert emu.eml_col2im_csr_emu(P(dA), Cp, P(offs), P(src), P(wv), P(dx2), Cp, B, ho * wo, h * w, None) == 0
    assert float((dx2 - want).abs().max()) <= 1e-5 * float(want.abs().max())
    dx3 = torch.zeros(B, h * w, Cp)
    assert emu.eml_col2im_csr_emu(P(dA), Cp, P(offs), P(src), P(wv), P(dx3), Cp, B, ho * wo, h * w, None) == 0 and torch.equal(dx2, dx3)
    assert emu.eml_col2im_lut_emu(P(dA), 6, P(idx), P(wgt), P(dx), 8, B, ho * wo, h * w, None) < 0       # Cp must be a multiple of 4
    assert emu.eml_col2im_lut_emu(None, Cp, P(idx), P(wgt), P(dx), Cp, B, ho * wo, h * w, None) < 0


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("M,C,pitch", [(37, 5, 8), (1000, 3, 4), (64, 130, 132)])
def test_act_bwd_and_bias_act_bwd(emu, act, M, C, pitch):
    gen = torch.Generator().manual_seed(M + C + act)
    x = torch.randn(M, pitch, generator=gen)
    bias = torch.randn(C, generator=gen)
    g = torch.randn(M, pitch, generator=gen)
    dx = g.clone()
    sums = torch.zeros(C, dtype=torch.float64)
    assert emu.eml_act_bwd_emu(P(dx), pitch, P(x), pitch, P(bias), act, M, C, P(sums), None) == 0
    want = g[:, :C] * _slope(x[:, :C] + bias, act)
    assert torch.allclose(dx[:, :C], want, atol=1e-6) and torch.equal(dx[:, C:], g[:, C:])          # padding untouched
    assert torch.allclose(sums, want.double().sum(0), rtol=1e-6, atol=1e-6)
    out = torch.zeros(M, pitch)
    u = x[:, :C] + bias
    out[:, :C] = F.relu(u) if act == 1 else F.leaky_relu(u, 0.2) if act == 2 else u
    dx2 = torch.full((M, pitch), 7.0)
    sums2 = torch.zeros(C, dtype=torch.float64)
    assert emu.eml_bias_act_bwd_emu(P(g), pitch, P(out), pitch, act, P(dx2), pitch, M, C, P(sums2), None) == 0
    assert torch.allclose(dx2[:, :C], want, atol=1e-6) and bool((dx2[:, C:] == 7.0).all())
    assert torch.allclose(sums2, want.double().sum(0), rtol=1e-6, atol=1e-6)
    assert emu.eml_act_bwd_emu(P(dx), pitch, P(x), pitch, P(bias), 3, M, C, P(sums), None) < 0


@pytest.mark.parametrize("leaky", [0, 1])
@pytest.mark.parametrize("M,C", [(50, 6), (700, 3), (32, 100)])
def test_spade_bwd_and_bn_free_bwd_match_autograd(emu, leaky, M, C):
    gen = torch.Generator().manual_seed(M * 3 + C + leaky)
    pitch, gbp = (C + 3) & ~3, (2 * C + 3) & ~3
    x = torch.randn(M, pitch, generator=gen) * 2 + 1
    gb = torch.randn(M, gbp, generator=gen) * 0.5
    bg, bb = torch.randn(C, generator=gen) * 0.1, torch.randn(C, generator=gen) * 0.1
    g = torch.randn(M, pitch, generator=gen)
    # reference: batch-statistic BatchNorm (no affine) + modulation (+ LeakyReLU) through autograd
    xr, gbr, bgr, bbr = x[:, :C].clone().requires_grad_(True), gb.clone().requires_grad_(True), bg.clone().requires_grad_(True), bb.clone().requires_grad_(True)
    mean = xr.mean(0)
    var = xr.var(0, unbiased=False)
    inv = torch.rsqrt(var + 1e-5)
    y = (xr - mean) * inv * (1 + gbr[:, :C] + bgr) + gbr[:, C:2 * C] + bbr
    out_ref = F.leaky_relu(y, 0.2) if leaky else y
    out_ref.backward(g[:, :C])
    out = F.pad(out_ref.detach(), (0, pitch - C))
    d_gb = torch.zeros(M, gbp)
    d_xhat = torch.zeros(M, pitch)
    sums = torch.zeros(4, C, dtype=torch.float64)
    m, i = mean.detach().contiguous(), inv.detach().contiguous()
    assert emu.eml_spade_bwd_emu(P(g), pitch, P(out), pitch, P(x), pitch, P(m), P(i), P(gb), gbp, P(bg), P(d_gb), P(d_xhat), pitch, M, C,
                                 leaky, P(sums), None) == 0
    assert torch.allclose(d_gb[:, :2 * C], gbr.grad[:, :2 * C], atol=1e-5)
    assert torch.allclose(sums[0].float(), bgr.grad, rtol=1e-4, atol=1e-4) and torch.allclose(sums[1].float(), bbr.grad, rtol=1e-4, atol=1e-4)
    dx = torch.zeros(M, pitch)
    assert emu.eml_bn_free_bwd_emu(P(d_xhat), pitch, P(x), pitch, P(m), P(i), P(sums[2:].contiguous()), float(M), P(dx), pitch, M, C, None) == 0
    assert float((dx[:, :C] - xr.grad).abs().max()) <= 1e-4 * float(xr.grad.abs().max()) + 1e-6
    # running-statistics mode: dx = inv_std * d_xhat
    dx2 = torch.zeros(M, pitch)
    assert emu.eml_bn_free_bwd_emu(P(d_xhat), pitch, None, 0, None, P(i), None, 0.0, P(dx2), pitch, M, C, None) == 0
    assert torch.allclose(dx2[:, :C], d_xhat[:, :C] * i, atol=1e-6)
    assert emu.eml_spade_bwd_emu(P(g), pitch, P(out), pitch, P(x), pitch, P(m), P(i), P(gb), C, P(bg), P(d_gb), P(d_xhat), pitch, M, C,
                                 leaky, P(sums), None) < 0                                            # gb pitch must hold gamma | beta


@pytest.mark.parametrize("leaky", [0, 1])
@pytest.mark.parametrize("B,HW,C", [(2, 48, 5), (3, 7, 8), (1, 300, 3)])
def test_instance_norm_bwd_matches_autograd(emu, leaky, B, HW, C):
    gen = torch.Generator().manual_seed(B * HW + C + leaky)
    pitch = (C + 3) & ~3
    raw = torch.randn(B, HW, pitch, generator=gen) * 3 + 0.5
    g = torch.randn(B, HW, pitch, generator=gen)
    xr = raw[..., :C].clone().requires_grad_(True)
    y = F.instance_norm(xr.permute(0, 2, 1).reshape(B, C, HW, 1), eps=1e-5).reshape(B, C, HW).permute(0, 2, 1)
    out_ref = F.leaky_relu(y, 0.2) if leaky else y
    out_ref.backward(g[..., :C])
    out = F.pad(out_ref.detach(), (0, pitch - C)).contiguous()
    sums = torch.zeros(B, 4, C, dtype=torch.float64)
    dx = torch.zeros(B, HW, pitch)
    assert emu.eml_instance_norm_bwd_emu(P(g), pitch, P(out), pitch, P(raw), pitch, B, HW, C, 1e-5, leaky, P(sums), P(dx), pitch, None) == 0
    assert float((dx[..., :C] - xr.grad).abs().max()) <= 2e-4 * float(xr.grad.abs().max()) + 1e-6


def test_upsample2_and_tanh_bwd_match_autograd(emu):
    gen = torch.Generator().manual_seed(4)
    B, H, W, C, pitch = 2, 3, 5, 6, 8
    x = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
    y = F.interpolate(x, scale_factor=2)
    g = torch.randn(B, 2 * H, 2 * W, pitch, generator=gen)
    y.backward(g[..., :C].permute(0, 3, 1, 2))
    dx = torch.zeros(B, H, W, pitch)
    assert emu.eml_upsample2_bwd_emu(P(g), pitch, P(dx), pitch, B, H, W, C, None) == 0
    assert torch.allclose(dx[..., :C].permute(0, 3, 1, 2), x.grad, atol=1e-6) and not dx[..., C:].any()
    # (tanh(raw + bias) + 1) * 25 -> NCHW
    C, pitch = 3, 4
    raw = torch.randn(B, H, W, C, generator=gen, requires_grad=True)
    bias = torch.randn(C, generator=gen, requires_grad=True)
    out = ((torch.tanh(raw + bias) + 1) * 25.0).permute(0, 3, 1, 2).contiguous()
    gn = torch.randn(B, C, H, W, generator=gen)
    out.backward(gn)
    d_raw = torch.zeros(B, H, W, pitch)
    sums = torch.zeros(C, dtype=torch.float64)
    assert emu.eml_tanh_nchw_bwd_emu(P(gn), P(out.detach()), 25.0, P(d_raw), pitch, B, H * W, C, P(sums), None) == 0
    assert torch.allclose(d_raw[..., :C], raw.grad, rtol=1e-4, atol=1e-5) and torch.allclose(sums.float(), bias.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("Hi,Wi", [(8, 12), (7, 9), (1, 1), (2, 5)])
def test_avg_pool_bwd_matches_autograd(emu, Hi, Wi):
    gen = torch.Generator().manual_seed(Hi * Wi)
    B, C, pitch = 2, 6, 8
    x = torch.randn(B, C, Hi, Wi, generator=gen, requires_grad=True)
    y = F.avg_pool2d(x, kernel_size=3, stride=2, padding=1, count_include_pad=False)
    Ho, Wo = y.shape[2], y.shape[3]
    g = torch.randn(B, Ho, Wo, pitch, generator=gen)
    y.backward(g[..., :C].permute(0, 3, 1, 2))
    dx = torch.zeros(B, Hi, Wi, pitch)
    assert emu.eml_pool2d_bwd_emu(P(g), pitch, None, 0, P(dx), pitch, Hi, Wi, C, B, 0, None) == 0
    assert torch.allclose(dx[..., :C].permute(0, 3, 1, 2), x.grad, atol=1e-6)


def test_max_pool_bwd_matches_autograd_including_ties(emu):
    gen = torch.Generator().manual_seed(8)
    B, C, Hi, Wi, pitch = 2, 5, 6, 8, 8
    x = torch.randn(B, C, Hi, Wi, generator=gen)
    x = torch.relu(x)                                    # many exact ties at 0, like VGG's post-ReLU maps
    x[0, 0, 0:2, 0:2] = 1.5                              # a 4-way tie at a positive value
    xr = x.clone().requires_grad_(True)
    y = F.max_pool2d(xr, 2, 2)
    g = torch.randn(B, Hi // 2, Wi // 2, pitch, generator=gen)
    y.backward(g[..., :C].permute(0, 3, 1, 2))
    xn = F.pad(x.permute(0, 2, 3, 1), (0, pitch - C)).contiguous()
    dx = torch.zeros(B, Hi, Wi, pitch)
    assert emu.eml_pool2d_bwd_emu(P(g), pitch, P(xn), pitch, P(dx), pitch, Hi, Wi, C, B, 1, None) == 0
    assert torch.equal(dx[..., :C].permute(0, 3, 1, 2), xr.grad)
    assert emu.eml_pool2d_bwd_emu(P(g), pitch, P(xn), pitch, P(dx), pitch, 5, Wi, C, B, 1, None) < 0        # odd size


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_loss_seeds_match_autograd_of_the_reductions(emu, mode):
    gen = torch.Generator().manual_seed(20 + mode)
    M, C, pitch = 60, 3, 4
    a = torch.randn(M, pitch, generator=gen) * 1.5
    b = torch.randn(M, pitch, generator=gen)
    mask = (torch.rand(M, generator=gen) > 0.5).float()
    if mode == 5:
        a[7, :C] = 0.0                                   # a zero vector: ATen clamps the norm
    ar = a[:, :C].clone().requires_grad_(True)
    bb = b[:, :C]
    if mode == 0:
        v = ar.sum()
    elif mode == 1:
        v = torch.clamp(ar - 1, max=0).sum()
    elif mode == 2:
        v = torch.clamp(-ar - 1, max=0).sum()
    elif mode == 3:
        v = (ar - bb).abs().sum()
    elif mode == 4:
        v = ((ar - bb).abs() * (mask + (1 - mask) * 50)[:, None]).sum()
    else:
        v = (1 - F.cosine_similarity(ar, bb, dim=1, eps=1e-20)).sum()
    (v * 0.37).backward()
    da = torch.zeros(M, pitch)
    assert emu.eml_loss_seed_emu(P(a), pitch, P(b), pitch, P(mask), M, C, mode, 0.74, P(torch.tensor([0.5])), P(da), pitch, None) == 0
    want = ar.grad
    ok = torch.isfinite(want).all(1)
    assert bool(ok.sum() >= M - 1)
    assert float((da[:, :C][ok] - want[ok]).abs().max()) <= 1e-4 * float(want[ok].abs().max())
    assert emu.eml_loss_seed_emu(P(a), pitch, None, pitch, P(mask), M, C, 3, 1.0, None, P(da), pitch, None) < 0


@pytest.mark.parametrize("kind,h,w,stride,C,act", [("sphere", 8, 16, 1, 5, 2), ("sphere", 8, 16, 2, 8, 0), ("conv", 9, 7, 2, 3, 1)])
def test_transposed_bf16_im2col_matches_the_gather(emu, kind, h, w, stride, C, act):
    """A^T in bf16 hi + lo == the fp32 gather to ~2^-16 relative; hi is the round-to-nearest-even bf16 of the value (the split of
    eml_split_bf16); padding rows / columns stay zero."""
    from emlight_b200.genprojector import _conv_lut, _sphere_lut
    idx, wgt, ho, wo = (_sphere_lut if kind == "sphere" else _conv_lut)(h, w, stride)
    idx, wgt = torch.from_numpy(idx).contiguous(), torch.from_numpy(wgt).contiguous()
    B, Cp = 2, (C + 3) & ~3
    gen = torch.Generator().manual_seed(h + w + C)
    x = torch.randn(B, h * w, Cp, generator=gen)
    bias = torch.randn(C, generator=gen)
    M = B * ho * wo
    Mp = (M + 63) // 64 * 64
    hi = torch.zeros(9 * Cp, Mp, dtype=torch.bfloat16)
    lo = torch.zeros(9 * Cp, Mp, dtype=torch.bfloat16)
    assert emu.eml_im2col_lut_bf16_t_emu(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(hi), P(lo), Mp, B, ho * wo, h * w, None) == 0
    u = x[..., :C] + bias
    u = F.relu(u) if act == 1 else F.leaky_relu(u, 0.2) if act == 2 else u
    u = F.pad(u, (0, Cp - C))
    A = torch.zeros(B, ho * wo * 9, Cp)
    for t in range(4):
        A += u[:, idx[:, :, t].reshape(-1).clamp_min(0).long()] * wgt[:, :, t].reshape(1, -1, 1)
    At = A.reshape(M, 9 * Cp).t()                                       # (9 Cp, M)
    got = hi.float() + lo.float()
    assert float((got[:, :M] - At).abs().max()) <= 2.0 ** -15 * float(At.abs().max())
    assert torch.equal(hi[:, :M], At.to(torch.bfloat16)) or float((hi[:, :M].float() - At).abs().max()) <= 2.0 ** -8 * float(At.abs().max())
    assert not got[:, M:].any()
    only_hi = torch.zeros_like(hi)
    assert emu.eml_im2col_lut_bf16_t_emu(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(only_hi), None, Mp, B, ho * wo, h * w, None) == 0
    assert torch.equal(only_hi, hi)
    assert emu.eml_im2col_lut_bf16_t_emu(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(hi), P(lo), M - 1, B, ho * wo, h * w, None) < 0


def test_product_library_validates_the_same_arguments(lib):
    """The real (CUDA) entry points reject bad arguments before any launch -- no GPU needed."""
    z = ctypes.create_string_buffer(64)
    a = ctypes.cast(z, c_void_p)
    assert lib.eml_col2im_lut(a, 6, a, a, a, 8, 1, 4, 4, None) < 0
    assert lib.eml_col2im_lut(None, 4, a, a, a, 4, 1, 4, 4, None) < 0
    assert lib.eml_col2im_csr(a, 4, None, a, a, a, 4, 1, 4, 4, None) < 0
    assert lib.eml_act_bwd(a, 4, a, 4, None, 5, 4, 4, None, None) < 0
    assert lib.eml_bias_act_bwd(a, 2, a, 4, 1, a, 4, 4, 4, None, None) < 0
    assert lib.eml_spade_bwd(a, 4, a, 4, a, 4, a, a, a, 4, None, a, a, 4, 4, 4, 0, a, None) < 0
    assert lib.eml_bn_free_bwd(a, 4, a, 4, a, a, a, 0.0, a, 4, 4, 4, None) < 0
    assert lib.eml_instance_norm_bwd(a, 4, a, 4, a, 4, 0, 4, 4, 1e-5, 0, a, a, 4, None) < 0
    assert lib.eml_upsample2_bwd(a, 2, a, 4, 1, 2, 2, 4, None) < 0
    assert lib.eml_tanh_nchw_bwd(a, a, 0.0, a, 4, 1, 4, 3, None, None) < 0
    assert lib.eml_pool2d_bwd(a, 4, None, 4, a, 4, 4, 4, 4, 1, 1, None) < 0
    assert lib.eml_loss_seed(a, 4, None, 4, None, 4, 4, 4, 1.0, None, a, 4, None) < 0
    assert np.isfinite(1.0) and c_double and c_float and c_int and c_long
