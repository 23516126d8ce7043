"""GPU: SphereConv2D / SPADE generator through the drop-in modules vs the reference-generated golden and the CPU oracle."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import genprojector_oracle as GO

pytestmark = pytest.mark.gpu


def _opt(ngf):
    return argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                              num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)


def test_sphere_conv_matches_reference_golden(cuda):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "generator.npz"))
    sc = E.SphereConv2D(5, 7, stride=2).to(cuda)
    sc.load_state_dict({"weight": torch.from_numpy(g["sc_weight"]), "bias": torch.from_numpy(g["sc_bias"])})
    y = sc(torch.from_numpy(g["sc_x"]).to(cuda))
    assert y.shape == g["sc_y"].shape
    assert np.abs(y.cpu().numpy() - g["sc_y"]).max() <= 1e-3 * np.abs(g["sc_y"]).max()


@pytest.mark.parametrize("h,w,cin,cout,stride", [(4, 8, 16, 300, 1), (32, 64, 3, 128, 1), (16, 32, 130, 20, 2)])
def test_sphere_conv_matches_oracle(cuda, h, w, cin, cout, stride):
    import emlight_b200 as E
    gen = torch.Generator().manual_seed(h * 7 + cin)
    sc = E.SphereConv2D(cin, cout, stride=stride).to(cuda)
    wt = torch.randn(cout, cin, 3, 3, generator=gen) / np.sqrt(9 * cin)
    b = torch.randn(cout, generator=gen)
    sc.load_state_dict({"weight": wt, "bias": b})
    x = torch.randn(2, cin, h, w, generator=gen)
    ref = GO.sphere_conv(x, wt, b, stride)
    y = sc(x.to(cuda)).cpu()
    assert y.shape == ref.shape
    assert (y - ref).abs().max() <= 1e-3 * ref.abs().max()


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp32", 1e-3), ("bf16", 5e-2)])
def test_generator_matches_reference_golden(cuda, precision, tol):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "generator.npz"))
    ngf = int(g["ngf"])
    G = E.SPADEGenerator(_opt(ngf), precision=precision).to(cuda).eval()
    G.load_state_dict(GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=ngf))
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(1, 3, 160, 160, generator=gen)
    out = G(guide.to(cuda), crop.to(cuda))
    assert out.shape == (1, 3, 128, 256) and out.dtype == torch.float32
    err = np.abs(out.cpu().numpy()[:, :, ::2, ::2] - g["out"]).max() / 50.0          # output range is [0, 50]
    assert err <= tol, err


def test_generator_batch_and_errors(cuda):
    """Batch of 3 vs the CPU oracle on fresh inputs; training mode / CPU tensors fail loudly."""
    import emlight_b200 as E
    ngf = 8
    sd = GO.init_generator_state_dict(seed=5, ngf=ngf)
    G = E.SPADEGenerator(_opt(ngf)).to(cuda).eval()
    G.load_state_dict(sd)
    gen = torch.Generator().manual_seed(11)
    guide = torch.rand(3, 3, 128, 256, generator=gen) * 3
    crop = torch.rand(3, 3, 128, 128, generator=gen)
    with torch.no_grad():
        ref = GO.generator_forward(sd, guide, crop, ngf)
    out = G(guide.to(cuda), crop.to(cuda)).cpu()
    assert (out - ref).abs().max() / 50.0 <= 1e-3
    again = G(guide[1:2].to(cuda), crop[1:2].to(cuda)).cpu()
    assert torch.equal(again, out[1:2])                                              # per-sample independence, run-to-run identical
    with pytest.raises(RuntimeError, match="CUDA"):
        G(guide, crop)


def test_generator_train_mode_forward_matches_reference_golden(cuda):
    """.train() forward: batch-statistic BatchNorm inside SPADE (eml_channel_stats) with running-stat update and one spectral-norm
    power iteration per wrapped convolution, two consecutive steps vs the reference module (tests/golden/generator_train.npz)."""
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "generator_train.npz"))
    ngf = int(g["ngf"])
    G = E.SPADEGenerator(_opt(ngf)).to(cuda).train()
    G.load_state_dict(GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=ngf))
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(2, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(2, 3, 128, 128, generator=gen)
    out1 = G(guide.to(cuda), crop.to(cuda)).cpu()
    out2 = G((guide * 0.5).to(cuda), crop.to(cuda)).cpu()
    assert not out1.requires_grad                                                     # forward values only: no autograd graph
    assert np.abs(out1.numpy()[:, :, ::2, ::2] - g["out1"]).max() / 50.0 <= 1e-3
    assert np.abs(out2.numpy()[:, :, ::2, ::2] - g["out2"]).max() / 50.0 <= 1e-3
    sd = G.state_dict()
    for k in g.files:
        if k.startswith("buf_"):
            want, got = g[k], sd[k[4:]].cpu().numpy()
            assert np.abs(got - want).max() <= 1e-3 * max(1.0, np.abs(want).max()), k
    # eval afterwards uses the updated running statistics / vectors and is deterministic
    G.eval()
    a = G(guide.to(cuda), crop.to(cuda))
    assert torch.equal(a, G(guide.to(cuda), crop.to(cuda)))


def test_generator_and_discriminator_match_reference_golden_at_baseline_width(cuda):
    """ngf = ndf = 64 (K up to 9216 per output, 4x the accumulation length of the ngf = 16 fixture): outputs of the
    reference's own modules (oracle/make_golden_ngf64.py) vs the bf16x3 tcgen05 path, at the north-star tolerance 1e-3."""
    import emlight_b200 as E
    from test_discriminator_cpu import d_opt
    g = np.load(os.path.join(GOLDEN, "genprojector_w64.npz"))
    w = int(g["width"])
    G = E.SPADEGenerator(_opt(w)).to(cuda).eval()
    G.load_state_dict(GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=w))
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(1, 3, 128, 128, generator=gen)
    with torch.no_grad():
        out = G(guide.to(cuda), crop.to(cuda))
    err = np.abs(out.cpu().numpy()[:, :, ::4, ::4] - g["out"]).max() / 50.0
    assert err <= 1e-3, err
    assert abs(float(out.double().mean()) - float(g["out_mean"])) <= 1e-3 * float(g["out_mean"])
    del G
    D = E.MultiscaleDiscriminator(d_opt(w)).to(cuda).eval()
    D.load_state_dict(GO.init_discriminator_state_dict(int(g["sd_seed"]), w))
    fake = torch.rand(1, 3, 128, 256, generator=gen) * 50 * torch.rand(1, 1, 128, 256, generator=gen) ** 4
    with torch.no_grad():
        feats = D(torch.cat([guide, fake], 1).to(cuda))
    for i, fl in enumerate(feats):
        for j, f in enumerate(fl):
            want = g["d%d_%d" % (i, j)]
            got = f.cpu().numpy()[:, ::max(1, f.shape[1] // 8), ::2, ::2]
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1e-3 * np.abs(want).max(), (i, j, np.abs(got - want).max() / np.abs(want).max())


@pytest.mark.parametrize("B,h,w,C,stride,kind,act", [(3, 5, 9, 3, 1, "sphere", 0), (2, 16, 32, 64, 1, "sphere", 2), (2, 8, 16, 10, 2, "sphere", 1),
                                                      (1, 32, 64, 128, 1, "sphere", 1), (2, 12, 20, 36, 2, "conv", 2), (1, 4, 8, 1024, 1, "sphere", 0),
                                                      (2, 7, 11, 20, 1, "sphere", 1)])
def test_im2col_bf16_operand_equals_fp32_im2col(cuda, B, h, w, C, stride, kind, act):
    """eml_im2col_lut_bf16 (the GEMM operand the tensor-core path reads: bf16 hi | lo, row length Kp, zero K padding) against
    eml_im2col_lut (fp32, checked against the CPU restatement through the fp32 generator tests): hi == bf16(value) bit for bit, hi + lo
    reconstructs the value to 2^-16, pad columns are zero.  Shapes: one channel quad per tap (C = 3), ragged quads (C = 10), tiles that
    cross an image boundary (45 pixels per image, tiles of 32), the widest layer (C = 1024), regular-grid LUT with stride 2."""
    from emlight_b200 import _lib, genprojector as GP
    lib, P, st = _lib.load(), _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(C * 3 + h)
    Cp = (C + 3) // 4 * 4
    x = torch.randn(B, h, w, Cp, generator=gen).to(cuda)
    bias = torch.randn(C, generator=gen).to(cuda)
    idx, wgt, ho, wo = GP._LUTS.get(kind, h, w, stride, cuda)
    M, K = B * ho * wo, 9 * Cp
    Kp = (K + 63) // 64 * 64
    A = torch.full((M, K), float("nan"), device=cuda)
    _lib.check(lib.eml_im2col_lut(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(A), B, ho * wo, h * w, st), "eml_im2col_lut")
    hi = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device=cuda)
    lo = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.eml_im2col_lut_bf16(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(hi), P(lo), Kp, B, ho * wo, h * w, st), "eml_im2col_lut_bf16")
    assert torch.equal(hi[:, :K], A.to(torch.bfloat16))
    assert float((hi[:, :K].float() + lo[:, :K].float() - A).abs().max()) <= 2.0 ** -16 * float(A.abs().max())
    assert float(hi[:, K:].float().abs().sum()) == 0.0 and float(lo[:, K:].float().abs().sum()) == 0.0
    only = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device=cuda)
    _lib.check(lib.eml_im2col_lut_bf16(P(x), Cp, C, Cp, P(idx), P(wgt), P(bias), act, P(only), None, Kp, B, ho * wo, h * w, st), "eml_im2col_lut_bf16")
    assert torch.equal(only, hi)                                                                    # single-pass bf16 tier: A_lo = NULL
    if C == Cp:
        # what _conv_raw runs: the input transform once per value (eml_bias_act), then the bias-free / activation-free fast path of the
        # gather -- same operand, bit for bit (the blend sees identical values in identical order)
        xt = torch.full((B, h, w, Cp), float("nan"), device=cuda)
        _lib.check(lib.eml_bias_act(P(x), Cp, P(bias), act, P(xt), Cp, B * h * w, C, st), "eml_bias_act")
        want = x + bias
        want = torch.relu(want) if act == 1 else (torch.where(want > 0, want, 0.2 * want) if act == 2 else want)
        assert torch.equal(xt, want)
        h2 = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device=cuda)
        l2 = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device=cuda)
        _lib.check(lib.eml_im2col_lut_bf16(P(xt), Cp, C, Cp, P(idx), P(wgt), None, 0, P(h2), P(l2), Kp, B, ho * wo, h * w, st), "eml_im2col_lut_bf16 (plain)")
        assert torch.equal(h2, hi) and torch.equal(l2, lo)
        _lib.check(lib.eml_im2col_lut_bf16(P(xt), Cp, C, Cp, P(idx), P(wgt), None, 0, P(h2), None, Kp, B, ho * wo, h * w, st), "eml_im2col_lut_bf16 (plain)")
        assert torch.equal(h2, hi)
        # the weight-gradient GEMM's operand is the TRANSPOSE (rows = (tap, channel), columns = pixels padded to 64): the tiled kernel
        # writes exactly the transposed forward operand and leaves the pad columns to the caller
        Mp = (M + 63) // 64 * 64
        ht = torch.full((K, Mp), 3.0, dtype=torch.bfloat16, device=cuda)
        lt = torch.full((K, Mp), 3.0, dtype=torch.bfloat16, device=cuda)
        _lib.check(lib.eml_im2col_lut_bf16_t(P(xt), Cp, C, Cp, P(idx), P(wgt), None, 0, P(ht), P(lt), Mp, B, ho * wo, h * w, st), "eml_im2col_lut_bf16_t")
        assert torch.equal(ht[:, :M], hi[:, :K].t()) and torch.equal(lt[:, :M], lo[:, :K].t())
        assert bool((ht[:, M:] == 3.0).all()) and bool((lt[:, M:] == 3.0).all())
        from emlight_b200 import gp_ops
        gh, gl = gp_ops.im2col_t(x, B, h, w, C, (idx, wgt, ho, wo), bias, act)                      # the host wrapper: transform + pad zeroing
        assert torch.equal(gh[:, :M], hi[:, :K].t()) and torch.equal(gl[:, :M], lo[:, :K].t())
        assert float(gh[:, M:].float().abs().sum()) == 0.0 and float(gl[:, M:].float().abs().sum()) == 0.0


@pytest.mark.parametrize("O,K", [(64, 27), (1024, 9216), (3, 576), (128, 1152), (37, 1001)])
def test_spectral_norm_kernel_equals_the_tensor_formulation(cuda, O, K):
    """eml_spectral_norm == torch.nn.utils.spectral_norm's arithmetic (architecture.py:37-40): v <- normalize(W^T u), u <- normalize(W v),
    sigma = u . (W v) in training mode (buffers updated in place), stored vectors in eval mode."""
    from emlight_b200 import _lib
    import torch.nn.functional as F
    lib, P, st = _lib.load(), _lib.ptr, _lib.stream_ptr()
    gen = torch.Generator().manual_seed(O + K)
    W = torch.randn(O, K, generator=gen).to(cuda)
    u0 = F.normalize(torch.randn(O, generator=gen), dim=0).to(cuda)
    v0 = F.normalize(torch.randn(K, generator=gen), dim=0).to(cuda)
    Wd = W.double()
    v_ref = F.normalize(Wd.t() @ u0.double(), dim=0, eps=1e-12)
    u_ref = F.normalize(Wd @ v_ref, dim=0, eps=1e-12)
    s_ref = torch.dot(u_ref, Wd @ v_ref)
    u, v = u0.clone(), v0.clone()
    scratch = torch.empty(O + K + 1, device=cuda)
    _lib.check(lib.eml_spectral_norm(P(W), O, K, P(u), P(v), 1, 1e-12, P(scratch), P(scratch[O + K:]), st), "eml_spectral_norm")
    assert float((v.double() - v_ref).abs().max()) <= 2e-6 and float((u.double() - u_ref).abs().max()) <= 2e-6
    assert abs(float(scratch[O + K]) - float(s_ref)) <= 2e-6 * float(s_ref)
    # eval mode: buffers untouched, sigma from the stored vectors
    u2, v2 = u0.clone(), v0.clone()
    _lib.check(lib.eml_spectral_norm(P(W), O, K, P(u2), P(v2), 0, 1e-12, P(scratch), P(scratch[O + K:]), st), "eml_spectral_norm")
    assert torch.equal(u2, u0) and torch.equal(v2, v0)
    s_eval = torch.dot(u0.double(), Wd @ v0.double())
    assert abs(float(scratch[O + K]) - float(s_eval)) <= 1e-5 * max(1.0, abs(float(s_eval)))
    # a second run from the same state gives the same bits (fixed summation order)
    u3, v3 = u0.clone(), v0.clone()
    _lib.check(lib.eml_spectral_norm(P(W), O, K, P(u3), P(v3), 1, 1e-12, P(scratch), P(scratch[O + K:]), st), "eml_spectral_norm")
    assert torch.equal(u3, u) and torch.equal(v3, v)
