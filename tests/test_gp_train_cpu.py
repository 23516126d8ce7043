"""CPU check of the GenProjector training tape (`emlight_b200/gp_train.py`): every FORWARD device primitive of `gp_ops` is replaced by a
torch stand-in that restates the kernel's contract from include/emlight_b200.h, the ADJOINT kernels of csrc/gp_bwd.cu run as
themselves through their host emulation build (see tests/test_gp_bwd_emulated.py), and the gradients the tape produces are compared with
torch autograd through the oracle (`oracle/genprojector_oracle.py`, the reference's modules restated functionally).  This pins the
backward ALGEBRA (adjoint of the sampling-table gather, SPADE / batch-statistic BatchNorm / InstanceNorm / spectral-norm adjoints,
loss seeds, gradient routing through the [fake; real] batches); the same tape on the real kernels is `tests/test_gp_train_gpu.py`."""
import argparse

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import genprojector_oracle as GO


def _up4(n):
    return (n + 3) & ~3


# ------------------------------------------------------------------------------------------------- torch stand-ins for gp_ops
class SimPackedConv:
    def __init__(self, weight, precision):
        O, C = weight.shape[0], weight.shape[1]
        self.C, self.Cp, self.O = C, _up4(C), O
        K = 9 * self.Cp
        if precision != "fp32":
            K = (K + 63) & ~63
        wk = torch.zeros(O, K)
        wk[:, :9 * self.Cp].view(O, 9, self.Cp)[:, :, :C] = weight.detach().float().permute(0, 2, 3, 1).reshape(O, 9, C)
        self.wk, self.K = wk, K


def _act(u, act):
    return F.relu(u) if act == 1 else F.leaky_relu(u, 0.2) if act == 2 else u


def sim_im2col(x, B, H, W, C, lut, bias_in, act):
    idx, wgt, ho, wo = lut
    Cp = _up4(C)
    u = x[..., :C].reshape(B, H * W, C)
    if bias_in is not None:
        u = u + bias_in
    u = F.pad(_act(u, act), (0, Cp - C))
    A = torch.zeros(B, ho * wo, 9, Cp)
    for t in range(4):
        A += u[:, idx[:, :, t].clamp_min(0).long().reshape(-1)].reshape(B, ho * wo, 9, Cp) * wgt[:, :, t].reshape(1, ho * wo, 9, 1)
    return A.reshape(B * ho * wo, 9 * Cp)


def sim_conv_raw(x, B, H, W, pc, lut, bias_in, act, precision):
    _, _, ho, wo = lut
    A = sim_im2col(x, B, H, W, pc.C, lut, bias_in, act)
    out = A @ pc.wk[:, :9 * pc.Cp].t()
    return F.pad(out, (0, _up4(pc.O) - pc.O)).reshape(B, ho, wo, _up4(pc.O))


def sim_bias_act(raw, bias, act, M, C):
    out = torch.zeros_like(raw)
    out[..., :C] = _act(raw[..., :C] + (bias if bias is not None else 0), act)
    return out


def sim_pool(x, B, H, W, C, mode):
    xi = x[..., :C].permute(0, 3, 1, 2)
    y = F.avg_pool2d(xi, 3, 2, 1, count_include_pad=False) if mode == 0 else F.max_pool2d(xi, 2, 2)
    out = F.pad(y.permute(0, 2, 3, 1), (0, x.shape[-1] - C)).contiguous()
    return out, out.shape[1], out.shape[2]


def sim_nchw_to_nhwc(x, pitch):
    return F.pad(x.float().permute(0, 2, 3, 1), (0, pitch - x.shape[1])).contiguous()


def sim_loss_sum(mode, a, M, C, a_pitch, b=None, b_pitch=0, mask=None):
    a2 = a.reshape(M, a_pitch)[:, :C].double()
    b2 = b.reshape(M, b_pitch)[:, :C].double() if b is not None else None
    if mode == 0:
        v = a2.sum()
    elif mode == 1:
        v = torch.clamp(a2 - 1, max=0).sum()
    elif mode == 2:
        v = torch.clamp(-a2 - 1, max=0).sum()
    elif mode == 3:
        v = (a2 - b2).abs().sum()
    elif mode == 4:
        m = mask.reshape(M, 1).double()
        v = ((a2 - b2).abs() * (m + (1 - m) * 50)).sum()
    else:
        v = (1 - F.cosine_similarity(a2, b2, dim=1, eps=1e-20)).sum()
    return v.reshape(1)


def sim_instance_norm(raw, B, HW, C, lrelu):
    x = raw[..., :C]
    y = (x - x.mean((1, 2), keepdim=True)) * torch.rsqrt(x.var((1, 2), unbiased=False, keepdim=True) + 1e-5)
    out = torch.zeros_like(raw)
    out[..., :C] = F.leaky_relu(y, 0.2) if lrelu else y
    return out


def sim_channel_sums(x, M, C):
    x2 = x.reshape(M, -1)[:, :C].double()
    return torch.stack([x2.sum(0), (x2 * x2).sum(0)])


def sim_spade_modulate(x, mean, inv, gb, bg, bb, M, C, lrelu):
    y = (x[..., :C] - mean) * inv * (1 + gb[..., :C] + bg) + gb[..., C:2 * C] + bb
    return F.pad(F.leaky_relu(y, 0.2) if lrelu else y, (0, _up4(C) - C))


def sim_bias_residual(a, bias_a, r, bias_r, M, C):
    y = a[..., :C] + (bias_a if bias_a is not None else 0)
    if r is not None:
        y = y + r[..., :C] + (bias_r if bias_r is not None else 0)
    return F.pad(y, (0, _up4(C) - C))


def sim_resize_nearest(x, x_pitch, Hi, Wi, Ho, Wo, C, B, src_is_nchw, out_pitch):
    xi = x.reshape(B, C, Hi, Wi) if src_is_nchw else x.reshape(B, Hi, Wi, x_pitch)[..., :C].permute(0, 3, 1, 2)
    y = F.interpolate(xi.float(), size=(Ho, Wo), mode="nearest")
    return F.pad(y.permute(0, 2, 3, 1), (0, out_pitch - C)).contiguous()


def sim_resize_bilinear_nchw(x, Ho, Wo):
    y = F.interpolate(x.float(), size=(Ho, Wo), mode="bilinear", align_corners=False)
    return F.pad(y.permute(0, 2, 3, 1), (0, _up4(x.shape[1]) - x.shape[1])).contiguous()


def sim_tanh_to_nchw(raw, bias, B, H, W, C, scale):
    return ((torch.tanh(raw[..., :C] + bias) + 1) * scale).permute(0, 3, 1, 2).contiguous()


def sim_mm_nt_split(a_hi, a_lo, M, K, b, precision="bf16x3"):
    a = a_hi.float() + (a_lo.float() if a_lo is not None else 0)
    return a[:M, :K] @ b.t()


def build_emu(outdir):
    """csrc/gp_bwd.cu compiled for the host (EML_EMULATE): the adjoint KERNELS themselves run inside these tests."""
    import ctypes
    import os
    import subprocess
    from emlight_b200 import _lib
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "emlight_b200", "csrc", "gp_bwd.cu")
    out = os.path.join(str(outdir), "libgp_bwd_emu.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DEML_EMULATE", "-x", "c++", src, "-o", out])
    lib = ctypes.CDLL(out)
    fns = {}
    for name in ("eml_col2im_lut", "eml_act_bwd", "eml_bias_act_bwd", "eml_spade_bwd", "eml_bn_free_bwd", "eml_instance_norm_bwd",
                 "eml_upsample2_bwd", "eml_tanh_nchw_bwd", "eml_pool2d_bwd", "eml_loss_seed", "eml_im2col_lut_bf16_t", "eml_col2im_csr"):
        fn = getattr(lib, name + "_emu")
        fn.restype, fn.argtypes = _lib.SIGNATURES[name]
        fns[name] = fn
    return fns


def install_sims(gp_ops, emu_fns, setter=setattr):
    setter(gp_ops, "_fn", lambda name: emu_fns[name])
    setter(gp_ops, "_st", lambda: None)
    for name, fn in dict(PackedConv=SimPackedConv, conv_raw=sim_conv_raw, im2col=sim_im2col, bias_act=sim_bias_act, pool=sim_pool,
                         nchw_to_nhwc=sim_nchw_to_nhwc, loss_sum=sim_loss_sum, instance_norm=sim_instance_norm,
                         channel_sums=sim_channel_sums, spade_modulate=sim_spade_modulate, bias_residual=sim_bias_residual,
                         resize_nearest=sim_resize_nearest, resize_bilinear_nchw=sim_resize_bilinear_nchw,
                         tanh_to_nchw=sim_tanh_to_nchw, linear=lambda a, w, b: a @ w.t() + b,
                         mm_nt=lambda a, b, precision="bf16x3": a @ b.t(), mm_nt_split=sim_mm_nt_split).items():
        setter(gp_ops, name, fn)


@pytest.fixture(scope="module")
def emu_lib(tmp_path_factory):
    return build_emu(tmp_path_factory.mktemp("emu"))


@pytest.fixture()
def sim(monkeypatch, emu_lib):
    from emlight_b200 import gp_ops
    install_sims(gp_ops, emu_lib, monkeypatch.setattr)
    from emlight_b200 import gp_train
    return gp_train


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _leaf(tape, t):
    """Registers a plain tensor as a tape tensor whose gradient can be read back after tape.backward()."""
    box = {}

    def bwd():
        box["g"] = tape.take(t)

    tape.record(bwd)
    return box


# ------------------------------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("stride,act,cin,cout", [(1, 0, 5, 7), (2, 2, 3, 6), (1, 1, 8, 4)])
def test_conv_adjoint_matches_autograd(sim, stride, act, cin, cout):
    gt = sim
    from emlight_b200 import gp_ops
    gen = torch.Generator().manual_seed(cin * 10 + stride)
    B, H, W = 2, 8, 16
    x = torch.randn(B, cin, H, W, generator=gen)
    wt = torch.randn(cout, cin, 3, 3, generator=gen)
    bin_ = torch.nn.Parameter(torch.randn(cin, generator=gen))
    # oracle: SphereConv(act(x + bias_in))
    xr, wr, br = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), bin_.detach().clone().requires_grad_(True)
    ref = GO.sphere_conv(_act(xr + br.view(1, -1, 1, 1), act), wr, None, stride)
    gy = torch.randn(ref.shape, generator=gen)
    ref.backward(gy)
    tape = gt.Tape()
    xn = sim_nchw_to_nhwc(x, _up4(cin))
    bx = _leaf(tape, xn)
    got = {}
    raw, ho, wo = gt.conv(tape, xn, B, H, W, cin, wt, gp_ops.lut("sphere", H, W, stride, x.device), bin_.detach(), bin_, act, "bf16x3",
                          lambda dw: got.setdefault("dw", dw))
    assert _rel(raw[..., :cout].permute(0, 3, 1, 2), ref.detach()) < 1e-5
    tape.add(raw, sim_nchw_to_nhwc(gy, _up4(cout)))
    tape.backward()
    assert _rel(got["dw"], wr.grad) < 1e-4
    assert _rel(bx["g"][..., :cin].permute(0, 3, 1, 2), xr.grad) < 1e-4
    assert _rel(tape.param_grads[bin_], br.grad) < 1e-4


def test_instance_norm_pool_and_spectral_adjoints(sim):
    gt = sim
    gen = torch.Generator().manual_seed(3)
    B, H, W, C = 2, 8, 12, 5
    x = torch.randn(B, C, H, W, generator=gen)
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(F.avg_pool2d(F.leaky_relu(F.instance_norm(xr, eps=1e-5), 0.2), 3, 2, 1, count_include_pad=False), 2, 2)
    gy = torch.randn(ref.shape, generator=gen)
    ref.backward(gy)
    tape = gt.Tape()
    xn = sim_nchw_to_nhwc(x, 8)
    bx = _leaf(tape, xn)
    y = gt.instance_norm(tape, xn, B, H, W, C, lrelu=True)
    y, h, w = gt._pool_with_grad(tape, y, B, H, W, C, 0)
    y, h, w = gt._pool_with_grad(tape, y, B, h, w, C, 1)
    assert _rel(y[..., :C].permute(0, 3, 1, 2), ref.detach()) < 1e-5
    tape.add(y, sim_nchw_to_nhwc(gy, 8))
    tape.backward()
    assert _rel(bx["g"][..., :C].permute(0, 3, 1, 2), xr.grad) < 1e-4
    # spectral norm: W_orig / (u^T W v), u and v constants
    mod = argparse.Namespace(weight_orig=torch.nn.Parameter(torch.randn(6, 4, 3, 3, generator=gen)),
                             weight_u=F.normalize(torch.randn(6, generator=gen), dim=0), weight_v=F.normalize(torch.randn(36, generator=gen), dim=0))
    u0, v0 = mod.weight_u.clone(), mod.weight_v.clone()
    w_eff, sigma, u, v = gt.spectral_weight(mod, update=True)
    assert not torch.equal(mod.weight_u, u0) and not torch.equal(mod.weight_v, v0)           # one power iteration, written in place
    sd, upd = {"c.weight_orig": mod.weight_orig.detach().clone().requires_grad_(True), "c.weight_u": u0, "c.weight_v": v0}, {}
    w_ref = GO.sn_weight(sd, "c", upd)
    assert _rel(w_eff, w_ref.detach()) < 1e-6 and _rel(mod.weight_u, upd["c.weight_u"]) < 1e-6
    gw = torch.randn(w_ref.shape, generator=gen)
    w_ref.backward(gw)
    tape = gt.Tape()
    gt._spectral_backward(tape, mod, w_eff, sigma, u, v, gw)
    assert _rel(tape.param_grads[mod.weight_orig], sd["c.weight_orig"].grad) < 1e-4


# ------------------------------------------------------------------------------------------------- whole generator
def _g_opt(ngf):
    return argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                              num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)


def test_generator_backward_matches_oracle_autograd(sim):
    """Train-mode forward + backward of the whole SPADEGenerator through `SPADEGenerator.forward` (autograd opt-in) against autograd
    of the oracle's train-mode forward: every parameter gradient, the updated spectral-norm vectors and running statistics."""
    from emlight_b200.genprojector import SPADEGenerator
    ngf = 2
    G = SPADEGenerator(_g_opt(ngf), precision="bf16x3").train()
    sd0 = GO.init_generator_state_dict(seed=4, ngf=ngf)
    G.load_state_dict(sd0)
    G.autograd = True
    gen = torch.Generator().manual_seed(9)
    guide = torch.rand(2, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(2, 3, 96, 112, generator=gen)
    gout = torch.randn(2, 3, 128, 256, generator=gen)
    import emlight_b200._lib as L
    L_require = L.require_cuda
    L.require_cuda = lambda *a: None                     # the stand-ins run on CPU tensors; the product keeps the check
    try:
        out = G(guide, crop)
        assert out.requires_grad
        (out * gout).sum().backward()
    finally:
        L.require_cuda = L_require
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))
              else v.clone()) for k, v in sd0.items()}
    upd = {}
    ref = GO.generator_forward(sd, guide, crop, ngf=ngf, upd=upd)
    assert _rel(out.detach(), ref.detach()) < 1e-4
    (ref * gout).sum().backward()
    worst = {}
    for name, p in G.named_parameters():
        want = sd[name].grad
        assert want is not None, name
        assert p.grad is not None, name
        scale = float(want.abs().max())
        worst[name] = (float((p.grad - want).abs().max()), scale)
    # conv biases that feed a batch-statistic BatchNorm have an analytically ZERO gradient (the shift cancels): both sides then hold
    # fp32 rounding noise, so the error is measured against max(|want|, noise floor of a reduction over B*H*W terms)
    top = max(sc for _, sc in worst.values())
    floor = 1e-4 * top
    bad = {k: (e, sc) for k, (e, sc) in worst.items() if e > 2e-3 * max(sc, floor)}
    assert not bad, (top, sorted(bad.items(), key=lambda kv: -kv[1][0])[:8])
    cancelled = [k for k, (_, sc) in worst.items() if sc < floor]
    assert cancelled and all(k.endswith(".bias") for k in cancelled), cancelled
    buffers = dict(G.named_buffers())
    for k, v in upd.items():                             # training side effects: u, v and running statistics
        assert _rel(buffers[k], v.detach()) < 1e-4, k


# ------------------------------------------------------------------------------------------------- discriminator + losses
def _d_opt(ndf):
    return argparse.Namespace(ndf=ndf, norm_D="spectralinstance", label_nc=3, output_nc=3, num_D=2, n_layers_D=4, netD_subarch="n_layer",
                              no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0])


def _model(ndf, train_d):
    from emlight_b200.genprojector import VGG19, MultiscaleDiscriminator
    opt = _d_opt(ndf)
    netD = MultiscaleDiscriminator(opt)
    sdd = GO.init_discriminator_state_dict(2, ndf)
    netD.load_state_dict(sdd)
    netD.train(train_d)
    vgg = VGG19()
    sdv = GO.init_vgg_state_dict(3)
    vgg.load_state_dict({k[4:]: v for k, v in sdv.items()})
    crit = argparse.Namespace(vgg=vgg, weights=[1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0])
    return argparse.Namespace(opt=opt, netD=netD, criterionVGG=crit), sdd, sdv


def _images(h=32, w=64, B=2, seed=5):
    gen = torch.Generator().manual_seed(seed)
    guide = torch.rand(B, 3, h, w, generator=gen) * 2
    fake = torch.rand(B, 3, h, w, generator=gen) * 50 * torch.rand(B, 1, h, w, generator=gen) ** 4
    real = torch.rand(B, 3, h, w, generator=gen) * 50 * torch.rand(B, 1, h, w, generator=gen) ** 4
    mask = (torch.rand(B, 1, h, w, generator=gen) > 0.3).float()
    return guide, fake, real, mask


def test_generator_loss_gradient_wrt_fake_image(sim):
    """d(GAN + GAN_Feat + VGG + COS)/d(fake) through the discriminator, VGG and the loss seeds vs autograd of the oracle composition."""
    gt = sim
    model, sdd, sdv = _model(8, train_d=False)            # eval: stored spectral-norm vectors, which is what the oracle's D uses
    guide, fake, real, mask = _images()
    fr = fake.clone().requires_grad_(True)
    ref = GO.generator_losses(sdd, sdv, guide, fr, real, mask)
    sum(v.sum() for v in ref.values()).backward()
    tape = gt.Tape()
    box = _leaf(tape, fake)
    losses = gt.generator_losses(tape, model, fake, guide, real, mask)
    for got, key in zip(losses, ("GAN", "GAN_Feat", "VGG", "COS")):
        assert abs(float(got.sum()) - float(ref[key])) <= 1e-4 * abs(float(ref[key])) + 1e-6, key
        tape.add(got, torch.ones_like(got))
    assert losses[1].shape == (1,)
    tape.backward()
    assert _rel(box["g"], fr.grad) < 2e-3


def test_discriminator_loss_parameter_gradients(sim):
    gt = sim
    model, sdd, _ = _model(8, train_d=True)               # train: one power iteration per spectral conv per forward
    guide, fake, real, _ = _images(seed=6)
    u_before = {k: v.clone() for k, v in model.netD.state_dict().items() if k.endswith("weight_u")}
    params = list(model.netD.parameters())
    outs = gt.run_with_tape(lambda tape: tuple(gt.discriminator_losses(tape, model, fake, guide, real)), params)
    (outs[0] + outs[1]).backward()
    after = model.netD.state_dict()
    assert all(not torch.equal(after[k], v) for k, v in u_before.items())
    # the oracle's discriminator uses the stored vectors: give it the ones the forward above left behind
    sd = {k: (v.clone().requires_grad_(True) if k.endswith(("weight", "weight_orig", "bias")) else after[k].clone()) for k, v in sdd.items()}
    ref = GO.discriminator_losses(sd, guide, fake, real)
    assert abs(float(outs[0]) - float(ref["D_Fake"])) <= 1e-4 * abs(float(ref["D_Fake"])) + 1e-6
    assert abs(float(outs[1]) - float(ref["D_real"])) <= 1e-4 * abs(float(ref["D_real"])) + 1e-6
    (ref["D_Fake"] + ref["D_real"]).backward()
    top = max(float(sd[name].grad.abs().max()) for name, _ in model.netD.named_parameters())
    for name, p in model.netD.named_parameters():
        want = sd[name].grad
        assert p.grad is not None and want is not None, name
        assert float((p.grad - want).abs().max()) <= 2e-3 * max(float(want.abs().max()), 1e-4 * top), name
    with pytest.raises(RuntimeError, match="already back-propagated"):
        (outs[0] + outs[1]).backward()


def test_product_modules_do_not_use_the_stand_ins():
    """Outside this file's fixture gp_ops points at the C-ABI wrappers, and CPU tensors are refused before any launch."""
    from emlight_b200 import gp_ops
    from emlight_b200.genprojector import SPADEGenerator, _PackedConv
    assert gp_ops.PackedConv is _PackedConv and gp_ops.mm_nt.__module__ == "emlight_b200.gp_ops"
    G = SPADEGenerator(_g_opt(2)).train()
    G.autograd = True
    with pytest.raises(Exception):
        G(torch.rand(1, 3, 128, 256), torch.rand(1, 3, 64, 64))
    assert np.isfinite(float(torch.zeros(1)))


def test_pix2pix_model_training_iterations_on_stand_ins(sim, monkeypatch):
    """The trainer's two steps through `Pix2PixModel` itself (pix2pix_model.py:40-141 + trainers): loss dictionaries with an autograd
    node, backward, Adam, twice.  `.cuda()` is patched to identity so the dispatch code of genprojector.py runs on CPU tensors over
    the stand-ins; first-iteration losses and the generator's gradients are compared with autograd of the oracle composition."""
    import emlight_b200._lib as L
    from emlight_b200.genprojector import Pix2PixModel
    monkeypatch.setattr(torch.nn.Module, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(L, "require_cuda", lambda *a: None)
    ngf = ndf = 2
    opt = _d_opt(ndf)
    opt.__dict__.update(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3, num_upsampling_layers="normal",
                        crop_size=256, aspect_ratio=2.0, isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
    model = Pix2PixModel(opt)
    sdg, sdd, sdv = GO.init_generator_state_dict(1, ngf), GO.init_discriminator_state_dict(2, ndf), GO.init_vgg_state_dict(3)
    model.netG.load_state_dict(sdg)
    model.netD.load_state_dict(sdd)
    model.criterionVGG.vgg.load_state_dict({k[4:]: v for k, v in sdv.items()})
    model.train()
    model.autograd = True
    og, od = model.create_optimizers(opt)
    gen = torch.Generator().manual_seed(11)
    data = {"input": torch.rand(1, 3, 128, 256, generator=gen) * 2, "crop": torch.rand(1, 3, 64, 64, generator=gen),
            "warped": torch.rand(1, 3, 128, 256, generator=gen) * 20, "map": (torch.rand(1, 1, 128, 256, generator=gen) > 0.4).float()}
    # ---- iteration 1, generator step, against the oracle
    og.zero_grad()
    g_losses, generated = model(data, "generator")
    assert set(g_losses) == {"GAN", "GAN_Feat", "VGG", "COS"} and generated.shape == (1, 3, 128, 256)
    assert all(v.requires_grad for v in g_losses.values())
    sum(v.sum() for v in g_losses.values()).mean().backward()
    sdg_r = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))
                 else v.clone()) for k, v in sdg.items()}
    fake_ref = GO.generator_forward(sdg_r, data["input"], data["crop"], ngf=ngf, upd={})
    sdd_now = {k: v.detach().clone() for k, v in model.netD.state_dict().items()}       # D after ITS power iteration (stored vectors)
    want = GO.generator_losses(sdd_now, sdv, data["input"], fake_ref, data["warped"], data["map"])
    for k, ref in want.items():
        assert abs(float(g_losses[k].sum()) - float(ref)) <= 1e-3 * abs(float(ref)) + 1e-5, (k, float(g_losses[k].sum()), float(ref))
    sum(v.sum() for v in want.values()).backward()
    top = max(float(sdg_r[n].grad.abs().max()) for n, _ in model.netG.named_parameters())
    for name, p in model.netG.named_parameters():
        ref = sdg_r[name].grad
        assert p.grad is not None, name
        assert float((p.grad - ref).abs().max()) <= 5e-3 * max(float(ref.abs().max()), 1e-4 * top), name
    assert all(p.grad is None for p in model.netD.parameters())                          # the G step leaves D's gradients alone
    before = [p.detach().clone() for p in model.netG.parameters()]
    og.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, model.netG.parameters()))
    # ---- iteration 1, discriminator step (its no-grad generator forward is the forward-only product path: real kernels only, so
    #      here it is replaced by the same network run over the stand-ins)
    gt = sim
    monkeypatch.setattr(model, "generate_fake", lambda inp, crop: gt.generator(gt.Tape(), model.netG, inp, crop, True))
    od.zero_grad()
    d_losses = model(data, "discriminator")
    assert set(d_losses) == {"D_Fake", "D_real"}
    sum(v.sum() for v in d_losses.values()).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.netD.parameters())
    od.step()
    # ---- iteration 2 runs on fresh tapes
    og.zero_grad()
    g2, _ = model(data, "generator")
    sum(v.sum() for v in g2.values()).mean().backward()
    assert all(torch.isfinite(p.grad).all() for p in model.netG.parameters())
    # ---- without the opt-in the model returns plain values (no graph), as before
    monkeypatch.undo()
    model.autograd = False
    with pytest.raises(Exception):
        model(data, "generator")                       # forward-only product path: CPU tensors are refused before any launch


@pytest.mark.parametrize("stride,bias", [(1, True), (2, False)])
def test_sphere_conv_module_autograd(sim, monkeypatch, stride, bias):
    """`SphereConv2D(...).autograd = True`: weight, bias and input gradients of the stand-alone layer vs autograd of the oracle."""
    import emlight_b200._lib as L
    from emlight_b200.genprojector import SphereConv2D
    monkeypatch.setattr(L, "require_cuda", lambda *a: None)
    gen = torch.Generator().manual_seed(5 + stride)
    sc = SphereConv2D(6, 5, stride=stride, bias=bias)
    with torch.no_grad():
        sc.weight.copy_(torch.randn(5, 6, 3, 3, generator=gen))
        if bias:
            sc.bias.copy_(torch.randn(5, generator=gen))
    sc.autograd = True
    x = torch.randn(2, 6, 8, 16, generator=gen, requires_grad=True)
    y = sc(x)
    gy = torch.randn(y.shape, generator=gen)
    (y * gy).sum().backward()
    xr, wr = x.detach().clone().requires_grad_(True), sc.weight.detach().clone().requires_grad_(True)
    br = sc.bias.detach().clone().requires_grad_(True) if bias else None
    ref = GO.sphere_conv(xr, wr, br, stride)
    assert _rel(y.detach(), ref.detach()) < 1e-5
    (ref * gy).sum().backward()
    assert _rel(x.grad, xr.grad) < 1e-4 and _rel(sc.weight.grad, wr.grad) < 1e-4
    if bias:
        assert _rel(sc.bias.grad, br.grad) < 1e-4
    # without input gradients only the parameters receive them
    sc.zero_grad()
    (sc(x.detach()) * gy).sum().backward()
    assert _rel(sc.weight.grad, wr.grad) < 1e-4
