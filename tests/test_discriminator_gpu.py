"""GPU: discriminator, GAN / feature-matching / VGG / cosine losses and the Pix2PixModel dispatch through the drop-in modules, against
the reference-generated golden (tests/golden/discriminator.npz) and the CPU oracle.  Tolerance: 1e-3 relative (BASELINE north_star)."""
import argparse
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import genprojector_oracle as GO
from test_discriminator_cpu import d_opt, golden_inputs

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("hi,wi,c", [(128, 256, 8), (17, 31, 5), (2, 2, 3)])
def test_avg_pool_no_pad_count(cuda, lib, hi, wi, c):
    from emlight_b200 import _lib
    x = torch.randn(2, hi, wi, c, generator=torch.Generator().manual_seed(hi)).to(cuda)
    ho, wo = (hi + 1) // 2, (wi + 1) // 2
    out = torch.empty(2, ho, wo, c, device=cuda)
    _lib.check(lib.eml_pool2d(_lib.ptr(x), c, hi, wi, _lib.ptr(out), c, c, 2, 0, None), "pool")
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 3, 2, [1, 1], count_include_pad=False).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and (out - ref).abs().max() <= 1e-6
    if hi % 2 == 0 and wi % 2 == 0:
        mp = torch.empty(2, hi // 2, wi // 2, c, device=cuda)
        _lib.check(lib.eml_pool2d(_lib.ptr(x), c, hi, wi, _lib.ptr(mp), c, c, 2, 1, None), "pool")
        assert torch.equal(mp, F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1))
    else:
        assert lib.eml_pool2d(_lib.ptr(x), c, hi, wi, _lib.ptr(out), c, c, 2, 1, None) != 0       # max-pool needs even sizes


def test_loss_reductions_match_torch(cuda, lib):
    from emlight_b200 import _lib
    g = torch.Generator().manual_seed(3)
    M, C, P = 1237, 13, 16
    a = torch.randn(M, P, generator=g).to(cuda)
    b = torch.randn(M, P, generator=g).to(cuda)
    m = torch.rand(M, generator=g).to(cuda)
    av, bv = a[:, :C].double(), b[:, :C].double()
    want = [av.sum(), torch.clamp(av - 1, max=0).sum(), torch.clamp(-av - 1, max=0).sum(), (av - bv).abs().sum(),
            ((av - bv).abs() * (m + (1 - m) * 50).double()[:, None]).sum(), (1 - F.cosine_similarity(av, bv, dim=1, eps=1e-20)).sum()]
    for mode, w in enumerate(want):
        acc = torch.zeros(1, dtype=torch.float64, device=cuda)
        _lib.check(lib.eml_loss_reduce(_lib.ptr(a), P, _lib.ptr(b), P, _lib.ptr(m), M, C, mode, _lib.ptr(acc), None), "reduce")
        _lib.check(lib.eml_loss_reduce(_lib.ptr(a), P, _lib.ptr(b), P, _lib.ptr(m), M, C, mode, _lib.ptr(acc), None), "reduce")   # accumulates
        assert abs(float(acc) - 2 * float(w)) <= 2e-6 * abs(float(w)) + 1e-9, mode
    assert lib.eml_loss_reduce(_lib.ptr(a), P, None, P, None, M, C, 3, _lib.ptr(torch.zeros(1, dtype=torch.float64, device=cuda)), None) != 0
    assert lib.eml_loss_reduce(_lib.ptr(a), P, None, P, None, M, C, 9, _lib.ptr(torch.zeros(1, dtype=torch.float64, device=cuda)), None) != 0


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp32", 1e-3), ("bf16", 5e-2)])
def test_discriminator_matches_reference_golden(cuda, precision, tol):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "discriminator.npz"))
    ndf = int(g["ndf"])
    D = E.MultiscaleDiscriminator(d_opt(ndf), precision=precision).to(cuda).eval()
    D.load_state_dict(GO.init_discriminator_state_dict(int(g["sd_seed"]), ndf))
    guide, fake, real, _ = golden_inputs(int(g["in_seed"]))
    x = torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0).to(cuda)
    out = D(x)
    assert len(out) == 2 and all(len(o) == 5 for o in out)
    for i, fl in enumerate(out):
        for j, f in enumerate(fl):
            want = g["d%d_%d" % (i, j)]
            got = f.cpu().numpy()[:, ::max(1, f.shape[1] // 8), ::2, ::2] if j < 4 else f.cpu().numpy()
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= tol * np.abs(want).max(), (i, j, np.abs(got - want).max() / np.abs(want).max())


def test_discriminator_no_feat_returns_final_only(cuda):
    import emlight_b200 as E
    D = E.MultiscaleDiscriminator(d_opt(8, no_ganFeat_loss=True)).to(cuda).eval()
    out = D(torch.rand(2, 6, 32, 64, device=cuda))
    assert [len(o) for o in out] == [1, 1] and out[0][0].shape == (2, 3, 4, 8) and out[1][0].shape == (2, 3, 2, 4)
    with pytest.raises(RuntimeError):
        D(torch.rand(2, 6, 32, 64))                                            # host tensors are rejected, there is no CPU path


def test_hinge_and_feature_losses_match_reference_golden(cuda):
    import emlight_b200 as E
    from emlight_b200.genprojector import cosine_loss, feature_matching_loss, _nchw_to_nhwc
    g = np.load(os.path.join(GOLDEN, "discriminator.npz"))
    ndf = int(g["ndf"])
    D = E.MultiscaleDiscriminator(d_opt(ndf)).to(cuda).eval()
    D.load_state_dict(GO.init_discriminator_state_dict(int(g["sd_seed"]), ndf))
    guide, fake, real, mask = [t.to(cuda) for t in golden_inputs(int(g["in_seed"]))]
    x = torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0)
    pred = D(x)
    pf = [[t[:1] for t in p] for p in pred]
    pr = [[t[1:] for t in p] for p in pred]
    crit = E.GANLoss("hinge")
    got = {"GAN": crit(pf, True, for_discriminator=False), "D_Fake": crit(pf, False), "D_real": crit(pr, True),
           "GAN_Feat": feature_matching_loss(D.features_nhwc(_nchw_to_nhwc(x, 8), 2, 128, 256), 1, mask), "COS": cosine_loss(fake, real) * 5}
    assert got["GAN_Feat"].shape == (1,) and got["GAN"].shape == ()
    for k, v in got.items():
        want = float(g["loss_" + k])
        assert abs(float(v) - want) <= 1e-3 * abs(want) + 1e-5, (k, float(v), want)
    with pytest.raises(NotImplementedError):
        E.GANLoss("ls")
    with pytest.raises(ValueError):
        E.GANLoss("nope")


@pytest.mark.parametrize("precision,tol", [("bf16x3", 1e-3), ("fp32", 1e-3)])
def test_vgg_features_and_loss_match_reference_golden(cuda, precision, tol):
    import emlight_b200 as E
    g = np.load(os.path.join(GOLDEN, "discriminator.npz"))
    crit = E.VGGLoss([0], precision=precision)
    crit.vgg.load_state_dict(GO.init_vgg_state_dict(int(g["vgg_seed"]), p=""))
    _, fake, real, _ = [t.to(cuda) for t in golden_inputs(int(g["in_seed"]))]
    feats = crit.vgg(fake)
    assert [tuple(f.shape) for f in feats] == [(1, 64, 128, 256), (1, 128, 64, 128), (1, 256, 32, 64), (1, 512, 16, 32), (1, 512, 8, 16)]
    for j, f in enumerate(feats):
        want = g["vgg_%d" % j]
        got = f.cpu().numpy()[:, ::max(1, f.shape[1] // 8), ::4, ::4]
        assert np.abs(got - want).max() <= tol * np.abs(want).max(), (j, np.abs(got - want).max() / np.abs(want).max())
    loss = crit(fake, real) * 5
    assert abs(float(loss) - float(g["loss_VGG"])) <= tol * float(g["loss_VGG"])


def test_vgg_loss_loads_a_torchvision_checkpoint(cuda, tmp_path, monkeypatch):
    """architecture.py:95 builds torchvision.models.vgg19(pretrained=True): a checkpoint in torchvision's `features.N.*` layout must load
    into slice1..5 (path argument, $EML_VGG19_WEIGHTS), and its absence must warn -- or raise when pretrained weights are required."""
    import torchvision
    import emlight_b200 as E
    torch.manual_seed(3)
    tv = torchvision.models.vgg19(weights=None)
    path = str(tmp_path / "vgg19-test.pth")
    torch.save(tv.state_dict(), path)
    monkeypatch.delenv("EML_VGG19_WEIGHTS", raising=False)
    monkeypatch.setenv("TORCH_HOME", str(tmp_path / "no_hub"))
    with pytest.warns(RuntimeWarning, match="RANDOMLY INITIALISED"):
        crit0 = E.VGGLoss([0])
    assert not crit0.pretrained
    with pytest.raises(RuntimeError, match="no ImageNet VGG19 checkpoint"):
        E.VGGLoss([0], require_pretrained=True)
    monkeypatch.setenv("EML_VGG19_WEIGHTS", path)
    crit = E.VGGLoss([0])
    assert crit.pretrained
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = tv.features[:2](x)                                                     # conv1_1 + relu = slice1
    got = crit.vgg(x.to(cuda))[0].cpu()
    assert float((got - want).abs().max()) <= 1e-3 * float(want.abs().max())
    with pytest.raises(RuntimeError, match="does not cover"):
        crit.load_torchvision_state_dict({"features.0.weight": tv.state_dict()["features.0.weight"]})


def test_pix2pix_model_modes_match_oracle(cuda):
    """Pix2PixModel.forward(data, mode) for the three modes (pix2pix_model.py:40-54) vs the oracle composition on small networks."""
    import emlight_b200 as E
    ngf = ndf = 8
    opt = d_opt(ndf, ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3, num_upsampling_layers="normal",
                crop_size=256, aspect_ratio=2.0, isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
    model = E.Pix2PixModel(opt)
    sdg, sdd, sdv = GO.init_generator_state_dict(1, ngf), GO.init_discriminator_state_dict(2, ndf), GO.init_vgg_state_dict(3)
    model.netG.load_state_dict(sdg)
    model.netD.load_state_dict(sdd)
    model.criterionVGG.vgg.load_state_dict({k[4:]: v for k, v in sdv.items()})
    gen = torch.Generator().manual_seed(11)
    data = {"input": torch.rand(2, 3, 128, 256, generator=gen) * 2, "crop": torch.rand(2, 3, 96, 128, generator=gen),
            "warped": torch.rand(2, 3, 128, 256, generator=gen) * 20, "map": (torch.rand(2, 1, 128, 256, generator=gen) > 0.4).float()}
    with torch.no_grad():
        fake_ref = GO.generator_forward(sdg, data["input"], data["crop"], ngf=ngf)
        gl_ref = GO.generator_losses(sdd, sdv, data["input"], fake_ref, data["warped"], data["map"])
        dl_ref = GO.discriminator_losses(sdd, data["input"], fake_ref, data["warped"])
    fake = model(data, "inference")
    assert _rel(fake.cpu(), fake_ref) <= 1e-3
    gl, generated = model(data, "generator")
    dl = model(data, "discriminator")
    assert torch.equal(generated, fake)
    assert set(gl) == {"GAN", "GAN_Feat", "VGG", "COS"} and set(dl) == {"D_Fake", "D_real"}
    for k, ref in list(gl_ref.items()) + list(dl_ref.items()):
        got = float((gl if k in gl else dl)[k])
        assert abs(got - float(ref)) <= 2e-3 * abs(float(ref)) + 1e-4, (k, got, float(ref))
    with pytest.raises(ValueError):
        model(data, "train")
    og, od = model.create_optimizers(opt)
    assert og.defaults["lr"] == opt.lr / 2 and od.defaults["lr"] == opt.lr * 2 and og.defaults["betas"] == (0.0, 0.9)
