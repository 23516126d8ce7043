"""CPU checks for the discriminator / loss rows: the oracle against the committed reference outputs
(tests/golden/discriminator.npz, made by oracle/make_golden.py running the reference's own classes and loss-composition methods),
and the drop-in modules' state_dict contract."""
import argparse
import os

import numpy as np
import torch

from oracle import genprojector_oracle as GO

GOLD = os.path.join(os.path.dirname(__file__), "golden", "discriminator.npz")


def d_opt(ndf=16, **kw):
    o = argparse.Namespace(ndf=ndf, norm_D="spectralinstance", label_nc=3, output_nc=3, num_D=2, n_layers_D=4, netD_subarch="n_layer",
                           no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0])
    o.__dict__.update(kw)
    return o


def golden_inputs(seed=5):
    gen = torch.Generator().manual_seed(seed)
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    fake = torch.rand(1, 3, 128, 256, generator=gen) * 50 * torch.rand(1, 1, 128, 256, generator=gen) ** 4
    real = torch.rand(1, 3, 128, 256, generator=gen) * 50 * torch.rand(1, 1, 128, 256, generator=gen) ** 4
    mask = (torch.rand(1, 1, 128, 256, generator=gen) > 0.3).float()
    return guide, fake, real, mask


def test_oracle_matches_reference_losses_and_features():
    g = np.load(GOLD)
    guide, fake, real, mask = golden_inputs(int(g["in_seed"]))
    sd = GO.init_discriminator_state_dict(int(g["sd_seed"]), int(g["ndf"]))
    sdv = GO.init_vgg_state_dict(int(g["vgg_seed"]))
    with torch.no_grad():
        feats = GO.multiscale_discriminator(sd, torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0))
        gl = GO.generator_losses(sd, sdv, guide, fake, real, mask)
        dl = GO.discriminator_losses(sd, guide, fake, real)
        vf = GO.vgg_features(sdv, fake)
    for i, fl in enumerate(feats):
        assert len(fl) == 5
        for j, f in enumerate(fl):
            want = g["d%d_%d" % (i, j)]
            got = f.numpy()[:, ::max(1, f.shape[1] // 8), ::2, ::2] if j < 4 else f.numpy()
            assert got.shape == want.shape and np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), (i, j)
    for j, f in enumerate(vf):
        want = g["vgg_%d" % j]
        got = f.numpy()[:, ::max(1, f.shape[1] // 8), ::4, ::4]
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max(), j
    for k, v in list(gl.items()) + list(dl.items()):
        assert abs(float(v) - float(g["loss_" + k])) <= 1e-6 * abs(float(g["loss_" + k])), k


def test_mask_is_resized_progressively():
    """The feature-matching mask is nearest-resized from its previous size (pix2pix_model.py:111): D1's first layer sees the 16x32 mask
    upsampled to 32x64, not the 128x256 mask downsampled -- the two differ, and the oracle follows the reference."""
    m = (torch.rand(1, 1, 128, 256, generator=torch.Generator().manual_seed(0)) > 0.5).float()
    import torch.nn.functional as F
    chained = F.interpolate(F.interpolate(F.interpolate(F.interpolate(m, size=(64, 128)), size=(32, 64)), size=(16, 32)), size=(32, 64))
    direct = F.interpolate(m, size=(32, 64))
    assert not torch.equal(chained, direct)


def test_discriminator_state_dict_contract():
    import emlight_b200 as E
    sd = GO.init_discriminator_state_dict(0, 16)
    D = E.MultiscaleDiscriminator(d_opt(16))
    mine = D.state_dict()
    assert list(mine) == list(sd)
    assert all(tuple(mine[k].shape) == tuple(sd[k].shape) for k in sd)
    D.load_state_dict(sd)
    full = E.MultiscaleDiscriminator(d_opt(64)).state_dict()
    # reference (instantiated in the build container): per scale model0.0.{weight,bias}, model{1,2,3}.0.0.weight_{orig,u,v}, model4.0.{weight,bias}
    assert len(full) == 26 and tuple(full["discriminator_1.model3.0.0.weight_orig"].shape) == (512, 256, 3, 3)
    assert tuple(full["discriminator_0.model4.0.weight"].shape) == (3, 512, 3, 3)


def test_vgg_state_dict_contract():
    import emlight_b200 as E
    sdv = GO.init_vgg_state_dict(0, p="")
    V = E.VGG19()
    assert set(V.state_dict()) == set(sdv)
    assert all(tuple(V.state_dict()[k].shape) == tuple(sdv[k].shape) for k in sdv)
    assert not any(p.requires_grad for p in V.parameters())
    import torchvision
    tv = torchvision.models.vgg19(weights=None).features.state_dict()
    assert sorted(k.split(".", 1)[1] for k in sdv) == sorted(k for k in tv if int(k.split(".")[0]) < 30)


def test_product_never_imports_the_oracle():
    import emlight_b200.genprojector as gp
    src = open(gp.__file__).read()
    assert "oracle" not in src.replace("oracle/", "")
