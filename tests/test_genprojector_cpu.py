"""CPU: GenProjector oracle vs the reference-generated golden; host-side sampling tables vs grid_sample; state_dict contract."""
import argparse
import os

import numpy as np
import torch
import torch.nn.functional as F

from conftest import GOLDEN
from oracle import genprojector_oracle as GO


def _opt(ngf):
    return argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                              num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)


def test_generator_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "generator.npz"))
    ngf = int(g["ngf"])
    sd = GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=ngf)
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(1, 3, 160, 160, generator=gen)
    with torch.no_grad():
        out = GO.generator_forward(sd, guide, crop, ngf)
    assert np.abs(out.numpy()[:, :, ::2, ::2] - g["out"]).max() <= 1e-3         # same op sequence: thread-order noise only
    assert abs(float(out.double().mean()) - float(g["out_mean"])) < 1e-4
    # the standalone SphereConv2D fixture (stride 2)
    y = GO.sphere_conv(torch.from_numpy(g["sc_x"]), torch.from_numpy(g["sc_weight"]), torch.from_numpy(g["sc_bias"]), stride=2)
    assert np.abs(y.numpy() - g["sc_y"]).max() <= 1e-5


def test_generator_train_mode_oracle_matches_reference_golden():
    """Train-mode forward (batch-statistic BatchNorm in SPADE, spectral-norm power iteration): two consecutive steps of the reference
    module in .train() (oracle/make_golden_gen_train.py) vs the oracle threading the updated buffers through."""
    g = np.load(os.path.join(GOLDEN, "generator_train.npz"))
    ngf = int(g["ngf"])
    sd = GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=ngf)
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(2, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(2, 3, 128, 128, generator=gen)
    with torch.no_grad():
        upd = {}
        out1 = GO.generator_forward(sd, guide, crop, ngf, upd=upd)
        sd.update(upd)
        upd = {}
        out2 = GO.generator_forward(sd, guide * 0.5, crop, ngf, upd=upd)
        sd.update(upd)
    assert np.abs(out1.numpy()[:, :, ::2, ::2] - g["out1"]).max() <= 1e-3
    assert np.abs(out2.numpy()[:, :, ::2, ::2] - g["out2"]).max() <= 1e-3
    for k in g.files:
        if k.startswith("buf_") and not k.endswith("num_batches_tracked"):
            want = g[k]
            assert np.abs(sd[k[4:]].numpy() - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), k


def test_sampling_tables_reproduce_grid_sample():
    """The product's 4-tap tables (emlight_b200.genprojector._sphere_lut) applied on the host == grid_sample on the reference grid."""
    from emlight_b200.genprojector import _conv_lut, _sphere_lut
    for (h, w, s) in ((4, 8, 1), (16, 32, 1), (16, 32, 2), (128, 256, 1)):
        idx, wgt, ho, wo = _sphere_lut(h, w, s)
        x = torch.randn(1, 2, h, w, generator=torch.Generator().manual_seed(h))
        ref = F.grid_sample(x, GO.sphere_grid(h, w, s), mode="bilinear", padding_mode="zeros", align_corners=False)   # (1,2,3ho,3wo)
        ref = ref.view(1, 2, ho, 3, wo, 3).permute(0, 1, 2, 4, 3, 5).reshape(2, ho * wo, 9)
        flat = torch.cat([x.view(2, h * w), torch.zeros(2, 1)], 1)                                                     # slot -1 -> 0
        got = (flat[:, torch.from_numpy(idx).long()] * torch.from_numpy(wgt)).sum(-1)
        assert (got - ref).abs().max() < 2e-5, (h, w, s)
    idx, wgt, ho, wo = _conv_lut(9, 12, 2)
    x = torch.randn(1, 1, 9, 12)
    cols = F.unfold(x, 3, padding=1, stride=2).view(9, ho * wo).t()                                                    # (pixels, taps)
    flat = torch.cat([x.view(-1), torch.zeros(1)])
    got = (flat[torch.from_numpy(idx).long()] * torch.from_numpy(wgt)).sum(-1)
    assert torch.equal(got, cols)


def test_generator_state_dict_contract():
    import emlight_b200 as E
    ngf = 8
    sd = GO.init_generator_state_dict(0, ngf)
    G = E.SPADEGenerator(_opt(ngf))
    mine = G.state_dict()
    assert set(mine) == set(sd)
    assert all(tuple(mine[k].shape) == tuple(sd[k].shape) for k in sd)
    G.load_state_dict(sd)
    full = E.SPADEGenerator(_opt(64))
    assert sum(v.numel() for v in full.state_dict().values()) == 118430576            # reference: 253 tensors, 118.43 M values
    assert len(full.state_dict()) == 253


def test_generator_oracle_matches_reference_golden_at_baseline_width():
    """The restatement at ngf = 64 against the reference's own SPADEGenerator (oracle/make_golden_ngf64.py)."""
    g = np.load(os.path.join(GOLDEN, "genprojector_w64.npz"))
    w = int(g["width"])
    sd = GO.init_generator_state_dict(seed=int(g["sd_seed"]), ngf=w)
    gen = torch.Generator().manual_seed(int(g["in_seed"]))
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(1, 3, 128, 128, generator=gen)
    with torch.no_grad():
        out = GO.generator_forward(sd, guide, crop, ngf=w)
    assert np.abs(out.numpy()[:, :, ::4, ::4] - g["out"]).max() / 50.0 <= 1e-5
