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
    with pytest.raises(NotImplementedError):
        G.train()(guide.to(cuda), crop.to(cuda))
