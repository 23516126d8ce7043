"""GPU: needlet basis kernel, projection / reconstruction GEMMs and sparsification through the C ABI vs tests/golden/needlets.npz
(outputs of the reference's own Needlets/ files, oracle/make_golden_needlets.py), vs the CPU oracle (oracle/needlets_oracle.py) and,
at BASELINE config-5 size, vs float64 torch matmul."""
import os

import numpy as np
import pytest
import torch

from oracle import needlets_oracle as NO

pytestmark = pytest.mark.gpu


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "needlets.npz")


def test_matches_reference_golden(cuda):
    """The reference's SNvertex outputs (jmax = 3 on grid points incl. poles / seam; jmax = 2 on a whole 16x32 grid) and its projection /
    sparsification / reconstruction script lines, against the basis kernel and the tcgen05 GEMMs."""
    from emlight_b200 import needlets as PN
    g = np.load(GOLD)
    SN1, SN2, SN = PN.SNvertex(g["theta3"], g["phi3"], 3, device=cuda)
    assert np.abs(SN.cpu().numpy() - g["SN_3"]).max() < 1e-12
    assert np.abs(SN1.cpu().numpy() - g["SN1_3"]).max() < 1e-12 and np.abs(SN2.cpu().numpy() - g["SN2_3"]).max() < 1e-12
    tr = PN.NeedletTransform(jmax=2, h=16, w=32, device=cuda)
    assert np.abs(tr.SN.cpu().numpy() - g["SN_2"]).max() < 1e-12
    assert np.array_equal(tr.omega.cpu().numpy(), g["omega2"].reshape(-1))
    pano = torch.from_numpy(g["pano2"]).float()[None]                                     # (1, P, 3) like im.reshape((-1, 3))
    coef = tr.project(pano.to(cuda))
    ref = (g["SN_2"] * g["omega2"].reshape(-1, 1)).T @ pano[0].double().numpy()            # same fp32-rounded input as the kernel saw
    assert np.abs(ref - g["coef2"]).max() < 1e-6 * np.abs(g["coef2"]).max()
    assert np.abs(coef[0].cpu().numpy() - g["coef2"]).max() < 1e-3 * np.abs(g["coef2"]).max()   # the north-star bar; bf16x3 measures ~1e-5
    rec = tr.reconstruct(torch.from_numpy(g["coef2"]).float()[None].to(cuda))
    assert np.abs(rec[0].cpu().numpy() - g["rec2"]).max() < 1e-3 * np.abs(g["rec2"]).max()
    tr3 = PN.NeedletTransform.__new__(PN.NeedletTransform)                               # sparsify needs only the level table of jmax = 3
    tr3.device, tr3.n, tr3.jmax, tr3.level_slices = torch.device(cuda), 1021, 3, [(1, 13), (13, 61), (61, 253), (253, 1021)]
    sp = tr3.sparsify(torch.from_numpy(g["sp_in"]).float()[None].to(cuda))[0].cpu().numpy()
    want = g["sp_out"]
    edge = np.zeros(want.shape, dtype=bool)                                             # fp32 input: entries within rounding of a block's threshold
    for lo, hi in ((253, 1021), (61, 253)):
        blk = np.abs(g["sp_in"][lo:hi])
        edge[lo:hi] = np.abs(blk - 0.1 * blk.max()) < 1e-6 * blk.max()
    assert np.array_equal((sp != 0) | edge, (want != 0) | edge)
    assert np.abs(sp - want)[~edge].max() <= 1e-6 * np.abs(want).max()


def test_basis_matches_oracle(cuda):
    from emlight_b200 import needlets as PN
    rng = np.random.default_rng(3)
    theta = np.concatenate(([0.0, np.pi, np.pi / 2], rng.uniform(0, np.pi, 400)))
    phi = np.concatenate(([0.0, 2 * np.pi, 1.0], rng.uniform(0, 2 * np.pi, 400)))
    ref = NO.needlet_matrix(theta, phi, 3)
    SN1, SN2, SN = PN.SNvertex(theta, phi, 3, device=cuda)
    assert SN.dtype == torch.float64 and tuple(SN.shape) == (403, 1021)
    assert np.abs(SN.cpu().numpy() - ref).max() < 1e-12
    pair, use = NO.spneedlet_pair(3)
    assert tuple(SN1.shape) == tuple(SN2.shape) == (403, 511)
    assert np.abs(SN1.cpu().numpy() - np.hstack((ref[:, :1], ref[:, 1:][:, use]))).max() < 1e-12
    assert np.abs(SN2.cpu().numpy() - np.hstack((ref[:, :1], ref[:, 1:][:, pair][:, use]))).max() < 1e-12


@pytest.fixture(scope="module")
def transform(cuda):
    from emlight_b200 import needlets as PN
    return PN.NeedletTransform(jmax=3, device=cuda)


@pytest.fixture(scope="module")
def oracle_sn():
    theta, phi = NO.pano_grid()
    return NO.needlet_matrix(theta, phi, 3), NO.solid_angle_map().reshape(-1)


def test_project_sparsify_reconstruct_match_oracle(cuda, transform, oracle_sn):
    SN, omega = oracle_sn
    assert np.abs(transform.SN.cpu().numpy() - SN).max() < 1e-12
    g = torch.Generator().manual_seed(11)
    pano = torch.exp(1.5 * torch.randn(2, 128 * 256, 3, generator=g))                  # HDR-like, (B, P, 3) as the reference reshapes it
    coef = transform.project(pano.to(cuda))
    assert tuple(coef.shape) == (2, 1021, 3)
    for b in range(2):
        ref = NO.project(pano[b].double().numpy(), SN, omega)
        err = np.abs(coef[b].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-3, err                                                          # parity bar; bf16x3 measures ~1e-5
    # NCHW input (our render's layout) gives the same coefficients
    coef2 = transform.project(pano.view(2, 128, 256, 3).permute(0, 3, 1, 2).contiguous().to(cuda))
    assert torch.allclose(coef, coef2, rtol=0, atol=1e-6 * float(coef.abs().max()))
    sp = transform.sparsify(coef)
    for b in range(2):
        ref = NO.sparsify(coef[b].cpu().numpy().astype(np.float64))
        assert np.array_equal(sp[b].cpu().numpy() != 0, ref != 0)
        assert np.abs(sp[b].cpu().numpy() - ref).max() <= 1e-6 * np.abs(ref).max()
    rec = transform.reconstruct(sp)
    assert tuple(rec.shape) == (2, 128 * 256, 3)
    for b in range(2):
        ref = NO.reconstruct(SN, sp[b].cpu().numpy().astype(np.float64))
        err = np.abs(rec[b].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err < 1e-3, err


def test_transform_at_config5_size(cuda, transform):
    """BASELINE configs[4]: batch 512 -- linearity of the projection and agreement with a float64 matmul on a slice."""
    B = 512
    g = torch.Generator(device=cuda).manual_seed(5)
    x = torch.exp(torch.randn(B, 3, 128, 256, generator=g, device=cuda))
    y = torch.rand(B, 3, 128, 256, generator=g, device=cuda)
    cx, cy, cxy = transform.project(x), transform.project(y), transform.project(2.0 * x + y)
    scale = float(cxy.abs().max())
    assert float((cxy - (2.0 * cx + cy)).abs().max()) < 1e-4 * scale
    W = (transform.SN * transform.omega[:, None])                                      # (P, n) float64
    ref = torch.einsum("bcp,pn->bnc", x[:4].reshape(4, 3, -1).double(), W)
    assert float((cx[:4].double() - ref).abs().max()) < 1e-3 * float(ref.abs().max())
    rec = transform.reconstruct(cx)
    assert tuple(rec.shape) == (B, 128 * 256, 3)
    ref_rec = torch.einsum("pn,bnc->bpc", transform.SN, cx[:2].double())
    assert float((rec[:2].double() - ref_rec).abs().max()) < 1e-3 * float(ref_rec.abs().max())


def test_errors(cuda, transform):
    with pytest.raises(ValueError):
        transform.project(torch.zeros(1, 100, 3, device=cuda))
    with pytest.raises(RuntimeError):
        transform.project(torch.zeros(1, 128 * 256, 3))
    with pytest.raises(ValueError):
        transform.reconstruct(torch.zeros(1, 10, 3, device=cuda))
