"""GPU: needlet basis kernel, projection / reconstruction GEMMs and sparsification through the C ABI vs the CPU oracle
(oracle/needlets_oracle.py; parity unpinned -- see its header) and, at BASELINE config-5 size, vs float64 torch matmul."""
import numpy as np
import pytest
import torch

from oracle import needlets_oracle as NO

pytestmark = pytest.mark.gpu


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
