"""CPU: the needlets oracle is pinned against tests/golden/needlets.npz -- outputs of the reference's own Needlets/ files run by
oracle/make_golden_needlets.py (healpy alone stays a restatement, checked against its documentation's examples) -- and checked against
itself (line-by-line transcription of sphere_needlets.py vs the addition-theorem closed form) and against analytic properties; the
product's host-side tables (emlight_b200.needlets) are checked against the oracle's independent implementations."""
import os
import numpy as np
import pytest

from oracle import needlets_oracle as NO


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "needlets.npz")


def test_oracle_matches_reference_golden():
    """Everything below was RETURNED by the reference's code (sphere_needlets.SNvertex / fun_b / spneedlet_pair, utils.getSolidAngleMap,
    the exec'd script lines gt_gen_j3.py:39-43, mat_gen2.py:43-51,55); the oracle must reproduce it to rounding."""
    g = np.load(GOLD)
    SN3 = NO.needlet_matrix(g["theta3"], g["phi3"], 3)                      # jmax = 3: poles, the phi = 0 / 2 pi seam, random grid points
    assert SN3.shape == g["SN_3"].shape == (24, 1021)
    assert np.abs(SN3 - g["SN_3"]).max() < 1e-12
    pair, use = NO.spneedlet_pair(3)
    assert np.array_equal(pair, g["pair3"]) and np.array_equal(use, g["use3"])
    assert np.abs(np.hstack((SN3[:, :1], SN3[:, 1:][:, use])) - g["SN1_3"]).max() < 1e-12
    assert np.abs(np.hstack((SN3[:, :1], SN3[:, 1:][:, pair][:, use])) - g["SN2_3"]).max() < 1e-12
    SN2 = NO.needlet_matrix(g["theta2"], g["phi2"], 2)                      # the whole 16x32 grid at jmax = 2
    assert np.abs(SN2 - g["SN_2"]).max() < 1e-12
    th, ph = NO.pano_grid(16, 32)
    assert np.array_equal(th, g["theta2"]) and np.array_equal(ph, g["phi2"])
    omega = NO.solid_angle_map(32)
    assert np.array_equal(omega, g["omega2"])
    coef = NO.project(g["pano2"], g["SN_2"], omega.reshape(-1))
    assert np.abs(coef - g["coef2"]).max() < 1e-12 * np.abs(g["coef2"]).max()
    assert np.abs(NO.reconstruct(g["SN_2"], g["coef2"]) - g["rec2"]).max() < 1e-12 * np.abs(g["rec2"]).max()
    assert np.array_equal(NO.sparsify(g["sp_in"]), g["sp_out"])             # mask and values, bit for bit
    assert np.array_equal(np.array([NO.fun_b(x) for x in g["b_x"]]), g["b_val"])


def test_product_window_matches_reference_golden():
    from emlight_b200 import needlets as PN
    g = np.load(GOLD)
    assert np.abs(np.array([PN.fun_b(x) for x in g["b_x"]]) - g["b_val"]).max() < 1e-15
    pair, use = PN.spneedlet_pair(3)
    assert np.array_equal(pair, g["pair3"]) and np.array_equal(use, g["use3"])


def test_transcription_equals_closed_form():
    rng = np.random.default_rng(0)
    theta = np.concatenate(([0.0, np.pi, np.pi / 2], rng.uniform(0, np.pi, 4)))
    phi = np.concatenate(([0.0, 2 * np.pi, 0.3], rng.uniform(0, 2 * np.pi, 4)))
    direct = NO.SNvertex_direct(theta, phi, 2)                 # sphere_needlets.py:34-104,182-238 step by step
    closed = NO.needlet_matrix(theta, phi, 2)
    assert direct.shape == closed.shape == (7, 1 + 12 + 48 + 192)
    assert np.abs(direct - closed).max() < 1e-13


@pytest.mark.parametrize("nside", [1, 2, 4, 8])
def test_healpix_ring_scheme(nside):
    v = NO.pix2vec(nside)
    th, ph = NO.pix2ang(nside)
    assert v.shape == (3, 12 * nside * nside)
    assert np.abs((v ** 2).sum(0) - 1).max() < 1e-14
    assert np.abs(v.sum(1)).max() < 1e-12                       # centres of an equal-area, symmetric pixelisation
    assert np.all(np.diff(th) >= -1e-15)                        # RING order: colatitude never decreases
    # every pixel has its antipode in the set (what spneedlet_pair relies on)
    corr = v.T @ v
    assert np.all((corr + 1 < 1e-10).sum(1) == 1)
    if nside == 1:                                              # published values: rings at z = 2/3, 0, -2/3; first pixel at phi = pi/4
        assert np.allclose(np.cos(th), np.repeat([2 / 3, 0, -2 / 3], 4))
        assert np.allclose(ph[:4], np.pi / 4 + np.arange(4) * np.pi / 2)
        assert np.allclose(ph[4:8], np.arange(4) * np.pi / 2)


def test_window_is_a_partition_of_unity():
    # sum_j b(l / B^j)^2 = 1 for 1 <= l <= B^jmax (needlet frame condition); b vanishes outside (1/B, B)
    bv = NO.b_vector(4, 16)
    assert np.allclose((bv ** 2).sum(0)[:16], 1.0, atol=1e-8)
    assert NO.fun_b(0.5) == 0.0 and NO.fun_b(2.0) < 1e-12 and abs(NO.fun_b(1.0) - 1.0) < 1e-12


def test_level_sizes_and_pairs():
    assert [NO.level_nside(j) for j in range(4)] == [1, 2, 4, 8]
    assert [NO.level_range(j, 16) for j in range(4)] == [(1, 2), (1, 4), (2, 8), (4, 16)]
    pair, use = NO.spneedlet_pair(1)
    assert len(pair) == 60 and len(use) == 30 and all(pair[pair[i]] == i for i in range(60))


def test_projection_reconstruction_of_a_band_limited_map():
    # the constant function is Y_00 * sqrt(4 pi): projecting it on [Y_00 | needlets] gives ~sqrt(4 pi) on column 0 and ~0 elsewhere
    theta, phi = NO.pano_grid(16, 32)
    SN = NO.needlet_matrix(theta, phi, 1)
    omega = NO.solid_angle_map(32).reshape(-1)
    assert abs(omega.sum() - 4 * np.pi) < 1e-10
    coef = NO.project(np.ones((16 * 32, 3)), SN, omega)
    # (approximately: the reference evaluates the basis on an endpoint-inclusive grid but integrates with half-pixel solid angles)
    assert abs(coef[0, 0] - np.sqrt(4 * np.pi)) < 0.05 and np.abs(coef[1:]).max() < 0.15
    sp = NO.sparsify(np.arange(30.0).reshape(10, 3) - 14, level_slices=((6, None), (2, 6)), frac=0.5)
    assert sp[:2].tolist() == [[-14, -13, -12], [-11, -10, -9]] and (sp[6:] != 0).sum() == 8


def test_product_host_tables_match_oracle():
    from emlight_b200 import needlets as PN
    for nside in (1, 2, 4, 8):
        assert np.abs(PN.healpix_centres(nside) - NO.pix2vec(nside).T).max() < 1e-14
    assert np.abs(PN.level_coefficients(3) - NO.level_coefficients(3)).max() < 1e-15
    assert PN.spneedlet_pair(1) == tuple(list(map(int, x)) for x in NO.spneedlet_pair(1))
    assert np.array_equal(PN.getSolidAngleMap(256), NO.solid_angle_map(256))
    pts, lev = PN.cubature_points(3)
    assert pts.shape == (1020, 3) and np.bincount(lev).tolist() == [12, 48, 192, 768]
    th, ph = PN.pano_grid()
    tho, pho = NO.pano_grid()
    assert np.array_equal(th, tho) and np.array_equal(ph, pho)


def test_healpix_known_answers_from_the_healpy_documentation():
    """Published outputs of the third-party dependency the reference calls (healpy, unpinned; sphere_needlets.py:5,52-54): the
    `healpy.pix2ang` documentation example  hp.pix2ang(16, [1440, 427, 1520, 0, 3071]) ->
        theta = [1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608], phi = [0., 0.78539816, 1.61988371, 0.78539816, 5.49778714]
    and the tutorial's  hp.pix2ang(16, 1440) -> (1.5291175943723188, 0.0).  Both the oracle's restatement of the RING scheme and the
    product's closed form must reproduce them (the needlet levels use nside 1..8 of the same formulas)."""
    from emlight_b200 import needlets as N
    pix = [1440, 427, 1520, 0, 3071]
    want_th = np.array([1.52911759, 0.78550497, 1.57079633, 0.05103658, 3.09055608])
    want_ph = np.array([0., 0.78539816, 1.61988371, 0.78539816, 5.49778714])
    th, ph = NO.pix2ang(16)
    assert np.abs(th[pix] - want_th).max() < 5e-9 and np.abs(ph[pix] - want_ph).max() < 5e-9
    assert th[1440] == 1.5291175943723188 and ph[1440] == 0.0
    v = N.healpix_centres(16)                                   # (npix, 3), what healpy.pix2vec returns
    th_p = np.arccos(v[:, 2])
    ph_p = np.mod(np.arctan2(v[:, 1], v[:, 0]), 2 * np.pi)
    assert np.abs(th_p[pix] - want_th).max() < 5e-9 and np.abs(ph_p[pix] - want_ph).max() < 5e-9
