"""CPU oracle for the spherical-needlet basis, projection and reconstruction (TEST INFRASTRUCTURE ONLY).

PARITY: pinned against the reference's own files, except for healpy.  ``oracle/make_golden_needlets.py`` imports
/root/reference/Needlets/sphere_needlets.py, sphere_harmonics.py and utils.py in place (shims: an ``lpmn`` with the documented contract
over scipy's ``assoc_legendre_p_all`` because scipy 1.18 removed it; a ``healpy`` stub) and runs ``SNvertex`` at jmax = 3 on 24 points
of the reference's 128x256 grid and at jmax = 2 on a whole 16x32 grid, then exec's the script lines gt_gen_j3.py:39-43,
mat_gen2.py:43-51 and :55 on those arrays -> tests/golden/needlets.npz.  This file reproduces all of it to <= 2e-14
(tests/test_needlets_cpu.py::test_oracle_matches_reference_golden).  What stays UNPINNED is third-party and absent from this image:
``healpy`` (unpinned -- the reference has no requirements file): its ``pix2ang / pix2vec / ringinfo`` are restated here from the
*published* HEALPix RING pixelisation (Gorski et al. 2005, ApJ 622:759, eqs. 2-9) and the golden run answers healpy's calls from
this restatement; it is checked against the healpy documentation's own example outputs (nside = 16, 5 pixels to 5e-9, one to the
last bit) and the scheme's symmetries only.  The pre-computed ``SN_Matrix3.npy``
is not shipped by the reference.  ``scipy.special.lpmv`` stands in for ``lpmn`` below (same Condon-Shortley convention).

Reference anchors (all under /root/reference/Needlets):
* ``fun_b / compute_f2 / compute_f3``   sphere_needlets.py:10-29     window b(x) = sqrt(f3(x/B) - f3(x)), C-infinity bump via quad
* ``spneedlet``                         sphere_needlets.py:34-104    per level j: Nside_j, cubature weight lambda_j = 4 pi / Npix_j,
                                                                     inverse SH transform ring by ring
* ``spneedlet_eval``                    sphere_needlets.py:182-191   coef[l, m+lmax] = conj(Y_lm(theta, phi)), lmax = floor(B^(jmax+1))
* ``spneedlet_pair``                    sphere_needlets.py:107-127   antipodal cubature pairs
* ``SNvertex``                          sphere_needlets.py:196-238   SN_matrix = [Y_00 | psi_0 | ... | psi_jmax]
* ``spharmonic_eval``                   sphere_harmonics.py:77-89
* ``getSolidAngle / getSolidAngleMap``  utils.py:35-50
* projection / reconstruction / sparsification   gt_gen_j3.py:39-43, mat_gen2.py:36-41,43-51,55; grid mat_gen2.py:22-25

Closed form (what the GPU kernel evaluates): sum_m conj(Y_lm(x)) Y_lm(x_k) = (2l+1)/(4 pi) P_l(x . x_k), hence
    psi_jk(x) = sqrt(lambda_j) * sum_{l=l_st(j)}^{l_en(j)} b(l / B^j) (2l+1)/(4 pi) P_l(x . xi_jk).
"""
import math

import numpy as np
from scipy.integrate import quad
from scipy.special import lpmv

PANO_H, PANO_W = 128, 256


# ----------------------------------------------------------------------------------------------------------- window function
def _bump(x):
    return np.exp(-1.0 / (1.0 - x ** 2))


def compute_f2(u):                                   # sphere_needlets.py:10-12
    return quad(_bump, -1, u + 1e-10)[0] / quad(_bump, -1, 1)[0]


def compute_f3(x, B):                                # sphere_needlets.py:15-23
    if x <= 1.0 / B:
        return 1.0
    if x <= 1:
        return compute_f2(1 - 2 * B / (B - 1) * (x - 1 / B))
    return 0.0


def fun_b(x, B=2.0):                                 # sphere_needlets.py:28-29
    return np.sqrt(compute_f3(x / B, B) - compute_f3(x, B))


def level_nside(j, B=2.0):                           # sphere_needlets.py:48
    return 2 ** math.ceil(math.log(math.floor(B ** (j + 1)) / 2, 2))


def level_range(j, lmax, B=2.0):                     # sphere_needlets.py:73-74
    return int(np.ceil(B ** (j - 1))), int(min(np.floor(B ** (j + 1)), lmax))


def b_vector(jmax, lmax, BW=2.0):                    # sphere_needlets.py:39-43
    bv = np.zeros((jmax + 1, lmax))
    for j in range(jmax + 1):
        for l in range(1, lmax + 1):
            bv[j, l - 1] = fun_b(l / BW ** j, BW)
    return bv


# ------------------------------------------------------------------------------------- HEALPix RING scheme (third party: healpy)
def healpix_rings(nside):
    """Ring table of the RING scheme: for ring r = 1 .. 4 nside - 1 -> (startpix, npix_in_ring, z, phi of first pixel, dphi).
    Gorski et al. 2005 eqs. 4-9."""
    rows = []
    start = 0
    for r in range(1, 4 * nside):
        if r < nside:                                # north polar cap
            n, z = 4 * r, 1.0 - r * r / (3.0 * nside * nside)
            phi0 = 0.5 * np.pi / (2.0 * r)
        elif r <= 3 * nside:                         # equatorial belt
            n, z = 4 * nside, (2.0 * nside - r) * 2.0 / (3.0 * nside)
            s = 1.0 if ((r + nside) & 1) else 0.5    # healpy: fodd = 0.5 * (1 + ((ring + nside) & 1)), phi = (j - fodd) pi / (2 nside)
            phi0 = (1.0 - s) * np.pi / (2.0 * nside)
        else:                                        # south polar cap
            rr = 4 * nside - r
            n, z = 4 * rr, -1.0 + rr * rr / (3.0 * nside * nside)
            phi0 = 0.5 * np.pi / (2.0 * rr)
        rows.append((start, n, z, phi0, 2.0 * np.pi / n))
        start += n
    assert start == 12 * nside * nside
    return rows


def pix2ang(nside):
    """(theta, phi) of every RING-ordered pixel centre (what healpy.pix2ang(nside, range(npix)) returns)."""
    th, ph = [], []
    for start, n, z, phi0, dphi in healpix_rings(nside):
        th.extend([math.acos(z)] * n)
        ph.extend(phi0 + dphi * k for k in range(n))
    return np.array(th), np.array(ph)


def pix2vec(nside):
    th, ph = pix2ang(nside)
    return np.stack((np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)))      # (3, npix)


def all_centres(jmax, B=2.0):
    """(3, sum_j Npix_j) cubature points of levels 0..jmax concatenated (sphere_needlets.py:109-116)."""
    return np.hstack([pix2vec(level_nside(j, B)) for j in range(jmax + 1)])


def spneedlet_pair(jmax, B=2.0):                     # sphere_needlets.py:107-127
    pix = all_centres(jmax, B)
    corr = pix.T.dot(pix)
    pair, use = [], []
    for i in range(pix.shape[1]):
        p = np.where(corr[i] + 1 < 1e-10)[0][0]
        pair.append(p)
        if p > i:
            use.append(i)
    return pair, use


# ------------------------------------------------------------------------------------------------ line-by-line transcription
def _fact(n):
    return float(math.factorial(n))


def spharmonic_eval(l, m, theta, phi):               # sphere_harmonics.py:77-89 (m >= 0 is all the hot path uses)
    sign_m = np.sign(m)
    m = abs(m)
    C = np.sqrt((2 * l + 1) / (4 * np.pi) * _fact(l - m) / _fact(l + m))
    Y = C * lpmv(m, l, np.cos(theta)) * np.exp(1j * m * phi)
    if sign_m < 0:
        Y = (-1) ** m * np.conjugate(Y)
    return Y


def spneedlet(coef, lmax, jmax, B=2.0):
    """sphere_needlets.py:34-104 restated step by step (ring-wise inverse spherical-harmonic transform)."""
    beta = {}
    bv = b_vector(jmax, lmax)
    for j in range(jmax + 1):
        nside = level_nside(j, B)
        npix = 12 * nside ** 2
        lamb = 4 * np.pi / npix
        rings = healpix_rings(nside)
        nring = 4 * nside - 1
        startpix = [r[0] for r in rings] + [npix]
        _, phis = pix2ang(nside)
        thetas = np.array([math.acos(rings[i][2]) for i in range(2 * nside)])            # north cap + equator ring (:53)
        pre = {}
        for l in range(1, lmax + 1):                                                     # :58-66
            norm = np.array([(-1) ** m * np.sqrt((l + 0.5) * _fact(l - m) / _fact(l + m)) for m in range(l + 1)])
            tm = np.zeros((l + 1, len(thetas)))
            for i in range(len(thetas)):
                tm[:, i] = np.array([lpmv(m, l, np.cos(thetas[i])) for m in range(l + 1)]) * norm
            tm2 = (np.fliplr(tm[:, :len(thetas) - 1]).T * (-1.0) ** (l + np.arange(l + 1))).T
            pre[l] = np.hstack((tm, tm2))
        l_st, l_en = level_range(j, lmax, B)
        alm = coef.copy()
        for l in range(l_st, l_en + 1):                                                  # :76-78
            alm[l, lmax:l + lmax + 1] *= bv[j, l - 1] * np.sqrt(lamb)
        beta[j] = np.zeros(npix)
        tmat = np.stack([pre[l][0, :] for l in range(l_st, l_en + 1)])
        term1 = np.conjugate(alm[l_st:l_en + 1, lmax]).dot(tmat) / np.sqrt(2 * np.pi)    # :83-86
        tmat2 = np.zeros((l_en, nring), dtype=complex)
        for m in range(1, l_en + 1):                                                     # :88-94
            l2 = max(m, l_st)
            tmm = np.stack([pre[l][m, :] for l in range(l2, l_en + 1)])
            tmat2[m - 1, :] = alm[l2:l_en + 1, m + lmax].dot(tmm) / np.sqrt(2 * np.pi) * (-1) ** m
        for r in range(nring):                                                           # :96-101
            for k in range(startpix[r], startpix[r + 1]):
                vec = np.exp(np.arange(1, l_en + 1) * 1j * phis[k])
                beta[j][k] = term1[r].real + 2 * vec.dot(tmat2[:, r]).real
    return beta


def spneedlet_eval(theta, phi, jmax, B=2.0):         # sphere_needlets.py:182-191
    lmax = int(np.floor(B ** (jmax + 1)))
    coef = np.zeros((lmax + 1, 2 * lmax + 1), dtype=complex)
    for l in range(1, lmax + 1):
        for m in range(l + 1):
            coef[l, m + lmax] = np.conjugate(spharmonic_eval(l, m, theta, phi))
    return spneedlet(coef, lmax, jmax, B)


def SNvertex_direct(theta, phi, jmax, B=2.0):
    """sphere_needlets.py:196-238 for a (small) list of points, through the transcription above. Returns SN_matrix."""
    rows = []
    for t, p in zip(theta, phi):
        sn = spneedlet_eval(t, p, jmax, B)
        rows.append(np.hstack([sn[j] for j in range(jmax + 1)]))
    sh00 = np.array([spharmonic_eval(0, 0, t, p).real for t, p in zip(theta, phi)]).reshape(-1, 1)
    return np.hstack((sh00, np.array(rows)))


# ------------------------------------------------------------------------------------------------------------- closed form
def level_coefficients(jmax, B=2.0):
    """c[j, l] = sqrt(lambda_j) b(l/B^j) (2l+1)/(4 pi) for l_st <= l <= l_en, else 0; shape (jmax+1, lmax+1)."""
    lmax = int(np.floor(B ** (jmax + 1)))
    bv = b_vector(jmax, lmax)
    c = np.zeros((jmax + 1, lmax + 1))
    for j in range(jmax + 1):
        lamb = 4 * np.pi / (12 * level_nside(j, B) ** 2)
        l_st, l_en = level_range(j, lmax, B)
        for l in range(l_st, l_en + 1):
            c[j, l] = np.sqrt(lamb) * bv[j, l - 1] * (2 * l + 1) / (4 * np.pi)
    return c


def needlet_matrix(theta, phi, jmax, B=2.0):
    """SN_matrix (npoints, 1 + sum_j Npix_j), float64, by the addition theorem + the Legendre three-term recurrence."""
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    x = np.stack((np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)), 1)      # (P, 3)
    c = level_coefficients(jmax, B)
    lmax = c.shape[1] - 1
    cols = [np.full((len(theta), 1), 1.0 / np.sqrt(4 * np.pi))]
    for j in range(jmax + 1):
        t = np.clip(x @ pix2vec(level_nside(j, B)), -1.0, 1.0)                                      # (P, Npix_j)
        p0, p1 = np.ones_like(t), t
        acc = c[j, 1] * p1
        for l in range(2, lmax + 1):
            p0, p1 = p1, ((2 * l - 1) * t * p1 - (l - 1) * p0) / l
            if c[j, l] != 0.0:
                acc = acc + c[j, l] * p1
        cols.append(acc)
    return np.hstack(cols)


def pano_grid(h=PANO_H, w=PANO_W):
    """theta (rows) / phi (columns) of mat_gen2.py:22-25: endpoint-inclusive linspace, flattened row-major."""
    X, Y = np.meshgrid(np.linspace(0, 2, w) * np.pi, np.linspace(0, 1, h) * np.pi)
    return Y.reshape(-1), X.reshape(-1)


# ------------------------------------------------------------------------------------------- projection / reconstruction
def solid_angle_map(width=PANO_W):                   # utils.py:35-50
    height = width // 2
    y = np.arange(0, height)
    theta = (1.0 - ((y + 0.5) / height)) * np.pi
    sa = (np.pi * 2) / width * (np.cos(theta - (np.pi / height / 2.0)) - np.cos(theta + (np.pi / height / 2.0)))
    return np.repeat(sa[:, None], width, axis=1)


def project(pano, SN, omega):
    """gt_gen_j3.py:39-43 / mat_gen2.py:36-41: pano (P,3) -> (nCoeffs,3), coef[i,ch] = sum_p pano[p,ch] SN[p,i] omega[p]."""
    return (SN * omega[:, None]).T @ pano


def sparsify(coef, level_slices=((253, None), (61, 253)), frac=0.1):     # mat_gen2.py:43-51 (j=3 then j=2)
    out = coef.copy()
    for lo, hi in level_slices:
        blk = out[lo:hi]
        out[lo:hi] = blk * (np.abs(blk) > np.abs(blk).max() * frac)
    return out


def reconstruct(SN, coef):                            # mat_gen2.py:55
    return SN @ coef
