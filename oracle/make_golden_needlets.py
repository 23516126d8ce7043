"""Generates tests/golden/needlets.npz by RUNNING the reference's own Needlets/ files (TEST INFRASTRUCTURE ONLY).

    python oracle/make_golden_needlets.py            (needs /root/reference; about a minute)

What is executed unmodified from /root/reference/Needlets (read-only, imported in place, nothing copied):
* ``sphere_needlets.SNvertex`` -> ``spneedlet_pair`` / ``spneedlet_eval`` / ``spneedlet`` / ``fun_b`` (sphere_needlets.py:10-29, 34-104,
  107-127, 182-191, 196-238) and ``sphere_harmonics.spharmonic_eval`` (sphere_harmonics.py:77-89): the basis matrices below are
  the return values of the reference's function on our grid points;
* ``utils.getSolidAngleMap`` (utils.py:35-50);
* the projection loop ``gt_gen_j3.py:39-43``, the sparsification ``mat_gen2.py:43-51`` and the reconstruction ``mat_gen2.py:55``:
  those script lines are read from the files at run time and exec'd on our arrays (the scripts themselves need a dataset).

What has to be shimmed for the files to import in this image, and why it does not touch the arithmetic under test:
* ``scipy.special.lpmn`` was removed in scipy 1.18 -> a shim with lpmn's documented contract (returns P[m, n] = P_n^m(z) with the
  Condon-Shortley phase, and its derivative) built on ``scipy.special.assoc_legendre_p_all`` (scipy's own replacement);
* ``healpy`` (third party, unpinned, absent): ``ringinfo`` / ``pix2ang`` / ``pix2vec`` answer from oracle/needlets_oracle.py's
  restatement of the published RING pixelisation (Gorski et al. 2005).  THIS is the one piece that stays unpinned: no healpy
  build exists here to compare with; the restatement is checked against the scheme's published values and symmetries
  (tests/test_needlets_cpu.py::test_healpix_ring_scheme);
* ``OpenEXR`` / ``Imath`` (imported by Needlets/utils.py, unused by the functions run here): empty stub modules;
* numpy >= 2 rejects generator arguments to ``np.vstack`` / ``np.hstack`` (sphere_needlets.py:113,116,209,238): the modules'
  ``np`` name is a proxy that turns a generator into a list first -- same call, same data.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/Needlets"
sys.path.insert(0, ROOT)


class _NumpyCompat:
    """numpy with vstack / hstack accepting generators, as numpy < 1.16 did."""

    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def vstack(tup, *a, **k):
        return np.vstack(list(tup) if isinstance(tup, types.GeneratorType) else tup, *a, **k)

    @staticmethod
    def hstack(tup, *a, **k):
        return np.hstack(list(tup) if isinstance(tup, types.GeneratorType) else tup, *a, **k)


def _lpmn(m, n, z):
    """scipy.special.lpmn(m, n, z) as documented up to scipy 1.14: (P, dP), each (m+1, n+1), P[i, j] = P_j^i(z) with the
    Condon-Shortley phase; |z| = 1 handled like the old routine for what the reference reads (values only)."""
    import scipy.special as sp
    p = sp.assoc_legendre_p_all(n, m, float(z), branch_cut=2, diff_n=1)                 # (2, n+1, 2m+1)
    P = np.array([[p[0][j, i] for j in range(n + 1)] for i in range(m + 1)])
    dP = np.array([[p[1][j, i] for j in range(n + 1)] for i in range(m + 1)])
    return P, dP


def import_reference():
    """Imports the reference's Needlets modules under the shims described in the header; returns (sphere_needlets, utils)."""
    import scipy.special as sp
    from oracle import needlets_oracle as NO
    if not hasattr(sp, "lpmn"):
        sp.lpmn = _lpmn
    hp = types.ModuleType("healpy")

    def ringinfo(nside, ring):
        rows = NO.healpix_rings(nside)
        ring = np.atleast_1d(ring)
        start = np.array([rows[r - 1][0] for r in ring])
        npix = np.array([rows[r - 1][1] for r in ring])
        z = np.array([rows[r - 1][2] for r in ring])
        return start, npix, z, np.sqrt(1 - z * z), np.array([True] * len(ring))

    def pix2ang(nside, ipix):
        th, ph = NO.pix2ang(nside)
        idx = np.asarray(list(ipix))
        return th[idx], ph[idx]

    def pix2vec(nside, ipix):
        v = NO.pix2vec(nside)
        idx = np.asarray(list(ipix))
        return v[0][idx], v[1][idx], v[2][idx]

    hp.ringinfo, hp.pix2ang, hp.pix2vec = ringinfo, pix2ang, pix2vec
    sys.modules["healpy"] = hp
    for name in ("OpenEXR", "Imath"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, REF)
    saved = sys.modules.pop("utils", None)
    try:
        import sphere_needlets as SNmod
        import utils as Umod
    finally:
        sys.path.remove(REF)
        for name in ("utils", "sphere_harmonics", "sphere_needlets"):
            sys.modules.pop(name, None)
        if saved is not None:
            sys.modules["utils"] = saved
    SNmod.np = _NumpyCompat()
    return SNmod, Umod


def script_lines(fname, first, last):
    """Source lines first..last (1-based, inclusive) of a reference script, dedented, for exec on our arrays."""
    import textwrap
    with open(os.path.join(REF, fname)) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def main():
    SNmod, Umod = import_reference()
    rng = np.random.default_rng(20261017)
    quiet = io.StringIO()

    # (1) basis at jmax = 3 (the BASELINE configs[4] level) on points of the reference's 128x256 evaluation grid (mat_gen2.py:22-25):
    #     both poles, the seam columns phi = 0 and 2 pi, and a random subset
    pix1 = np.linspace(0, 1, 128) * np.pi
    pix2 = np.linspace(0, 2, 256) * np.pi
    rows = np.concatenate(([0, 127, 64, 64, 1], rng.integers(0, 128, 19)))
    cols = np.concatenate(([0, 255, 0, 255, 128], rng.integers(0, 256, 19)))
    theta3, phi3 = pix1[rows], pix2[cols]
    with contextlib.redirect_stdout(quiet):
        SN1_3, SN2_3, SN_3 = SNmod.SNvertex(theta=theta3, phi=phi3, jmax=3)
    assert SN_3.shape == (24, 1021) and SN1_3.shape == SN2_3.shape == (24, 511)

    # (2) the whole pipeline at jmax = 2 on a coarse 16x32 grid built exactly like mat_gen2.py:22-25
    h, w = 16, 32
    X, Y = np.meshgrid(np.linspace(0, 2, w) * np.pi, np.linspace(0, 1, h) * np.pi)
    X, Y = X.reshape(-1), Y.flatten().reshape(-1)
    with contextlib.redirect_stdout(quiet):
        _, _, SN_2 = SNmod.SNvertex(theta=Y, phi=X, jmax=2)
    assert SN_2.shape == (512, 253)
    omega = Umod.getSolidAngleMap(w)
    pano = np.exp(1.5 * rng.standard_normal((h * w, 3)))
    # gt_gen_j3.py:39-43 -- the coefficient loop (exr := our pano, SN_Matrix := the reference's own basis on this grid)
    ns = {"np": np, "exr": pano, "SN_Matrix": SN_2, "solidAngles": omega.reshape((-1)), "nCoeffs": SN_2.shape[1]}
    src = script_lines("gt_gen_j3.py", 39, 43)
    assert src.startswith("SN_Coeffs = np.zeros((nCoeffs, 3))") and "solidAngles" in src, src
    exec(src, ns)
    coef_2 = ns["SN_Coeffs"]
    # mat_gen2.py:55 -- reconstruction
    src = script_lines("mat_gen2.py", 55, 55)
    assert src.strip() == "rec = np.dot(SN_Matrix, SN_Coeffs)", src
    exec(src, ns)
    rec_2 = ns["rec"]

    # (3) sparsification, mat_gen2.py:43-51 (hard-wired j = 3 / j = 2 column ranges 253: and 61:253 -> needs 1021 coefficients)
    sp_in = rng.standard_normal((1021, 3)) * np.exp(rng.standard_normal((1021, 1)))
    ns = {"np": np, "SN_Coeffs": sp_in.copy()}
    src = script_lines("mat_gen2.py", 43, 51)
    assert src.startswith("j3 = SN_Coeffs[253:, :]") and "SN_Coeffs[61:253, :] = j2 * mask" in src, src
    with contextlib.redirect_stdout(quiet):
        exec(src, ns)
    sp_out = ns["SN_Coeffs"]

    # (4) window function and pair table straight from the reference
    xs = np.array([0.5, 0.51, 0.6, 0.75, 0.9, 1.0, 1.2, 1.5, 1.9, 1.99, 2.0])
    bvals = np.array([SNmod.fun_b(x, 2.0) for x in xs])
    pair, use = SNmod.spneedlet_pair(3)

    out = os.path.join(ROOT, "tests", "golden", "needlets.npz")
    np.savez_compressed(out, theta3=theta3, phi3=phi3, SN_3=SN_3, SN1_3=SN1_3, SN2_3=SN2_3,
                        theta2=Y, phi2=X, SN_2=SN_2, omega2=omega, pano2=pano, coef2=coef_2, rec2=rec_2,
                        sp_in=sp_in, sp_out=sp_out, b_x=xs, b_val=bvals, pair3=np.array(pair), use3=np.array(use))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
