"""Golden vectors for the GT parameter extraction: exec's the reference's own `extract_mesh` class
(RegressionNetwork/representation/distribution_representation.py:65-120) and the two helpers it takes from representation/util.py
(:184-203) -- the files themselves import vtk / cv2 / imageio and run a dataset loop at import time -- on synthetic HDR panoramas and
writes tests/golden/extract.npz.  Run from the repo root: python oracle/make_golden_extract.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference/RegressionNetwork/representation/"


def synthetic_pano(seed, h=128, w=256):
    """A few bright lobes on a dim textured background (HDR range), float32 (h, w, 3)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = 0.05 * rng.random((h, w, 3)) + 0.02
    for _ in range(4):
        cy, cx, s = rng.integers(10, h - 10), rng.integers(0, w), rng.uniform(2, 9)
        col = rng.uniform(0.3, 1.0, 3) * rng.uniform(5, 300)
        dx = np.minimum(np.abs(xx - cx), w - np.abs(xx - cx))
        img += col * np.exp(-((yy - cy) ** 2 + dx ** 2) / (2 * s * s))[..., None]
    return img.astype(np.float32)


def main():
    usrc = open(REF + "util.py").read().split("\n")
    util = types.ModuleType("util")
    util.np = np
    exec("\n".join(usrc[183:203]), util.__dict__)          # polar_to_cartesian, sphere_points
    dsrc = open(REF + "distribution_representation.py").read().split("\n")
    ns = {"np": np, "util": util}
    exec("\n".join(dsrc[64:120]), ns)                       # class extract_mesh
    out = {}
    for ln in (64, 128):
        ex = ns["extract_mesh"](ln=ln)
        out["idx_%d" % ln] = ex.idx.astype(np.int16)
        for seed in (0, 1):
            hdr = synthetic_pano(10 * ln + seed)
            pl, mp = ex.compute(hdr)
            tag = "%d_%d" % (ln, seed)
            out["dist_" + tag] = pl["distribution"]; out["int_" + tag] = np.float64(pl["intensity"])
            out["rgb_" + tag] = pl["rgb_ratio"]; out["amb_" + tag] = pl["ambient"]; out["map_" + tag] = np.packbits(mp)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "extract.npz"), **out)
    print("wrote extract.npz", {k: v.shape for k, v in out.items() if k.startswith("dist")})


if __name__ == "__main__":
    main()
