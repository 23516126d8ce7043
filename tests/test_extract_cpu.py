"""CPU: the GT-parameter-extraction oracle vs golden vectors produced by the reference's own `extract_mesh` class
(oracle/make_golden_extract.py exec's distribution_representation.py:65-120), and the product's host-side LUT vs the golden LUT."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import extract_oracle as XO
from oracle.make_golden_extract import synthetic_pano


def test_extract_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "extract.npz"))
    for ln in (64, 128):
        ex = XO.ExtractMesh(ln=ln)
        assert np.array_equal(ex.idx, g["idx_%d" % ln])
        for seed in (0, 1):
            pl, mp = ex.compute(synthetic_pano(10 * ln + seed))
            tag = "%d_%d" % (ln, seed)
            assert np.allclose(pl["distribution"], g["dist_" + tag], rtol=1e-10, atol=1e-14)
            assert abs(pl["intensity"] - float(g["int_" + tag])) <= 1e-10 * float(g["int_" + tag])
            assert np.allclose(pl["rgb_ratio"], g["rgb_" + tag], rtol=1e-10) and np.allclose(pl["ambient"], g["amb_" + tag], rtol=1e-10)
            assert np.array_equal(np.packbits(mp), g["map_" + tag])
            assert abs(pl["distribution"].sum() - 1) < 1e-12 and abs(np.linalg.norm(pl["rgb_ratio"]) - 1) < 1e-12


def test_product_lut_matches_reference_golden():
    import emlight_b200.representation as R
    g = np.load(os.path.join(GOLDEN, "extract.npz"))
    for ln in (64, 128):
        ex = R.extract_mesh.__new__(R.extract_mesh)                     # host tables only (no CUDA in this test)
        import torch
        R.extract_mesh.__init__(ex, ln=ln, device=torch.device("cpu"))
        assert np.array_equal(ex.idx, g["idx_%d" % ln])
