"""CPU: TonemapHDR oracle vs golden vectors from the reference class (oracle/make_golden_tonemap.py exec's util.py:36-66)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle import tonemap_oracle as TO
from oracle.make_golden_tonemap import synthetic_crop


def test_tonemap_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "tonemap.npz"))
    for i in range(3):
        pct, mm = g["cfg_%d" % i]
        y, alpha = TO.tonemap_hdr(synthetic_crop(40 + i, zeros=0.0 if i == 2 else 0.1), percentile=int(pct), max_mapping=float(mm))   # python scalars like the reference call (numpy 2 promotes by scalar type)
        assert abs(float(alpha) - float(g["alpha_%d" % i])) <= 1e-7 * float(alpha)
        assert np.array_equal(y[::3, ::3], g["y_%d" % i])
    y, a = TO.tonemap_hdr(synthetic_crop(40), clip=False, alpha=0.7, use_gamma=False)
    assert a == 0.7 and np.array_equal(y[::3, ::3], g["y_given"])
