"""CPU oracle for TonemapHDR (TEST INFRASTRUCTURE ONLY): numpy restatement of RegressionNetwork/util.py:36-66, pinned against the
reference class itself (exec'd from those lines) by oracle/make_golden_tonemap.py -> tests/golden/tonemap.npz."""
import numpy as np


def tonemap_hdr(img, gamma=2.4, percentile=50, max_mapping=0.5, clip=True, alpha=None, use_gamma=True):
    p = np.power(img, 1 / gamma) if use_gamma else img                      # :50-53
    nz = p > 0
    r = np.percentile(p[nz], percentile) if nz.any() else np.percentile(p, percentile)   # :54-58
    if alpha is None:
        alpha = max_mapping / (r + 1e-10)                                   # :59-60
    out = np.multiply(alpha, p)
    if clip:
        out = np.clip(out, 0, 1)                                            # :63-64
    return out.astype("float32"), alpha
