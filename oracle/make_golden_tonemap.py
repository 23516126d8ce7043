"""Golden vectors for TonemapHDR: exec's the reference class from RegressionNetwork/util.py:36-66 (the file itself is not importable:
merge-conflict markers, OpenEXR) on synthetic HDR crops and writes tests/golden/tonemap.npz.  python oracle/make_golden_tonemap.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_crop(seed, h=96, w=128, zeros=0.1):
    rng = np.random.default_rng(seed)
    img = np.exp(rng.normal(-1.0, 1.5, (h, w, 3))).astype(np.float32)       # log-normal radiance
    img[rng.random((h, w, 3)) < zeros] = 0.0                                 # exact zeros are excluded from the percentile
    return img


def main():
    src = open("/root/reference/RegressionNetwork/util.py").read().split("\n")
    ns = {"np": np}
    exec("\n".join(src[35:66]), ns)                                          # class TonemapHDR
    out = {}
    for i, (pct, mm) in enumerate(((50, 0.5), (99, 0.9), (50, 0.5))):
        tone = ns["TonemapHDR"](gamma=2.4, percentile=pct, max_mapping=mm)
        img = synthetic_crop(40 + i, zeros=0.0 if i == 2 else 0.1)
        y, alpha = tone(img)
        out["alpha_%d" % i] = np.float64(alpha); out["y_%d" % i] = y[::3, ::3]
        out["cfg_%d" % i] = np.array([pct, mm])
    y, a = ns["TonemapHDR"]()(synthetic_crop(40), clip=False, alpha=0.7, gamma=False)
    out["y_given"] = y[::3, ::3]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tonemap.npz"), **out)
    print("wrote tonemap.npz", [float(out["alpha_%d" % i]) for i in range(3)])


if __name__ == "__main__":
    main()
