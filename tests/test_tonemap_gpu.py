"""GPU: eml_tonemap_hdr (per-image radix-select percentile) through emlight_b200.tonemap.TonemapHDR vs the reference-generated golden
and the oracle; batch of 64 crops at the network's input size."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import tonemap_oracle as TO
from oracle.make_golden_tonemap import synthetic_crop

pytestmark = pytest.mark.gpu


def test_tonemap_matches_reference_golden(cuda):
    from emlight_b200.tonemap import TonemapHDR
    g = np.load(os.path.join(GOLDEN, "tonemap.npz"))
    for i in range(3):
        pct, mm = g["cfg_%d" % i]
        tone = TonemapHDR(gamma=2.4, percentile=float(pct), max_mapping=float(mm))
        y, alpha = tone(torch.from_numpy(synthetic_crop(40 + i, zeros=0.0 if i == 2 else 0.1)).to(cuda))
        want = float(g["alpha_%d" % i])
        assert isinstance(alpha, float) and abs(alpha - want) <= 1e-5 * want          # powf vs numpy pow: a few ulp on the selected value
        assert y.dtype == torch.float32 and np.abs(y.cpu().numpy()[::3, ::3] - g["y_%d" % i]).max() <= 1e-5
    y, a = TonemapHDR()(torch.from_numpy(synthetic_crop(40)).to(cuda), clip=False, alpha=0.7, gamma=False)
    assert a == 0.7 and np.abs(y.cpu().numpy()[::3, ::3] - g["y_given"]).max() <= 1e-6 * np.abs(g["y_given"]).max()


def test_tonemap_batch_matches_oracle(cuda):
    from emlight_b200.tonemap import TonemapHDR
    B = 64
    imgs = np.stack([synthetic_crop(100 + b, h=192, w=256, zeros=0.05 * (b % 3)) for b in range(B)])
    imgs[5] = 0.0                                                                     # no positive value at all: alpha = 0.5 / 1e-10
    y, alpha = TonemapHDR()(torch.from_numpy(imgs).to(cuda))
    assert tuple(y.shape) == (B, 192, 256, 3) and tuple(alpha.shape) == (B,)
    for b in (0, 1, 2, 5, 33, 63):
        ry, ra = TO.tonemap_hdr(imgs[b])
        assert abs(float(alpha[b]) - ra) <= 1e-5 * ra
        assert np.abs(y[b].cpu().numpy() - ry).max() <= 1e-5
    assert float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    # the median of the positive tonemapped values is max_mapping by construction
    v = y[0][y[0] > 0]
    assert abs(float(v.median()) - 0.5) < 1e-3


def test_tonemap_numpy_call_like_train_py(cuda):
    """RegressionNetwork/train.py:124,135: `tone(env_pred)[0].transpose((1, 2, 0)).astype('float32') * 255.0` on a numpy (3,128,256)
    array -- numpy in, (numpy float32, float alpha) out, computed by the kernel."""
    from emlight_b200.tonemap import TonemapHDR
    env = np.ascontiguousarray(synthetic_crop(77, h=128, w=256).transpose(2, 0, 1)) * 40.0
    y, alpha = TonemapHDR(gamma=2.4, percentile=50, max_mapping=0.5)(env)
    ry, ra = TO.tonemap_hdr(env)
    assert isinstance(y, np.ndarray) and y.dtype == np.float32 and y.shape == (3, 128, 256) and isinstance(alpha, float)
    assert abs(alpha - ra) <= 1e-5 * ra and np.abs(y - ry).max() <= 1e-5
    img = y.transpose((1, 2, 0)).astype('float32') * 255.0
    assert img.shape == (128, 256, 3) and img.max() <= 255.0
