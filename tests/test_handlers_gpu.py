"""GPU: the `data.ParameterDataset` / `util.tonemapping` shims (tone mapping on the kernel) against the reference arithmetic restated
in numpy.  PENDING FIRST B200 RUN like tests/test_gp_train_gpu.py: runs only with EML_PENDING_GPU=1 (`tools/gpu_pending.sh`)."""
import importlib.util
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("EML_PENDING_GPU") != "1", reason="not yet run on a B200 (set EML_PENDING_GPU=1)")]


def _np_tonemap(img, percentile, max_mapping, gamma=2.4):
    p = np.power(img, 1 / gamma)                                   # util.py:44-66 / :187-200
    nz = p > 0
    r = np.percentile(p[nz], percentile) if nz.any() else np.percentile(p, percentile)
    alpha = max_mapping / (r + 1e-10)
    return np.clip(alpha * p, 0, 1), alpha


def test_dataset_items_and_tonemapping(cuda, tmp_path):
    from emlight_b200 import handlers, wire
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin", "data.py")
    spec = importlib.util.spec_from_file_location("_dropin_data_gpu", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    os.makedirs(tmp_path / "pkl")
    os.makedirs(tmp_path / "crop")
    crop = np.exp(rng.normal(-1, 1, (48, 64, 3))).astype(np.float32)
    gt = {"distribution": rng.random(96).astype(np.float32), "intensity": np.float32(7.5), "rgb_ratio": rng.random(3).astype(np.float32),
          "ambient": rng.random(3).astype(np.float32)}
    wire.write_exr(str(tmp_path / "crop" / "im0.exr"), crop)
    with open(tmp_path / "pkl" / "im0.pickle", "wb") as f:
        pickle.dump(gt, f)
    ds = mod.ParameterDataset(str(tmp_path) + "/")
    item = ds[0]
    want, alpha = _np_tonemap(crop, 50, 0.5)
    assert item["name"] == "im0" and item["crop"].shape == (3, 48, 64) and item["crop"].is_cuda
    assert np.abs(item["crop"].cpu().numpy() - want.transpose(2, 0, 1)).max() <= 1e-5
    assert abs(float(item["intensity"]) - 7.5 * alpha / 500) <= 1e-5 * 7.5 * alpha / 500
    assert np.allclose(item["ambient"].cpu().numpy(), gt["ambient"] * alpha / (128 * 256), rtol=1e-5)
    assert torch.equal(item["distribution"].cpu(), torch.from_numpy(gt["distribution"]))
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=1)))
    assert batch["crop"].shape == (1, 3, 48, 64)
    tm = handlers.tonemapping(crop)
    assert isinstance(tm, np.ndarray) and np.abs(tm - _np_tonemap(crop, 99, 0.8)[0]).max() <= 1e-5
