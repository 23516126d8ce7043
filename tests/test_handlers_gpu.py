"""GPU: the `data.ParameterDataset` / `util.tonemapping` shims (tone mapping on the kernel) against the reference arithmetic restated
in numpy."""
import importlib.util
import os
import pickle

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu]


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


def test_genprojector_dataset_item(cuda, tmp_path):
    """LavalIndoorDataset item against GenProjector/data.py:59-107 restated with the render oracle."""
    import argparse
    import emlight_b200 as E
    from emlight_b200 import wire
    from oracle import render_oracle as RO
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin_genprojector", "data.py")
    spec = importlib.util.spec_from_file_location("_dropin_gp_data_gpu", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(2)
    for d in ("pkl", "warped", "crop"):
        os.makedirs(tmp_path / d)
    crop = np.exp(rng.normal(-1, 1, (96, 128, 3))).astype(np.float32)
    pano = np.exp(rng.normal(-2, 1, (128, 256, 3))).astype(np.float32)
    pano[20:24, 100:110] += 300
    dist = rng.random(128).astype(np.float32)
    dist /= dist.sum()
    pkl = {"distribution": dist, "intensity": np.float32(420.0), "rgb_ratio": np.array([0.6, 0.6, 0.52], np.float32),
           "ambient": np.array([900.0, 800.0, 700.0], np.float32)}
    wire.write_exr(str(tmp_path / "crop" / "s.exr"), crop)
    wire.write_exr(str(tmp_path / "warped" / "s.exr"), pano)
    with open(tmp_path / "pkl" / "s.pickle", "wb") as f:
        pickle.dump(pkl, f)
    item = mod.LavalIndoorDataset(argparse.Namespace(dataroot=str(tmp_path)))[0]
    _, alpha = _np_tonemap(crop, 50, 0.5)
    assert item["name"] == "s" and item["input"].shape == (3, 128, 256) and item["crop"].shape == (3, 128, 128)
    assert item["warped"].shape == (3, 128, 256) and item["map"].shape == (1, 128, 256)
    assert np.allclose(item["warped"].numpy(), pano.transpose(2, 0, 1) * alpha, rtol=1e-5)
    dirs = torch.from_numpy(E.sphere_points(128)).float().view(1, -1)
    colors = (torch.from_numpy(dist).view(1, 128, 1) * (420.0 * 0.01) * torch.from_numpy(pkl["rgb_ratio"]).view(1, 1, 3)).reshape(1, -1)
    want = (RO.convert_to_panorama_torch(dirs, torch.full((1, 128), 0.0025), colors)[0]
            + torch.from_numpy(pkl["ambient"]).view(3, 1, 1) / (128 * 256)) * alpha
    assert float((item["input"].cpu() - want).abs().max()) <= 1e-3 * float(want.abs().max())
    assert torch.equal(item["distribution"].cpu(), torch.from_numpy(dist).view(1, 128, 1).repeat(1, 1, 3))


def test_save_test_images_writes_the_reference_outputs(cuda, tmp_path):
    """GenProjector/test.py:30-39 output stage: the HDR result as EXR (bit-exact round trip) plus the tone-mapped previews."""
    from collections import OrderedDict
    from PIL import Image
    from emlight_b200 import handlers, wire
    gen = torch.Generator().manual_seed(4)
    fake = (torch.rand(3, 128, 256, generator=gen) * 50).to(cuda)
    images = OrderedDict([("input", torch.rand(3, 128, 256, generator=gen).to(cuda)), ("fake_image", fake),
                          ("warped", torch.rand(3, 128, 256, generator=gen).to(cuda) * 20), ("im", torch.rand(3, 128, 128, generator=gen).to(cuda))])
    out = str(tmp_path / "results")
    handlers.save_test_images(images, "scene7", out_dir=out)
    assert sorted(os.listdir(out)) == ["scene7_fake_image.exr", "scene7_fake_image.jpg", "scene7_input.jpg", "scene7_warped.jpg"]
    assert np.array_equal(wire.load_exr(os.path.join(out, "scene7_fake_image.exr")), fake.permute(1, 2, 0).cpu().numpy())
    prev = np.asarray(Image.open(os.path.join(out, "scene7_fake_image.jpg")))
    want, _ = _np_tonemap(fake.permute(1, 2, 0).cpu().numpy(), 50, 0.5)
    # JPEG is lossy (white noise loses ~25 levels on average): compare with the same 8-bit image pushed through the same encoder
    import io
    buf = io.BytesIO()
    Image.fromarray((want * 255.0).astype("uint8")).save(buf, format="JPEG")
    same = np.asarray(Image.open(io.BytesIO(buf.getvalue())))
    assert prev.shape == (128, 256, 3) and np.abs(prev.astype(np.float32) - same.astype(np.float32)).mean() < 1.0
