"""Host-side helpers behind the `util` / `data` module-name shims (`emlight_b200/handlers.py`, `dropin/data.py`) against golden
vectors produced by the reference's own code (`oracle/make_golden_handlers.py` exec's util.py:69-220)."""
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN
from emlight_b200 import handlers, wire
from oracle.make_golden_handlers import synthetic_pano

H = handlers.PanoramaHandler


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "handlers.npz"))


def test_panorama_handler_matches_reference(gold):
    pano = synthetic_pano()
    assert np.array_equal(H.rgb_to_intenisty(pano), gold["intensity"])
    assert np.array_equal(H.horizontal_rotate_panorama(pano, 77.0), gold["rot"])
    assert np.allclose(H.generate_steradian(64, 128), gold["ster"], rtol=1e-6, atol=0)
    assert np.allclose(H.generate_steradian(16, 32, multiply=False), gold["ster_raw"], rtol=1e-6, atol=0)
    work = pano.copy()
    gt, amb = H.prepare_gt_panorama(work)
    assert gt is work and np.array_equal(gt, gold["gt"]) and np.allclose(amb, gold["ambient"], rtol=1e-5)
    gt2, amb2 = H.prepare_gt_panorama(pano.copy(), threshold=1e-9)
    assert np.array_equal(gt2, gold["gt2"]) and np.array_equal(amb2, gold["ambient2"]) and not amb2.any()
    assert np.allclose(H.resize_panorama(pano, 16), gold["resized"], rtol=1e-6)
    assert np.allclose(H.resize_panorama(pano, (40, 24)), gold["resized_t"], rtol=1e-6)
    assert np.allclose(H.crop_panorama(pano, 60.0, crop_image_h=30), gold["crop"], rtol=1e-5, atol=1e-7)


def test_polar_conversions_match_reference(gold):
    phi, theta = handlers.cartesian_to_polar(gold["xyz"])
    assert np.array_equal(phi, gold["phi"]) and np.array_equal(theta, gold["theta"])
    back = handlers.polar_to_cartesian((phi, theta))
    assert np.array_equal(back, gold["back"]) and np.allclose(back.T, gold["xyz"], atol=1e-12)


def test_read_exr_and_read_hdr_go_through_the_wire_reader(tmp_path):
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(1)
    rgba = rng.random((9, 12, 4)).astype(np.float32) * 5
    path = str(tmp_path / "rgba.exr")
    assert cv2.imwrite(path, rgba[:, :, [2, 1, 0, 3]], [cv2.IMWRITE_EXR_TYPE, cv2.IMWRITE_EXR_TYPE_FLOAT])    # cv2 stores BGRA
    hdr, alpha = H.read_exr(path)
    assert np.array_equal(hdr, rgba[..., :3]) and np.array_equal(alpha, rgba[..., 3])
    assert np.array_equal(H.read_hdr(path), rgba[..., :3])
    rgb_only = str(tmp_path / "rgb.exr")
    wire.write_exr(rgb_only, rgba[..., :3])
    with pytest.raises(ValueError):
        H.read_exr(rgb_only)                                    # no alpha channel (the reference's channel('A') raises too)


def test_print_model_parm_nums(capsys):
    torch = pytest.importorskip("torch")
    handlers.print_model_parm_nums(torch.nn.Linear(1000, 1000))
    assert "Number of params: 1.00M" in capsys.readouterr().out


def test_dataset_discovers_pairs_like_the_reference(tmp_path):
    """File discovery of data.py:24-36 (a pickle without its crop is skipped); items need the GPU tonemap (gpu test)."""
    import importlib.util
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin", "data.py")
    spec = importlib.util.spec_from_file_location("_dropin_data", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    os.makedirs(tmp_path / "pkl")
    os.makedirs(tmp_path / "crop")
    for nm in ("a", "b", "c"):
        with open(tmp_path / "pkl" / (nm + ".pickle"), "wb") as f:
            pickle.dump({"distribution": np.ones(96, np.float32) / 96, "intensity": np.float32(3.0), "rgb_ratio": np.ones(3, np.float32),
                         "ambient": np.ones(3, np.float32)}, f)
    for nm in ("a", "c"):
        wire.write_exr(str(tmp_path / "crop" / (nm + ".exr")), np.ones((6, 8, 3), np.float32))
    ds = mod.ParameterDataset(str(tmp_path) + "/", device="cpu")
    assert len(ds) == 2 and [os.path.basename(p[0]) for p in ds.pairs] == ["a.exr", "c.exr"]


def test_genprojector_dataset_paths_and_mask(tmp_path):
    """GenProjector/data.py:40-57 pairing (pickle <-> warped panorama) and :75-80 light mask; items need the GPU (gpu test)."""
    import argparse
    import importlib.util
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin_genprojector", "data.py")
    spec = importlib.util.spec_from_file_location("_dropin_gp_data", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for d in ("pkl", "warped", "crop"):
        os.makedirs(tmp_path / d)
    for nm in ("x", "y"):
        with open(tmp_path / "pkl" / (nm + ".pickle"), "wb") as f:
            pickle.dump({"distribution": np.ones(128, np.float32) / 128}, f)
    wire.write_exr(str(tmp_path / "warped" / "y.exr"), np.ones((4, 8, 3), np.float32))
    ds = mod.LavalIndoorDataset(argparse.Namespace(dataroot=str(tmp_path)))
    assert len(ds) == 1 and ds.pairs[0][1].endswith("warped/y.exr")
    hdr = np.zeros((4, 8, 3), np.float32)
    hdr[1, 2] = (10, 10, 10)
    hdr[3, 3] = (0.6, 0.6, 0.6)                 # above 5 % of the maximum
    hdr[0, 0] = (0.4, 0.4, 0.4)                 # below
    m = mod.light_mask(hdr)
    assert m.shape == (1, 4, 8) and m.sum() == 2 and m[0, 1, 2] == 1 and m[0, 3, 3] == 1
    loader = mod.create_dataloader(argparse.Namespace(dataroot=str(tmp_path), batchSize=1, serial_batches=True, isTrain=False))
    assert len(loader) == 1


def test_genprojector_util_shim_names():
    import importlib.util
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin_genprojector", "util.py")
    spec = importlib.util.spec_from_file_location("_dropin_gp_util", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("TonemapHDR", "load_exr", "write_exr", "sphere_points", "convert_to_panorama", "tonemapping", "PanoramaHandler",
                 "save_test_images", "save_current_images", "print_current_errors", "save_network", "load_network"):
        assert hasattr(mod, name), name


def test_output_side_helpers_on_cpu(tmp_path, capsys):
    import argparse
    torch = pytest.importorskip("torch")
    vis = handlers.convert_visuals_to_numpy({"a": torch.arange(24.0).reshape(3, 2, 4)})
    assert vis["a"].shape == (2, 4, 3) and vis["a"][1, 2, 0] == 6.0
    handlers.print_current_errors(3, 40, {"GAN": torch.tensor([1.5, 2.5]), "VGG": torch.tensor(0.25)}, 0.5)
    assert "(epoch: 3, iters: 40, time: 0.500)GAN: 2.000 VGG: 0.250" in capsys.readouterr().out
    opt = argparse.Namespace(checkpoints_dir=str(tmp_path), name="run")
    os.makedirs(tmp_path / "run")
    net = torch.nn.Linear(3, 2)
    handlers.save_network(net, "G", "latest", opt)
    assert os.path.exists(tmp_path / "run" / "latest_net_G.pth")
    other = torch.nn.Linear(3, 2)
    handlers.load_network(other, "G", "latest", opt)
    assert torch.equal(other.weight, net.weight)
