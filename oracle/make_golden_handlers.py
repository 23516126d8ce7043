"""TEST INFRASTRUCTURE.  Golden vectors for the host-side helpers of RegressionNetwork/util.py (PanoramaHandler :69-186,
cartesian_to_polar / polar_to_cartesian :206-220): the file itself is not importable (merge-conflict markers, OpenEXR), so the class
and the two functions are exec'd from their line ranges, run on a seeded synthetic panorama, and the results written to
tests/golden/handlers.npz.   python oracle/make_golden_handlers.py"""
import os

import cv2
import numpy as np
from scipy import interpolate

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def synthetic_pano(seed=3, h=64, w=128):
    rng = np.random.default_rng(seed)
    img = np.exp(rng.normal(-2.0, 1.0, (h, w, 3))).astype(np.float32)
    img[10:14, 30:36] += 200.0                                  # a light source: most pixels fall under max/20
    return img


def main():
    src = open("/root/reference/RegressionNetwork/util.py").read().split("\n")
    ns = {"np": np, "cv2": cv2, "interpolate": interpolate}
    exec("\n".join(src[68:186]), ns)                            # class PanoramaHandler
    exec("\n".join(src[205:221]), ns)                           # cartesian_to_polar, polar_to_cartesian
    H = ns["PanoramaHandler"]
    pano = synthetic_pano()
    out = {"intensity": H.rgb_to_intenisty(pano), "rot": H.horizontal_rotate_panorama(pano, 77.0),
           "ster": H.generate_steradian(64, 128), "ster_raw": H.generate_steradian(16, 32, multiply=False)}
    gt, amb = H.prepare_gt_panorama(pano.copy())
    out["gt"], out["ambient"] = gt, amb
    gt2, amb2 = H.prepare_gt_panorama(pano.copy(), threshold=1e-9)          # nothing below the threshold
    out["gt2"], out["ambient2"] = gt2, amb2
    out["resized"] = H.resize_panorama(pano, 16)
    out["resized_t"] = H.resize_panorama(pano, (40, 24))
    out["crop"] = H.crop_panorama(pano, 60.0, crop_image_h=30).astype(np.float32)
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(3, 50))
    xyz /= np.linalg.norm(xyz, axis=0)
    phi, theta = ns["cartesian_to_polar"](xyz)
    out["xyz"], out["phi"], out["theta"] = xyz, phi, theta
    out["back"] = ns["polar_to_cartesian"]((phi, theta))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "handlers.npz"), **out)
    print("wrote handlers.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
