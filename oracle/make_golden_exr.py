"""TEST INFRASTRUCTURE.  Writes the OpenEXR fixtures of tests/golden/exr/ with an independent implementation of the file format:
OpenCV's bundled OpenEXR library (`cv2.imwrite` with OPENCV_IO_ENABLE_OPENEXR=1) — the reference's own bindings (OpenEXR + Imath,
`GenProjector/util.py:248-277`) are not installed in this image.  One small image per codec / pixel type the reader claims, plus the
pixels OpenCV itself reads back (`expected.npz`).  Run from the repo root:  python oracle/make_golden_exr.py
"""
import os

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "exr")

CODECS = {
    "none": cv2.IMWRITE_EXR_COMPRESSION_NO, "rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS,
    "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP, "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ,
}
TYPES = {"float": cv2.IMWRITE_EXR_TYPE_FLOAT, "half": cv2.IMWRITE_EXR_TYPE_HALF}


def image(h=37, w=53):
    """Smooth HDR-like content with a few saturated pixels and a zero band: compressible, so every codec really runs."""
    rng = np.random.default_rng(7)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([np.sin(x / 9.0) + 1.1, np.cos(y / 7.0) * 3 + 3.5, (x + y) / 50.0], -1).astype(np.float32)
    img = np.round(img * 8) / 8
    img[rng.integers(0, h, 6), rng.integers(0, w, 6)] = 3000.0
    img[5:8] = 0.0
    return img


def main():
    os.makedirs(OUT, exist_ok=True)
    img = image()
    expected = {}
    for cname, c in CODECS.items():
        for tname, t in TYPES.items():
            path = os.path.join(OUT, f"{cname}_{tname}.exr")
            assert cv2.imwrite(path, img[:, :, ::-1], [cv2.IMWRITE_EXR_TYPE, t, cv2.IMWRITE_EXR_COMPRESSION, c])
            expected[f"{cname}_{tname}"] = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1].copy()
    np.savez_compressed(os.path.join(OUT, "expected.npz"), **expected)
    print("wrote", len(expected), "fixtures,", sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)), "bytes")


if __name__ == "__main__":
    main()
