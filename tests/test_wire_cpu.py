"""Wire formats: the OpenEXR reader / writer against an independent implementation of the format — OpenCV's
bundled OpenEXR library — through committed fixtures (`oracle/make_golden_exr.py`) and, when cv2 is importable, live in both
directions; the parameter pickle of `RegressionNetwork/test.py:79-85` as `GenProjector/data.py:64-94` reads it.  Bit-exact."""
import os
import pickle

import numpy as np
import pytest

from emlight_b200 import wire

EXR_DIR = os.path.join(os.path.dirname(__file__), "golden", "exr")
CODECS = ["none", "rle", "zips", "zip", "piz"]


def _cv2():
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    return pytest.importorskip("cv2")


@pytest.mark.parametrize("ptype", ["float", "half"])
@pytest.mark.parametrize("codec", CODECS)
def test_load_exr_matches_openexr_fixture(codec, ptype):
    expected = np.load(os.path.join(EXR_DIR, "expected.npz"))[f"{codec}_{ptype}"]
    got = wire.load_exr(os.path.join(EXR_DIR, f"{codec}_{ptype}.exr"))
    assert got.dtype == np.float32 and got.shape == expected.shape == (37, 53, 3)
    assert np.array_equal(got, expected)


def test_read_exr_channels_keeps_stored_type():
    ch = wire.read_exr_channels(os.path.join(EXR_DIR, "piz_half.exr"))
    assert sorted(ch) == ["B", "G", "R"] and all(v.dtype == np.float16 for v in ch.values())


@pytest.mark.parametrize("shape", [(1, 1), (3, 100), (16, 32), (17, 5), (128, 256)])
def test_write_exr_round_trip_and_layout(tmp_path, shape):
    rng = np.random.default_rng(sum(shape))
    img = rng.lognormal(0.0, 1.5, shape + (3,)).astype(np.float32)
    if shape[0] > 2:
        img[1] = 0.0  # a compressible line next to incompressible ones
    path = str(tmp_path / "out.exr")
    wire.write_exr(path, img)
    assert np.array_equal(wire.load_exr(path), img)
    raw = open(path, "rb").read()
    # what OpenEXR.Header(W, H) + writePixels produce (util.py:301-306): FLOAT B, G, R; ZIP; increasing Y
    assert raw[:8] == bytes([0x76, 0x2F, 0x31, 0x01, 2, 0, 0, 0])
    i = raw.index(b"compression\0compression\0")
    assert raw[i + 24:i + 29] == b"\x01\0\0\0\x03"
    ch = wire.read_exr_channels(path)
    assert list(ch) == ["B", "G", "R"] and all(v.dtype == np.float32 for v in ch.values())


def test_write_exr_accepts_float64_and_rejects_bad_shapes(tmp_path):
    img = np.linspace(0, 5, 4 * 6 * 3).reshape(4, 6, 3)
    wire.write_exr(str(tmp_path / "d.exr"), img)
    assert np.array_equal(wire.load_exr(str(tmp_path / "d.exr")), img.astype(np.float32))
    with pytest.raises(ValueError):
        wire.write_exr(str(tmp_path / "bad.exr"), np.zeros((4, 6)))
    with pytest.raises(ValueError):
        wire.write_exr(str(tmp_path / "bad.exr"), np.zeros((0, 6, 3)))


def test_load_exr_rejects_garbage_and_unsupported(tmp_path):
    p = tmp_path / "x.exr"
    p.write_bytes(b"not an exr file at all")
    with pytest.raises(ValueError, match="magic"):
        wire.load_exr(str(p))
    good = open(os.path.join(EXR_DIR, "zip_float.exr"), "rb").read()
    tiled = bytearray(good)
    tiled[5] |= 0x02  # version flag 0x200: tiled
    p.write_bytes(bytes(tiled))
    with pytest.raises(ValueError, match="tiled"):
        wire.load_exr(str(p))
    lossy = bytearray(good)
    i = good.index(b"compression\0compression\0")
    lossy[i + 28] = 8  # DWAA
    p.write_bytes(bytes(lossy))
    with pytest.raises(ValueError, match="DWAA"):
        wire.load_exr(str(p))
    trunc = good[:len(good) - 40]
    p.write_bytes(trunc)
    with pytest.raises(Exception):
        wire.load_exr(str(p))


def test_openexr_library_reads_what_we_write(tmp_path):
    cv2 = _cv2()
    rng = np.random.default_rng(3)
    for shape in [(128, 256), (33, 7)]:
        img = rng.lognormal(0.0, 1.0, shape + (3,)).astype(np.float32)
        img[::3] = np.round(img[::3])
        path = str(tmp_path / "ours.exr")
        wire.write_exr(path, img)
        back = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        assert back is not None and np.array_equal(back[:, :, ::-1], img)


@pytest.mark.parametrize("codec", CODECS)
def test_we_read_what_the_openexr_library_writes(tmp_path, codec):
    cv2 = _cv2()
    flag = {"none": cv2.IMWRITE_EXR_COMPRESSION_NO, "rle": cv2.IMWRITE_EXR_COMPRESSION_RLE, "zips": cv2.IMWRITE_EXR_COMPRESSION_ZIPS,
            "zip": cv2.IMWRITE_EXR_COMPRESSION_ZIP, "piz": cv2.IMWRITE_EXR_COMPRESSION_PIZ}[codec]
    y, x = np.mgrid[0:128, 0:256]
    pano = np.stack([np.exp(4 * np.cos(x / 40.0) * np.sin(y / 20.0)), (x + 2 * y) / 64.0, np.full(x.shape, 0.25)], -1).astype(np.float32)
    for ptype in (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_TYPE_HALF):
        path = str(tmp_path / "theirs.exr")
        assert cv2.imwrite(path, pano[:, :, ::-1], [cv2.IMWRITE_EXR_TYPE, ptype, cv2.IMWRITE_EXR_COMPRESSION, flag])
        expected = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1]
        assert np.array_equal(wire.load_exr(path), expected)


def test_parametric_lights_pickle_contract(tmp_path):
    torch = pytest.importorskip("torch")
    N = 128
    dist = torch.softmax(torch.randn(2, N), 1)
    rgb = torch.rand(2, 3)
    inten = torch.rand(2, 1)
    path = str(tmp_path / "im.pickle")
    # test.py:79-82 passes sample 0's heads
    wire.save_parametric_lights(path, dist[0].view(N), rgb[0].view(3), inten[0])
    with open(path, "rb") as handle:  # GenProjector/data.py:64-66 reads it with plain pickle.load
        pkl = pickle.load(handle)
    assert set(pkl) == {"distribution", "rgb_ratio", "intensity"}
    assert pkl["distribution"].shape == (N,) and pkl["rgb_ratio"].shape == (3,) and pkl["intensity"].shape == ()
    assert all(isinstance(v, np.ndarray) and v.dtype == np.float32 for v in pkl.values())
    rec = wire.load_parametric_lights(path)
    assert np.array_equal(rec["distribution"], dist[0].numpy()) and float(rec["intensity"]) == float(inten[0])
    with pytest.raises(ValueError):
        wire.save_parametric_lights(path, dist, rgb[0], inten[0])  # a whole batch is not one record
    with open(path, "wb") as handle:
        pickle.dump({"distribution": 1}, handle)
    with pytest.raises(KeyError):
        wire.load_parametric_lights(path)


def test_dropin_util_exposes_exr_io():
    import importlib.util
    here = os.path.join(os.path.dirname(__file__), "..", "emlight_b200", "dropin", "util.py")
    spec = importlib.util.spec_from_file_location("_dropin_util_wire", here)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.load_exr is wire.load_exr and mod.write_exr is wire.write_exr


def _raw_exr(path, planes, x0=0, y0=0, line_order=0, ptype=2, shuffle=None):
    """Hand-built uncompressed scan-line file (published layout): channels in `planes` (name -> (H,W) array), data window at (x0,y0),
    chunks stored in `shuffle` order -- exercises what a writer other than ours may legally produce."""
    import struct
    names = sorted(planes)
    h, w = planes[names[0]].shape
    dt = {0: "<u4", 1: "<f2", 2: "<f4"}[ptype]

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iB3xii", ptype, 0, 1, 1) for n in names) + b"\0"
    win = struct.pack("<4i", x0, y0, x0 + w - 1, y0 + h - 1)
    header = struct.pack("<ii", 20000630, 2) + attr("channels", "chlist", chlist) + attr("compression", "compression", b"\0") + \
        attr("dataWindow", "box2i", win) + attr("displayWindow", "box2i", win) + attr("lineOrder", "lineOrder", bytes([line_order])) + \
        attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)) + attr("screenWindowCenter", "v2f", struct.pack("<ff", 0, 0)) + \
        attr("screenWindowWidth", "float", struct.pack("<f", 1.0)) + b"\0"
    rows = list(range(h))
    stored = list(shuffle) if shuffle is not None else (rows[::-1] if line_order == 1 else rows)
    chunks = {}
    for r in stored:
        body = b"".join(np.ascontiguousarray(planes[n][r]).astype(dt).tobytes() for n in names)
        chunks[r] = struct.pack("<ii", y0 + r, len(body)) + body
    pos = len(header) + 8 * h
    offs = {}
    blob = b""
    for r in stored:
        offs[r] = pos + len(blob)
        blob += chunks[r]
    table_rows = rows[::-1] if line_order == 1 else rows                  # the offset table follows the line order
    with open(path, "wb") as f:
        f.write(header + struct.pack("<%dQ" % h, *[offs[r] for r in table_rows]) + blob)


def test_reader_handles_window_offset_line_order_and_uint(tmp_path):
    rng = np.random.default_rng(9)
    planes = {c: rng.random((7, 5)).astype(np.float32) * 3 for c in "RGBA"}
    want = np.stack([planes[c] for c in "RGB"], -1)
    p = str(tmp_path / "w.exr")
    _raw_exr(p, planes, x0=-3, y0=11)                                      # data window not at the origin
    assert np.array_equal(wire.load_exr(p), want)
    _raw_exr(p, planes, line_order=1)                                      # DECREASING_Y
    assert np.array_equal(wire.load_exr(p), want)
    _raw_exr(p, planes, shuffle=[3, 0, 6, 1, 5, 2, 4])                     # RANDOM_Y-style chunk order: rows are placed by their y
    assert np.array_equal(wire.load_exr(p), want)
    ids = {c: rng.integers(0, 2 ** 31, (4, 6)).astype(np.uint32) for c in "RGB"}
    _raw_exr(p, ids, ptype=0)                                              # UINT channels convert to float like OpenEXR's FLOAT request
    assert np.array_equal(wire.load_exr(p), np.stack([ids[c].astype(np.float32) for c in "RGB"], -1))
    ch = wire.read_exr_channels(p)
    assert ch["R"].dtype == np.uint32 and np.array_equal(ch["G"], ids["G"])
    half = {c: rng.random((3, 4)).astype(np.float16) for c in "BGR"}
    _raw_exr(p, half, ptype=1)
    assert np.array_equal(wire.load_exr(p), np.stack([half[c].astype(np.float32) for c in "RGB"], -1))


def test_reader_fuzz_against_the_openexr_library(tmp_path):
    """Random sizes / contents (smooth, noise, sparse highlights, constant, quantised, HDR with saturated lights) x {PIZ, ZIP, RLE} x
    {FLOAT, HALF}: every file OpenCV's OpenEXR writes reads back bit-identically (a 1440-file run of the same loop found no mismatch)."""
    cv2 = _cv2()
    rng = np.random.default_rng(123)
    kinds = ["smooth", "noise", "sparse", "const", "steps", "hdr"]
    path = str(tmp_path / "fz.exr")
    for it in range(18):
        h, w = int(rng.integers(1, 140)), int(rng.integers(1, 300))
        k = kinds[it % len(kinds)]
        y, x = np.mgrid[0:h, 0:w]
        if k == "smooth":
            a = np.stack([np.sin(x / rng.uniform(3, 30)) + 1.2, np.cos(y / rng.uniform(3, 30)) + 1.5, (x * y) / (h * w + 1.0)], -1)
        elif k == "noise":
            a = rng.lognormal(0, 2, (h, w, 3))
        elif k == "sparse":
            a = np.zeros((h, w, 3))
            m = rng.random((h, w)) < 0.02
            a[m] = rng.uniform(0, 5000, (int(m.sum()), 3))
        elif k == "const":
            a = np.full((h, w, 3), rng.uniform(0, 10))
        elif k == "steps":
            a = np.floor(rng.random((h, w, 3)) * rng.integers(2, 2000)) / 7.0
        else:
            a = np.exp(rng.normal(-2, 1, (h, w, 3)))
            a[h // 3:h // 3 + 2, w // 2:w // 2 + 3] = rng.uniform(1e3, 6e4)
        a = a.astype(np.float32)
        for comp in (cv2.IMWRITE_EXR_COMPRESSION_PIZ, cv2.IMWRITE_EXR_COMPRESSION_ZIP, cv2.IMWRITE_EXR_COMPRESSION_RLE):
            for typ in (cv2.IMWRITE_EXR_TYPE_FLOAT, cv2.IMWRITE_EXR_TYPE_HALF):
                assert cv2.imwrite(path, a[:, :, ::-1], [cv2.IMWRITE_EXR_TYPE, typ, cv2.IMWRITE_EXR_COMPRESSION, comp])
                ref = cv2.imread(path, cv2.IMREAD_UNCHANGED)[:, :, ::-1]
                assert np.array_equal(wire.load_exr(path), ref), (it, k, h, w, comp, typ)
