"""Wire formats either side of the hot path: OpenEXR scan-line images and the parameter pickle.

Host-side file IO only — no device work happens here.  The reference reads / writes these through the OpenEXR + Imath Python
bindings, which this image does not have, so the container format is implemented directly from the published OpenEXR file layout
(magic / version / attribute list / offset table / scan-line chunks):

* ``load_exr(path)``   — `GenProjector/util.py:248-270`: channels R, G, B converted to FLOAT, returned as an `(H, W, 3)` float32 array.
* ``write_exr(path, data)`` — `RegressionNetwork/util.py:301-306`, `GenProjector/util.py:272-277`: what `OpenEXR.Header(W, H)` +
  `writePixels({'R','G','B'})` produce — three FLOAT channels, ZIP compression (16-line blocks), increasing-Y line order.
* ``save_parametric_lights`` / ``load_parametric_lights`` — `RegressionNetwork/test.py:79-85` writes
  `{"distribution": (N,), "rgb_ratio": (3,), "intensity": ()}` with `pickle.HIGHEST_PROTOCOL`; `GenProjector/data.py:64-66, 86-94`
  reads it back.

Decoders: NONE, RLE, ZIPS, ZIP and PIZ (what the Laval panoramas and OpenCV/OpenEXR writers use); UINT / HALF / FLOAT channels.
Tiled, deep and multi-part files and the lossy codecs (PXR24, B44, DWA) raise `ValueError`.  Pinned in `tests/test_wire_cpu.py`
against OpenCV's bundled OpenEXR library in both directions.
"""
from __future__ import annotations

import pickle
import struct
import zlib

import numpy as np

_MAGIC = 20000630
_NONE, _RLE, _ZIPS, _ZIP, _PIZ = 0, 1, 2, 3, 4
_LINES_PER_BLOCK = {_NONE: 1, _RLE: 1, _ZIPS: 1, _ZIP: 16, _PIZ: 32}
_CODEC_NAMES = {5: "PXR24", 6: "B44", 7: "B44A", 8: "DWAA", 9: "DWAB"}
_UINT, _HALF, _FLOAT = 0, 1, 2
_DTYPES = {_UINT: np.dtype("<u4"), _HALF: np.dtype("<f2"), _FLOAT: np.dtype("<f4")}


# ------------------------------------------------------------------------------------------------- header

def _cstr(buf: bytes, pos: int) -> tuple[str, int]:
    end = buf.index(b"\0", pos)
    return buf[pos:end].decode("latin-1"), end + 1


def _parse_header(buf: bytes) -> tuple[dict, int]:
    if len(buf) < 8:
        raise ValueError("not an OpenEXR file (too short)")
    magic, version = struct.unpack_from("<ii", buf, 0)
    if magic != _MAGIC:
        raise ValueError("not an OpenEXR file (bad magic number)")
    if version & 0xFF != 2:
        raise ValueError(f"unsupported OpenEXR version {version & 0xFF}")
    if version & 0x200:
        raise ValueError("tiled OpenEXR files are not supported")
    if version & 0x1800:
        raise ValueError("deep / multi-part OpenEXR files are not supported")
    pos, attrs = 8, {}
    while True:
        name, pos = _cstr(buf, pos)
        if not name:
            break
        typ, pos = _cstr(buf, pos)
        (size,) = struct.unpack_from("<i", buf, pos)
        pos += 4
        attrs[name] = (typ, buf[pos:pos + size])
        pos += size
    for need in ("channels", "compression", "dataWindow"):
        if need not in attrs:
            raise ValueError(f"OpenEXR header has no '{need}' attribute")
    channels, raw, p = [], attrs["channels"][1], 0
    while raw[p] != 0:
        cname, p = _cstr(raw, p)
        ptype, _plinear, xs, ys = struct.unpack_from("<iB3xii", raw, p)
        p += 16
        if ptype not in _DTYPES:
            raise ValueError(f"channel {cname}: unknown pixel type {ptype}")
        if xs != 1 or ys != 1:
            raise ValueError(f"channel {cname}: sub-sampled channels are not supported")
        channels.append((cname, ptype))
    x0, y0, x1, y1 = struct.unpack("<4i", attrs["dataWindow"][1])
    head = {
        "channels": channels,  # already sorted by name in the file
        "compression": attrs["compression"][1][0],
        "window": (x0, y0, x1, y1),
        "line_order": attrs["lineOrder"][1][0] if "lineOrder" in attrs else 0,
    }
    return head, pos


# ------------------------------------------------------------------------------------------------- ZIP / RLE byte transforms

def _unpredict_deinterleave(t: np.ndarray) -> np.ndarray:
    """Inverse of OpenEXR's byte predictor (`t[i] = t[i] - t[i-1] + 128`) and half/half byte split."""
    n = t.size
    d = t.astype(np.uint8).copy()
    if n > 1:
        d[1:] -= 128
    d = np.cumsum(d, dtype=np.uint8)
    out = np.empty(n, np.uint8)
    half = (n + 1) // 2
    out[0::2] = d[:half]
    out[1::2] = d[half:]
    return out


def _interleave_predict(raw: np.ndarray) -> np.ndarray:
    t = np.concatenate([raw[0::2], raw[1::2]])
    d = t.copy()
    if t.size > 1:
        d[1:] = t[1:] - t[:-1] + np.uint8(128)
    return d


def _rle_decode(src: bytes, expect: int) -> np.ndarray:
    out, i = bytearray(), 0
    while i < len(src):
        c = src[i] - 256 if src[i] > 127 else src[i]
        i += 1
        if c < 0:
            out += src[i:i - c]
            i += -c
        else:
            out += src[i:i + 1] * (c + 1)
            i += 1
    if len(out) != expect:
        raise ValueError("corrupt RLE block in OpenEXR file")
    return np.frombuffer(bytes(out), np.uint8)


# ------------------------------------------------------------------------------------------------- PIZ

_HUF_ENCBITS, _HUF_DECBITS = 16, 14
_HUF_ENCSIZE = (1 << _HUF_ENCBITS) + 1
_SHORT_ZEROCODE_RUN, _LONG_ZEROCODE_RUN = 59, 63
_SHORTEST_LONG_RUN = 2 + _LONG_ZEROCODE_RUN - _SHORT_ZEROCODE_RUN


class _BitReader:
    __slots__ = ("buf", "pos", "c", "lc")

    def __init__(self, buf: bytes, pos: int = 0):
        self.buf, self.pos, self.c, self.lc = buf, pos, 0, 0

    def get(self, n: int) -> int:
        while self.lc < n:
            self.c = ((self.c << 8) | self.buf[self.pos]) & 0xFFFFFFFFFFFFFFFF
            self.pos += 1
            self.lc += 8
        self.lc -= n
        return (self.c >> self.lc) & ((1 << n) - 1)


def _huf_unpack_lengths(buf: bytes, im: int, iM: int) -> tuple[list[int], int]:
    """Packed code-length table: 6 bits per symbol, values 59..62 = short zero runs, 63 + 8 bits = long zero runs."""
    lens = [0] * _HUF_ENCSIZE
    br = _BitReader(buf)
    i = im
    while i <= iM:
        l = br.get(6)
        if l == _LONG_ZEROCODE_RUN:
            i += br.get(8) + _SHORTEST_LONG_RUN
        elif l >= _SHORT_ZEROCODE_RUN:
            i += l - _SHORT_ZEROCODE_RUN + 2
        else:
            lens[i] = l
            i += 1
    return lens, br.pos


def _huf_canonical_codes(lens: list[int]) -> list[int]:
    """Canonical codes as OpenEXR assigns them: shorter codes have numerically larger prefixes."""
    n = [0] * 59
    for l in lens:
        n[l] += 1
    c = 0
    for i in range(58, 0, -1):
        nc = (c + n[i]) >> 1
        n[i] = c
        c = nc
    codes = [0] * len(lens)
    for i, l in enumerate(lens):
        if l > 0:
            codes[i] = n[l]
            n[l] += 1
    return codes


def _huf_decode(buf: bytes, n_raw: int) -> np.ndarray:
    im, iM, tbl_len, n_bits, _ = struct.unpack_from("<5i", buf, 0)
    if not (0 <= im < _HUF_ENCSIZE and 0 <= iM < _HUF_ENCSIZE):
        raise ValueError("corrupt PIZ block in OpenEXR file")
    lens, _ = _huf_unpack_lengths(buf[20:20 + tbl_len] + b"\0\0", im, iM)
    codes = _huf_canonical_codes(lens)
    rlc = iM  # the run-length symbol is the largest one
    # decode by (length, code) lookup, shortest codes first
    table = {}
    for sym in range(im, iM + 1):
        if lens[sym]:
            table[(lens[sym], codes[sym])] = sym
    min_len = min((l for l in lens if l), default=0)
    out = np.empty(n_raw, np.uint16)
    data = buf[20 + tbl_len:]
    total, pos_bits, o = n_bits, 0, 0
    c, lc, p = 0, 0, 0

    def need(n):
        nonlocal c, lc, p
        while lc < n:
            c = (c << 8) | (data[p] if p < len(data) else 0)
            p += 1
            lc += 8

    while pos_bits < total and o < n_raw:
        l, code = 0, 0
        while True:
            step = min_len if l == 0 else 1
            need(step)
            lc -= step
            code = (code << step) | ((c >> lc) & ((1 << step) - 1))
            c &= (1 << lc) - 1
            l += step
            sym = table.get((l, code))
            if sym is not None:
                break
            if l > 58:
                raise ValueError("corrupt PIZ block in OpenEXR file")
        pos_bits += l
        if sym == rlc:
            need(8)
            lc -= 8
            run = (c >> lc) & 0xFF
            c &= (1 << lc) - 1
            pos_bits += 8
            if o == 0 or o + run > n_raw:
                raise ValueError("corrupt PIZ block in OpenEXR file")
            out[o:o + run] = out[o - 1]
            o += run
        else:
            out[o] = sym
            o += 1
    if o != n_raw:
        raise ValueError("corrupt PIZ block in OpenEXR file")
    return out


def _wdec14(l, h):
    """Inverse 14-bit wavelet step (signed 16-bit arithmetic), vectorised."""
    ls = l.astype(np.int16).astype(np.int32)
    hs = h.astype(np.int16).astype(np.int32)
    ai = ls + (hs & 1) + (hs >> 1)
    return (ai & 0xFFFF).astype(np.uint16), ((ai - hs) & 0xFFFF).astype(np.uint16)


def _wdec16(l, h):
    """Inverse 16-bit (modulo) wavelet step, vectorised."""
    m = l.astype(np.int32)
    d = h.astype(np.int32)
    bb = (m - (d >> 1)) & 0xFFFF
    aa = (d + bb - 0x8000) & 0xFFFF
    return aa.astype(np.uint16), bb.astype(np.uint16)


def _wav2_decode(a: np.ndarray, nx: int, ny: int, mx: int) -> None:
    """In-place inverse 2-D Haar-like wavelet of OpenEXR's PIZ codec on an (ny, nx) uint16 plane."""
    dec = _wdec14 if mx < (1 << 14) else _wdec16
    n = min(nx, ny)
    p = 1
    while p <= n:
        p <<= 1
    p >>= 1
    p2 = p
    p >>= 1
    while p >= 1:
        ys = np.arange(0, ny - p2 + 1, p2) if ny - p2 >= 0 else np.arange(0)
        xs = np.arange(0, nx - p2 + 1, p2) if nx - p2 >= 0 else np.arange(0)
        if ys.size and xs.size:
            Y, X = np.meshgrid(ys, xs, indexing="ij")
            i00, i10 = a[Y, X], a[Y + p, X]
            i01, i11 = a[Y, X + p], a[Y + p, X + p]
            a00, a10 = dec(i00, i10)
            a01, a11 = dec(i01, i11)
            o00, o01 = dec(a00, a01)
            o10, o11 = dec(a10, a11)
            a[Y, X], a[Y, X + p], a[Y + p, X], a[Y + p, X + p] = o00, o01, o10, o11
        if nx & p:  # odd column at this level: 1-D step down the rows
            x = xs[-1] + p2 if xs.size else 0
            if ys.size:
                o0, o1 = dec(a[ys, x], a[ys + p, x])
                a[ys, x], a[ys + p, x] = o0, o1
        if ny & p:  # odd row at this level: 1-D step along the columns
            y = ys[-1] + p2 if ys.size else 0
            if xs.size:
                o0, o1 = dec(a[y, xs], a[y, xs + p])
                a[y, xs], a[y, xs + p] = o0, o1
        p2 = p
        p >>= 1


def _piz_decode(src: bytes, chan_bytes: list[int], width: int, lines: int) -> np.ndarray:
    """One PIZ chunk -> the uncompressed chunk bytes (line-interleaved channel rows)."""
    min_nz, max_nz = struct.unpack_from("<HH", src, 0)
    pos = 4
    bitmap = np.zeros(8192, np.uint8)
    if min_nz <= max_nz:
        cnt = max_nz - min_nz + 1
        bitmap[min_nz:min_nz + cnt] = np.frombuffer(src, np.uint8, cnt, pos)
        pos += cnt
    present = np.unpackbits(bitmap, bitorder="little").astype(bool)
    present[0] = True  # zero is always mapped
    lut = np.zeros(65536, np.uint16)
    vals = np.nonzero(present)[0].astype(np.uint16)
    lut[:vals.size] = vals
    max_value = vals.size - 1
    (length,) = struct.unpack_from("<i", src, pos)
    pos += 4
    halves = [b // 2 for b in chan_bytes]  # uint16 words per pixel per channel
    n_raw = sum(h * width * lines for h in halves)
    raw = _huf_decode(src[pos:pos + length], n_raw)
    # layout of the wavelet domain: per channel a block of lines x (width * words), words interleaved per pixel
    planes, o = [], 0
    for h in halves:
        blk = raw[o:o + h * width * lines].reshape(lines, width, h).copy()
        for j in range(h):
            plane = np.ascontiguousarray(blk[:, :, j])
            _wav2_decode(plane, width, lines, max_value)
            blk[:, :, j] = plane
        planes.append(lut[blk].reshape(lines, width * h))
        o += h * width * lines
    rows = [np.concatenate([pl[y] for pl in planes]) for y in range(lines)]
    return np.concatenate(rows).astype("<u2").view(np.uint8)


# ------------------------------------------------------------------------------------------------- public: EXR

def read_exr_channels(path: str) -> dict[str, np.ndarray]:
    """All channels of a scan-line OpenEXR file as `(H, W)` arrays in their stored type (uint32 / float16 / float32)."""
    with open(path, "rb") as f:
        buf = f.read()
    head, pos = _parse_header(buf)
    comp = head["compression"]
    if comp not in _LINES_PER_BLOCK:
        raise ValueError(f"OpenEXR compression {_CODEC_NAMES.get(comp, comp)} is not supported")
    x0, y0, x1, y1 = head["window"]
    width, height = x1 - x0 + 1, y1 - y0 + 1
    if width <= 0 or height <= 0:
        raise ValueError("OpenEXR file has an empty data window")
    lpb = _LINES_PER_BLOCK[comp]
    n_blocks = (height + lpb - 1) // lpb
    offsets = struct.unpack_from(f"<{n_blocks}Q", buf, pos)
    chans = head["channels"]
    cbytes = [_DTYPES[t].itemsize for _, t in chans]
    line_bytes = width * sum(cbytes)
    out = {name: np.empty((height, width), _DTYPES[t]) for name, t in chans}
    for off in offsets:
        y, size = struct.unpack_from("<ii", buf, off)
        data = buf[off + 8:off + 8 + size]
        r0 = y - y0
        if not 0 <= r0 < height or len(data) != size:
            raise ValueError("corrupt OpenEXR chunk table")
        lines = min(lpb, height - r0)
        expect = line_bytes * lines
        if size == expect and comp != _NONE:
            raw = np.frombuffer(data, np.uint8)  # codec stored the block raw because it did not shrink
        elif comp == _NONE:
            raw = np.frombuffer(data, np.uint8)
        elif comp in (_ZIP, _ZIPS):
            raw = _unpredict_deinterleave(np.frombuffer(zlib.decompress(data), np.uint8))
        elif comp == _RLE:
            raw = _unpredict_deinterleave(_rle_decode(data, expect))
        else:
            raw = _piz_decode(data, cbytes, width, lines)
        if raw.size != expect:
            raise ValueError("corrupt OpenEXR chunk (wrong decoded size)")
        rows = raw.reshape(lines, line_bytes)
        o = 0
        for (name, t), cb in zip(chans, cbytes):
            out[name][r0:r0 + lines] = rows[:, o:o + width * cb].copy().view(_DTYPES[t])
            o += width * cb
    return out


def load_exr(in_file: str) -> np.ndarray:
    """`util.load_exr` (`GenProjector/util.py:248-270`): R, G, B as FLOAT -> `(H, W, 3)` float32."""
    ch = read_exr_channels(in_file)
    missing = [c for c in "RGB" if c not in ch]
    if missing:
        raise ValueError(f"OpenEXR file has no channel(s) {missing}")
    return np.stack([ch[c].astype(np.float32) for c in "RGB"], axis=-1)


def write_exr(out_file: str, data) -> None:
    """`util.write_exr` (`RegressionNetwork/util.py:301-306`): `(H, W, 3)` -> FLOAT R/G/B, ZIP, increasing Y."""
    data = np.asarray(data)
    if data.ndim != 3 or data.shape[2] < 3:
        raise ValueError(f"write_exr expects an (H, W, 3) array, got {data.shape}")
    height, width = data.shape[:2]
    if height == 0 or width == 0:
        raise ValueError("write_exr: empty image")
    planes = [np.ascontiguousarray(data[:, :, c], dtype="<f4") for c in (2, 1, 0)]  # stored alphabetically: B, G, R

    def attr(name: str, typ: str, payload: bytes) -> bytes:
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    chlist = b"".join(n + b"\0" + struct.pack("<iB3xii", _FLOAT, 0, 1, 1) for n in (b"B", b"G", b"R")) + b"\0"
    window = struct.pack("<4i", 0, 0, width - 1, height - 1)
    header = struct.pack("<ii", _MAGIC, 2) + b"".join([
        attr("channels", "chlist", chlist),
        attr("compression", "compression", bytes([_ZIP])),
        attr("dataWindow", "box2i", window),
        attr("displayWindow", "box2i", window),
        attr("lineOrder", "lineOrder", b"\0"),
        attr("pixelAspectRatio", "float", struct.pack("<f", 1.0)),
        attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0)),
        attr("screenWindowWidth", "float", struct.pack("<f", 1.0)),
    ]) + b"\0"
    lpb = _LINES_PER_BLOCK[_ZIP]
    chunks = []
    for r0 in range(0, height, lpb):
        r1 = min(r0 + lpb, height)
        raw = np.concatenate([pl[y].view(np.uint8) for y in range(r0, r1) for pl in planes])
        packed = zlib.compress(_interleave_predict(raw).tobytes())
        body = packed if len(packed) < raw.size else raw.tobytes()
        chunks.append(struct.pack("<ii", r0, len(body)) + body)
    table_at = len(header)
    offs, pos = [], table_at + 8 * len(chunks)
    for c in chunks:
        offs.append(pos)
        pos += len(c)
    with open(out_file, "wb") as f:
        f.write(header + struct.pack(f"<{len(offs)}Q", *offs) + b"".join(chunks))


# ------------------------------------------------------------------------------------------------- public: parameter pickle

def save_parametric_lights(path: str, distribution, rgb_ratio, intensity) -> None:
    """`RegressionNetwork/test.py:79-85`: sample 0's heads as numpy arrays in one dict, `pickle.HIGHEST_PROTOCOL`."""
    def to_np(v):
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        return np.squeeze(np.asarray(v))
    rec = {"distribution": to_np(distribution), "rgb_ratio": to_np(rgb_ratio), "intensity": to_np(intensity)}
    if rec["distribution"].ndim != 1 or rec["rgb_ratio"].shape != (3,) or rec["intensity"].ndim != 0:
        raise ValueError("expected distribution (N,), rgb_ratio (3,), intensity scalar for ONE sample")
    with open(path, "wb") as handle:
        pickle.dump(rec, handle, protocol=pickle.HIGHEST_PROTOCOL)


def load_parametric_lights(path: str) -> dict:
    """`GenProjector/data.py:64-66`: the dict written above (`distribution`, `rgb_ratio`, `intensity`)."""
    with open(path, "rb") as handle:
        rec = pickle.load(handle)
    for key in ("distribution", "rgb_ratio", "intensity"):
        if key not in rec:
            raise KeyError(f"parameter pickle has no '{key}'")
    return rec
