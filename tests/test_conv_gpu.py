"""GPU: tcgen05 implicit-GEMM convolution (bf16 / bf16x3) and its fp32 SIMT companion vs a torch CPU reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

MODES = {"1x1": 0, "3x3": 1, "pool": 2}
TOL = {"fp32": 2e-5, "bf16x3": 1e-4, "bf16": 2e-2}     # relative to the output's max magnitude


def conv_ref(x, c_in, scale, shift, w, mode, relu):
    """x (B,H,W,pitch) fp32 NHWC -> (B,Ho,Wo,C_out) in float64."""
    a = x[..., :c_in].double() * scale.double() + shift.double()
    if relu:
        a = a.clamp_min(0)
    a = a.permute(0, 3, 1, 2)
    if mode == "pool":
        a = F.avg_pool2d(a, 2)
    out = F.conv2d(a, w.double(), padding=1 if mode == "3x3" else 0)
    return out.permute(0, 2, 3, 1).contiguous()


def run_conv(lib, cuda, x, c_in, scale, shift, w, mode, relu, precision, out_pitch, choff, want_stats):
    from emlight_b200 import _lib
    from emlight_b200._lib import ConvParams
    B, H, W, pitch = x.shape
    c_out = w.shape[0]
    taps = 9 if mode == "3x3" else 1
    Ho, Wo = (H // 2, W // 2) if mode == "pool" else (H, W)
    xd, wd = x.to(cuda), w.contiguous().to(cuda)
    pad = (c_in + 3) & ~3
    sc = torch.zeros(pad, device=cuda); sc[:c_in] = scale.to(cuda)
    sh = torch.zeros(pad, device=cuda); sh[:c_in] = shift.to(cuda)
    out = torch.full((B, Ho, Wo, out_pitch), float("nan"), device=cuda)
    stats = torch.zeros(2 * out_pitch, dtype=torch.float64, device=cuda) if want_stats else None
    wp = torch.empty(lib.eml_conv_wpack_bytes(c_out, c_in, taps), dtype=torch.uint8, device=cuda)
    _lib.check(lib.eml_conv_pack_weights(_lib.ptr(wd), _lib.ptr(wp), c_out, c_in, taps, _lib.stream_ptr()))
    p = ConvParams()
    p.in_, p.scale, p.shift, p.w_oihw, p.wpack, p.out = (t.data_ptr() for t in (xd, sc, sh, wd, wp, out))
    p.stats = stats.data_ptr() + 8 * choff if want_stats else None
    p.stats_stride = out_pitch
    p.B, p.H, p.W, p.C_in, p.in_pitch = B, H, W, c_in, pitch
    p.C_out, p.out_pitch, p.out_choff = c_out, out_pitch, choff
    p.mode, p.relu, p.precision = MODES[mode], int(relu), _lib.PRECISIONS[precision]
    _lib.check(lib.eml_conv_forward(p, _lib.stream_ptr()), "eml_conv_forward")
    torch.cuda.synchronize()
    return out.cpu(), (stats.cpu() if want_stats else None)


CASES = [
    # mode, B, H, W, c_in, pitch, c_out, out_pitch, choff, relu
    ("1x1", 1, 8, 16, 24, 216, 48, 48, 0, True),          # denseblock1.denselayer1.conv1, one exact tile
    ("1x1", 2, 6, 10, 150, 344, 48, 48, 0, True),         # block-3 channel count (not a multiple of 4/64), ragged M
    ("1x1", 1, 16, 16, 330, 344, 48, 48, 0, True),        # 6 K-chunks: stage ring wraps 3 times
    ("3x3", 1, 8, 16, 48, 48, 12, 216, 24, False),        # conv2 writing 12 channels in place at offset 24
    ("3x3", 2, 5, 7, 48, 48, 12, 344, 150, False),        # ragged M, image borders dominate
    ("pool", 1, 16, 16, 216, 216, 108, 300, 0, True),     # transition1
    ("pool", 2, 4, 8, 342, 344, 171, 172, 0, True),       # transition3: N_pad 176, odd C_out (scalar stores)
    ("1x1", 3, 24, 32, 204, 216, 48, 48, 0, True),        # 18 tiles
    ("3x3", 1, 4, 128, 48, 48, 12, 216, 24, False),       # planar no-im2col kernel (W % 128 == 0): block-2 geometry
    ("3x3", 2, 3, 256, 48, 48, 12, 344, 150, False),      # block-1 geometry, two tiles per row, top/bottom borders
    ("3x3", 1, 1, 128, 48, 48, 12, 12, 0, True),          # single row: both vertical neighbours are padding
]


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-B%d-%dx%d-c%d-o%d" % (c[0], c[1], c[2], c[3], c[4], c[6]))
def test_conv_matches_reference(lib, cuda, case, precision):
    mode, B, H, W, c_in, pitch, c_out, out_pitch, choff, relu = case
    g = torch.Generator().manual_seed(hash(case) & 0xffff)
    x = torch.randn(B, H, W, pitch, generator=g)
    x[..., c_in:] = float("nan")                           # channels beyond C_in must never be consumed
    scale = 0.5 + torch.rand(c_in, generator=g)
    shift = 0.3 * torch.randn(c_in, generator=g)
    taps = 3 if mode == "3x3" else 1
    w = torch.randn(c_out, c_in, taps, taps, generator=g) / np.sqrt(c_in * taps * taps)
    out, stats = run_conv(lib, cuda, x, c_in, scale, shift, w, mode, relu, precision, out_pitch, choff, True)
    ref = conv_ref(x, c_in, scale, shift, w, mode, relu)
    got = out[..., choff:choff + c_out].double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= TOL[precision], err
    # untouched channels of the destination stay untouched (in-place concat contract)
    rest = torch.cat([out[..., :choff], out[..., choff + c_out:]], -1)
    assert torch.isnan(rest).all()
    # batch statistics of the written values
    s1 = stats[choff:choff + c_out]; s2 = stats[out_pitch + choff:out_pitch + choff + c_out]
    r1 = got.reshape(-1, c_out).sum(0); r2 = (got.reshape(-1, c_out) ** 2).sum(0)
    assert (s1 - r1).abs().max() <= 1e-5 * r1.abs().max() + 1e-6
    assert (s2 - r2).abs().max() <= 1e-5 * r2.abs().max() + 1e-6
    assert stats[:choff].abs().max() == 0 if choff else True


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("case", [c for c in CASES if c[0] != "pool"] + [("1x1", 40, 48, 64, 132, 216, 48, 48, 0, True),
                                  # transition1 geometry without statistics -> the TMA-fed pipeline of dense_layer.cu with a pooling
                                  # epilogue (row pairs accumulate in TMEM, column pairs by shuffle): two tiles per row / one tile per row
                                  ("pool", 2, 8, 256, 216, 216, 108, 304, 0, True), ("pool", 3, 4, 128, 204, 216, 100, 112, 4, True),
                                  ("pool", 37, 12, 128, 216, 216, 108, 108, 0, True)],
                         ids=lambda c: "%s-B%d-%dx%d-c%d" % (c[0], c[1], c[2], c[3], c[4]))
def test_conv_persistent_kernels(lib, cuda, case, precision):
    """Without a statistics epilogue the dispatcher picks the persistent warp-specialised kernels (one CTA per SM walking
    the tiles); the last case has 960 tiles so every CTA loops > 6 times and the stage/accumulator rings wrap."""
    mode, B, H, W, c_in, pitch, c_out, out_pitch, choff, relu = case
    g = torch.Generator().manual_seed((hash(case) + 17) & 0xffff)
    x = torch.randn(B, H, W, pitch, generator=g)
    x[..., c_in:] = float("nan")
    scale = 0.5 + torch.rand(c_in, generator=g)
    shift = 0.3 * torch.randn(c_in, generator=g)
    taps = 3 if mode == "3x3" else 1
    w = torch.randn(c_out, c_in, taps, taps, generator=g) / np.sqrt(c_in * taps * taps)
    out, _ = run_conv(lib, cuda, x, c_in, scale, shift, w, mode, relu, precision, out_pitch, choff, False)
    ref = conv_ref(x, c_in, scale, shift, w, mode, relu)
    got = out[..., choff:choff + c_out].double()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err <= TOL[precision], err
    rest = torch.cat([out[..., :choff], out[..., choff + c_out:]], -1)
    assert torch.isnan(rest).all()


def test_conv_argument_errors(lib, cuda):
    from emlight_b200._lib import ConvParams
    x = torch.zeros(1, 8, 8, 24, device=cuda); o = torch.zeros(1, 8, 8, 48, device=cuda)
    p = ConvParams()
    p.in_, p.out = x.data_ptr(), o.data_ptr()
    p.B, p.H, p.W, p.C_in, p.in_pitch, p.C_out, p.out_pitch = 1, 8, 8, 24, 24, 48, 48
    p.precision = 1
    assert lib.eml_conv_forward(p, None) == -1            # wpack missing
    p.in_pitch = 22
    assert lib.eml_conv_forward(p, None) == -3            # pitch not a multiple of 4 / smaller than C_in
    p.in_pitch, p.C_out = 24, 300
    assert lib.eml_conv_forward(p, None) == -2
    p.C_out, p.mode, p.H = 48, 2, 7
    assert lib.eml_conv_forward(p, None) == -2            # odd height cannot be pooled
