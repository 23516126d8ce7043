"""GPU: the one-kernel dense layer (composite 3x3 filter, csrc/dense_layer.cu) through the C ABI vs a float64 torch
restatement of the reference layer (RegressionNetwork/DenseNet.py:26-55, eval-mode BatchNorm) and vs the two-kernel path."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

G, NB = 12, 48
TOL = {"bf16x3": 1e-4, "bf16": 2e-2}      # relative to the output's max magnitude (parity bar: 1e-3 for the fp32-grade mode)


def layer_ref(x, c_in, s1, t1, w1, s2, t2, w2):
    """float64: norm1 -> relu -> conv1 -> norm2 -> conv2 on x (B,H,W,pitch) NHWC; returns (B,H,W,12)."""
    a = (x[..., :c_in].double() * s1.double() + t1.double()).clamp_min(0).permute(0, 3, 1, 2)
    b = F.conv2d(a, w1.double())
    n = b * s2.double().view(1, -1, 1, 1) + t2.double().view(1, -1, 1, 1)
    return F.conv2d(n, w2.double(), padding=1).permute(0, 2, 3, 1).contiguous()


def compose(w1, s2, t2, w2):
    """Host-side composite filter + bias table exactly as emlight_b200.densenet._compose_layers builds them."""
    weff = torch.einsum("obyx,b,bc->yxoc", w2.double(), s2.double(), w1.double().view(NB, -1)).reshape(9 * G, -1).float().contiguous()
    beta = torch.einsum("obyx,b->yxo", w2.double(), t2.double())
    valid = ((1, 2), (0, 1, 2), (0, 1))
    bias9 = torch.stack([torch.stack([beta[list(valid[r])][:, list(valid[c])].sum((0, 1)) for c in range(3)]) for r in range(3)])
    return weff, bias9.float().contiguous()


def run_fused(lib, cuda, x, c_in, s1, t1, w1, s2, t2, w2, precision, rows=None):
    """The product path: eml_dense_layer_compose (composite filter + bias table on the device, written straight into the packed operand)
    then eml_dense_layer_forward.  The device-built bias table is also checked against the host restatement above."""
    from emlight_b200 import _lib
    from emlight_b200._lib import DenseLayerParams
    B, H, W, pitch = x.shape
    xd = x.to(cuda)
    pad = (c_in + 3) & ~3
    sc = torch.zeros(pad, device=cuda); sc[:c_in] = s1.to(cuda)
    sh = torch.zeros(pad, device=cuda); sh[:c_in] = t1.to(cuda)
    wp = torch.empty(lib.eml_dense_layer_wpack_bytes(c_in), dtype=torch.uint8, device=cuda)
    bd = torch.empty(9 * G, device=cuda)
    w1d, w2d, s2d, t2d = w1.reshape(NB, c_in).contiguous().to(cuda), w2.contiguous().to(cuda), s2.to(cuda), t2.to(cuda)   # keep them alive
    _lib.check(lib.eml_dense_layer_compose(_lib.ptr(w1d), _lib.ptr(w2d), _lib.ptr(s2d), _lib.ptr(t2d), NB, c_in, G, _lib.ptr(wp), _lib.ptr(bd),
                                           _lib.stream_ptr()), "eml_dense_layer_compose")
    torch.cuda.synchronize()
    _, bias9 = compose(w1, s2, t2, w2)
    assert float((bd.cpu() - bias9.reshape(-1)).abs().max()) <= 1e-5 * float(bias9.abs().max())
    p = DenseLayerParams()
    p.in_, p.scale, p.shift, p.wpack, p.bias9, p.out = (t.data_ptr() for t in (xd, sc, sh, wp, bd, xd))
    p.B, p.H, p.W, p.C_in, p.in_pitch = B, H, W, c_in, pitch
    p.growth, p.out_pitch, p.out_choff, p.precision = G, pitch, c_in, _lib.PRECISIONS[precision]
    if rows is not None:
        os.environ["EML_DENSE_ROWS"] = str(rows)
    try:
        _lib.check(lib.eml_dense_layer_forward(p, _lib.stream_ptr()), "eml_dense_layer_forward")
        torch.cuda.synchronize()
    finally:
        os.environ.pop("EML_DENSE_ROWS", None)
    return xd.cpu()


CASES = [
    # B, H, W, c_in, pitch, forced band rows
    (1, 2, 128, 24, 216, None),        # smallest image: every row touches a border
    (2, 6, 256, 24, 216, None),        # block-1 geometry, first layer (one k-chunk, two tiles per row)
    (2, 6, 256, 60, 216, 3),           # bands of 3 rows: halo rows recomputed, top/bottom of image in different bands
    (1, 8, 256, 204, 216, 4),          # block-1 last layer: 4 K-chunks (2-stage ring), partial last chunk (12 channels)
    (3, 8, 128, 108, 304, 8),          # block-2 geometry: one tile per row, whole image per band
    (2, 4, 128, 288, 304, 2),          # block-2 last layer: 5 K-chunks; the 4 filler channels end exactly at the pitch
    (40, 16, 128, 132, 300, None),     # 640 rows -> several bands per CTA; pitch not a multiple of 8 -> plain 48-byte stores
    (5, 12, 256, 96, 216, 6),          # odd image count
    (2, 6, 64, 150, 344, None),        # block-3 geometry: a tile = the same row of two images; channel offset only 8-byte aligned
    (6, 8, 64, 318, 344, 4),           # block-3 second-to-last layer: 5 K-chunks, three image pairs, bands of 4 rows
    (4, 48, 64, 174, 344, None),       # full block-3 image height
]


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "B%d-%dx%d-c%d-r%s" % (c[0], c[1], c[2], c[3], c[5]))
def test_dense_layer_matches_reference(lib, cuda, case, precision):
    B, H, W, c_in, pitch, rows = case
    assert lib.eml_dense_layer_supported(H, W, c_in, G, 1) == 1
    g = torch.Generator().manual_seed(1000 + c_in + H)
    x = torch.randn(B, H, W, pitch, generator=g)
    x[..., c_in:] = float("nan")                       # channels beyond C_in must never be consumed
    s1 = 0.5 + torch.rand(c_in, generator=g)
    t1 = 0.3 * torch.randn(c_in, generator=g)
    w1 = torch.randn(NB, c_in, 1, 1, generator=g) / np.sqrt(c_in)
    s2 = 0.5 + torch.rand(NB, generator=g)
    t2 = 0.3 * torch.randn(NB, generator=g)
    w2 = torch.randn(G, NB, 3, 3, generator=g) / np.sqrt(9 * NB)
    ref = layer_ref(x, c_in, s1, t1, w1, s2, t2, w2)
    out = run_fused(lib, cuda, x, c_in, s1, t1, w1, s2, t2, w2, precision, rows)
    got = out[..., c_in:c_in + G].double()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL[precision], err
    # only the 12 new channels were written -- plus, with full-sector stores (pitch % 8 == 0 and c_in % 8 == 0), the 4 channels
    # behind them zeroed (include/emlight_b200.h)
    assert torch.equal(out[..., :c_in], x[..., :c_in])
    tail = c_in + G
    if pitch % 8 == 0 and c_in % 8 == 0 and c_in + G + 4 <= pitch:
        assert (out[..., tail:tail + 4] == 0).all()
        tail += 4
    assert torch.isnan(out[..., tail:]).all()


def test_dense_layer_rejects_unsupported(lib, cuda):
    assert lib.eml_dense_layer_supported(48, 64, 330, G, 1) == 0        # block-3 last layer: 6 weight chunks do not fit
    assert lib.eml_dense_layer_supported(24, 32, 150, G, 1) == 0        # W = 32: two-kernel path
    assert lib.eml_dense_layer_supported(96, 128, 108, 16, 1) == 0      # other growth rates
    assert lib.eml_dense_layer_supported(96, 128, 108, G, 2) == 0       # fp32 SIMT mode has no fused kernel


def test_densenet_fused_equals_two_kernel_path(cuda):
    """Whole network, eval mode: fused dense layers vs conv1 + conv2 kernels (same weights, same input)."""
    import emlight_b200 as E
    torch.manual_seed(5)
    net = E.DenseNet(n_anchors=128, precision="bf16x3").to(cuda).eval()
    with torch.no_grad():
        for m in net.modules():                        # non-trivial running statistics / affine parameters
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
    x = torch.rand(4, 3, 192, 256, generator=torch.Generator().manual_seed(9)).to(cuda)     # even batch: block 3 fuses too (pairs)
    with torch.no_grad():
        net.fuse_dense_layers = True
        a = {k: v.clone() for k, v in net(x).items()}
        net.fuse_dense_layers = False
        b = net(x)
    for k in a:
        err = (a[k] - b[k]).abs().max().item() / b[k].abs().max().item()
        assert err < 1e-4, (k, err)
    assert torch.equal(a["distribution"].argmax(1), b["distribution"].argmax(1))
