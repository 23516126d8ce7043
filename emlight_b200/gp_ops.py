"""Device primitives of the GenProjector TRAINING path (`gp_train.py`), one function per kernel entry point of include/emlight_b200.h.

Every function here launches hand-written CUDA through the C ABI on NHWC fp32 tensors `(B, H, W, pitch)` with `pitch = up4(C)`;
nothing runs on the CPU.  They are kept in one namespace so that `tests/test_gp_train_cpu.py` can swap them for torch stand-ins and
check the tape / backward algebra of `gp_train.py` against torch autograd of the CPU restatement on a box without a GPU -- the product
never does.
The forward-only inference path (`genprojector.py`) does not go through this module.
"""
import torch

from . import _lib
from .genprojector import (_LUTS, _PackedConv, _bias_act, _conv_raw, _im2col, _nchw_to_nhwc, _pool, _reduce, _up4)  # noqa: F401

PackedConv = _PackedConv


def lut(kind, h, w, stride, device):
    """(idx (P,9,4) int32, wgt (P,9,4) fp32, ho, wo): 'sphere' = SphereConv2D's tangent-plane bilinear taps, 'conv' = regular 3x3 pad 1."""
    return _LUTS.get(kind, h, w, stride, device)


def conv_raw(x, B, H, W, pc, lut_, bias_in, act, precision):
    """(B,ho,wo,up4(O)) = Wk * S(act(x + bias_in)): LUT gather + tcgen05 GEMM (genprojector._conv_raw)."""
    return _conv_raw(x, B, H, W, pc, lut_, bias_in, act, precision)


def im2col(x, B, H, W, C, lut_, bias_in, act):
    """fp32 operand A (B*ho*wo, 9*up4(C)) of the convolution above (recomputed in the backward for the weight gradient)."""
    return _im2col(x, B, H, W, C, lut_, bias_in, act)[0]


def bias_act(raw, bias, act, M, C):
    return _bias_act(raw, bias, act, M, C)


def pool(x, B, H, W, C, mode):
    return _pool(x, B, H, W, C, mode)


def nchw_to_nhwc(x, pitch):
    return _nchw_to_nhwc(x, pitch)


def loss_sum(mode, a, M, C, a_pitch, b=None, b_pitch=0, mask=None):
    """float64 scalar tensor: eml_loss_reduce(mode) over M rows x C channels."""
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    _reduce(acc, mode, a, M, C, a_pitch, b, b_pitch, mask)
    return acc


def instance_norm(raw, B, HW, C, lrelu):
    out = torch.empty_like(raw)
    _lib.check(_lib.load().eml_instance_norm(_lib.ptr(raw), raw.shape[-1], _lib.ptr(out), out.shape[-1], B, HW, C, 1e-5, int(lrelu),
                                             _lib.stream_ptr()), "eml_instance_norm")
    return out


def channel_sums(x, M, C):
    """(2, C) float64: per-channel sum and sum of squares over the M rows."""
    sums = torch.zeros(2, C, dtype=torch.float64, device=x.device)
    _lib.check(_lib.load().eml_channel_stats(_lib.ptr(x), x.shape[-1], M, C, _lib.ptr(sums), _lib.stream_ptr()), "eml_channel_stats")
    return sums


def spade_modulate(x, mean, inv, gb, bias_gamma, bias_beta, M, C, lrelu):
    shape = x.shape[:-1] + (_up4(C),)
    out = torch.empty(shape, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().eml_spade_modulate(_lib.ptr(x), x.shape[-1], _lib.ptr(mean.contiguous()), _lib.ptr(inv.contiguous()), _lib.ptr(gb),
                                              gb.shape[-1], _lib.ptr(bias_gamma), _lib.ptr(bias_beta), _lib.ptr(out), out.shape[-1], M, C,
                                              int(lrelu), _lib.stream_ptr()), "eml_spade_modulate")
    return out


def bias_residual(a, bias_a, r, bias_r, M, C):
    shape = a.shape[:-1] + (_up4(C),)
    out = torch.empty(shape, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().eml_bias_residual(_lib.ptr(a), a.shape[-1], _lib.ptr(bias_a), _lib.ptr(r), r.shape[-1] if r is not None else 0,
                                             _lib.ptr(bias_r), _lib.ptr(out), out.shape[-1], M, C, _lib.stream_ptr()), "eml_bias_residual")
    return out


def resize_nearest(x, x_pitch, Hi, Wi, Ho, Wo, C, B, src_is_nchw, out_pitch):
    out = torch.zeros(B, Ho, Wo, out_pitch, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().eml_resize_nearest(_lib.ptr(x), x_pitch, Hi, Wi, _lib.ptr(out), out_pitch, Ho, Wo, C, B, int(src_is_nchw),
                                              _lib.stream_ptr()), "eml_resize_nearest")
    return out


def resize_bilinear_nchw(x, Ho, Wo):
    """(B,3,Hi,Wi) NCHW -> (B,Ho,Wo,4) NHWC, F.interpolate(mode='bilinear', align_corners=False)."""
    B, C, Hi, Wi = x.shape
    out = torch.zeros(B, Ho, Wo, _up4(C), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().eml_resize_bilinear_nchw(_lib.ptr(x.contiguous().float()), Hi, Wi, _lib.ptr(out), out.shape[-1], Ho, Wo, C, B,
                                                    _lib.stream_ptr()), "eml_resize_bilinear_nchw")
    return out


def tanh_to_nchw(raw, bias, B, H, W, C, scale):
    out = torch.empty(B, C, H, W, dtype=torch.float32, device=raw.device)
    _lib.check(_lib.load().eml_tanh_to_nchw(_lib.ptr(raw), raw.shape[-1], _lib.ptr(bias), _lib.ptr(out), B, H * W, C, float(scale),
                                            _lib.stream_ptr()), "eml_tanh_to_nchw")
    return out


def linear(a, w, bias):
    """(M,N) = a (M,K) @ w (N,K)^T + bias: fp32 FFMA kernel (the encoder's fc)."""
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().eml_linear_fp32(_lib.ptr(a.contiguous()), _lib.ptr(w.contiguous()), _lib.ptr(bias), _lib.ptr(out), M, N, K,
                                           _lib.stream_ptr()), "eml_linear_fp32")
    return out


@_lib.on_tensor_device
def mm_nt(a, b, precision="bf16x3"):
    """(M,N) fp32 = a (M,K) @ b (N,K)^T on the TMA-fed tcgen05 GEMM: `a` is split into bf16 hi/lo rows (eml_split_bf16), `b` is packed
    as the resident operand in slices of <= 256 rows (eml_conv_pack_weights); short-and-deep products (few row tiles, long K -- the
    weight gradients, K = pixels) take the split-K kernel so that every SM gets a piece."""
    lib = _lib.load()
    _lib.require_cuda(a, b)
    a = a.contiguous().float()
    b = b.contiguous().float()
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K:
        raise ValueError("mm_nt: inner dimensions differ (%d vs %d)" % (K, b.shape[1]))
    st = _lib.stream_ptr()
    if precision == "fp32":
        out = torch.empty(M, N, dtype=torch.float32, device=a.device)
        _lib.check(lib.eml_linear_fp32(_lib.ptr(a), _lib.ptr(b), None, _lib.ptr(out), M, N, K, st), "eml_linear_fp32(mm_nt %dx%dx%d)" % (M, N, K))
        return out
    Kp = (K + 63) // 64 * 64
    split = precision == "bf16x3"
    a_hi = torch.empty(M, Kp, dtype=torch.bfloat16, device=a.device)
    a_lo = torch.empty_like(a_hi) if split else None
    _lib.check(lib.eml_split_bf16(_lib.ptr(a), M, K, K, _lib.ptr(a_hi), _lib.ptr(a_lo), Kp, st), "eml_split_bf16")
    return mm_nt_split(a_hi, a_lo, M, K, b, precision)


def mm_nt_split(a_hi, a_lo, M, K, b, precision="bf16x3"):
    """(M,N) fp32 = A (M,K) @ b (N,K)^T with A already split into bf16 hi / lo rows of length Kp = round_up(K, 64) (zero padded)."""
    lib = _lib.load()
    st = _lib.stream_ptr()
    b = b.contiguous().float()
    N = b.shape[0]
    Kp = a_hi.shape[1]
    pitch = _up4(N)
    mtiles = (M + 127) // 128
    ksplit = max(1, min(Kp // 64, 148 // mtiles))
    deep = ksplit > 1 and Kp >= 2048
    out = (torch.zeros if deep else torch.empty)(M, pitch, dtype=torch.float32, device=b.device)
    prec = _lib.PRECISIONS[precision]
    n_start = 0
    nfull = N // 256
    if nfull >= 2:
        # all whole 256-row slices of b: one pack launch, one GEMM launch (work item = (slice, K range, 128-row tile))
        sb = lib.eml_conv_wpack_bytes(256, K, 1)
        buf = torch.empty(sb * nfull, dtype=torch.uint8, device=b.device)
        _lib.check(lib.eml_gemm_pack_slices(_lib.ptr(b), _lib.ptr(buf), nfull, 256, K, sb, st), "eml_gemm_pack_slices(mm_nt)")
        _lib.check(lib.eml_gemm_bf16_slices(_lib.ptr(a_hi), _lib.ptr(a_lo), M, Kp, _lib.ptr(buf), sb, nfull, 256, None, _lib.ptr(out), pitch, 0,
                                            prec, ksplit if deep else 1, st), "eml_gemm_bf16_slices(%dx%dx%d)" % (M, nfull * 256, K))
        n_start = nfull * 256
    for n0 in range(n_start, N, 256):
        rows = min(256, N - n0)
        buf = torch.empty(lib.eml_conv_wpack_bytes(rows, K, 1), dtype=torch.uint8, device=b.device)
        _lib.check(lib.eml_conv_pack_weights(_lib.ptr(b[n0:n0 + rows]), _lib.ptr(buf), rows, K, 1, st), "eml_conv_pack_weights(mm_nt)")
        if deep:
            _lib.check(lib.eml_gemm_bf16_splitk(_lib.ptr(a_hi), _lib.ptr(a_lo), M, Kp, _lib.ptr(buf), rows, None, _lib.ptr(out), pitch, n0,
                                                prec, ksplit, st), "eml_gemm_bf16_splitk(%dx%dx%d)" % (M, rows, K))
        else:
            _lib.check(lib.eml_gemm_bf16(_lib.ptr(a_hi), _lib.ptr(a_lo), M, Kp, _lib.ptr(buf), rows, None, _lib.ptr(out), pitch, n0, prec, st),
                       "eml_gemm_bf16(%dx%dx%d)" % (M, rows, K))
    return out[:, :N] if pitch != N else out


# ------------------------------------------------------------------------------------------------- adjoint kernels (csrc/gp_bwd.cu)
def _fn(name):
    return getattr(_lib.load(), name)


def _st():
    return _lib.stream_ptr()


_CSR = {}


def lut_csr(lut_, in_pixels):
    """The sampling table inverted into CSR over input pixels (host, once per table): offs (in_pixels+1) int32, src = p*9 + tap int32,
    w fp32, entries of one input pixel ordered by (p, tap, t) so that the summation order is fixed."""
    idx, wgt, ho, wo = lut_
    key = (idx.data_ptr(), in_pixels, str(idx.device))
    if key not in _CSR:
        i = idx.reshape(-1).cpu().numpy().astype("int64")                   # (P*9*4,)
        w = wgt.reshape(-1).cpu().numpy()
        import numpy as np
        keep = np.nonzero((i >= 0) & (w != 0))[0]
        order = keep[np.argsort(i[keep], kind="stable")]
        counts = np.bincount(i[order], minlength=in_pixels)
        offs = np.concatenate(([0], np.cumsum(counts))).astype("int32")
        src = (order // 4).astype("int32")
        dev = idx.device
        _CSR[key] = (torch.from_numpy(offs).to(dev), torch.from_numpy(src).to(dev), torch.from_numpy(w[order].astype("float32")).to(dev), idx)
    return _CSR[key][:3]


def col2im(dA, Cp, lut_, B, in_pixels):
    """dx (B, in_pixels, Cp) = adjoint of the 4-tap gather applied to dA (B*out_pixels, 9*Cp): gather form over the inverted table
    (eml_col2im_csr; no atomics, deterministic)."""
    idx, wgt, ho, wo = lut_
    offs, src, w = lut_csr(lut_, in_pixels)
    dA = dA.contiguous()
    dx = torch.empty(B, in_pixels, Cp, dtype=torch.float32, device=dA.device)
    _lib.check(_fn("eml_col2im_csr")(_lib.ptr(dA), Cp, _lib.ptr(offs), _lib.ptr(src), _lib.ptr(w), _lib.ptr(dx), Cp, B, ho * wo, in_pixels,
                                     _st()), "eml_col2im_csr")
    return dx


def act_bwd(dx, x, bias, act, M, C, want_sums):
    """dx[..., :C] *= act'(x[..., :C] + bias) in place; returns the per-channel sums of the result (float64) when asked."""
    sums = torch.zeros(C, dtype=torch.float64, device=dx.device) if want_sums else None
    if act or want_sums:
        _lib.check(_fn("eml_act_bwd")(_lib.ptr(dx), dx.shape[-1], _lib.ptr(x), x.shape[-1] if x is not None else 0, _lib.ptr(bias), int(act), M,
                                      C, _lib.ptr(sums), _st()), "eml_act_bwd")
    return sums


def bias_act_bwd(g, out, act, M, C, want_sums):
    g = g.contiguous()
    dx = torch.zeros_like(out)
    sums = torch.zeros(C, dtype=torch.float64, device=out.device) if want_sums else None
    _lib.check(_fn("eml_bias_act_bwd")(_lib.ptr(g), g.shape[-1], _lib.ptr(out), out.shape[-1], int(act), _lib.ptr(dx), dx.shape[-1], M, C,
                                       _lib.ptr(sums), _st()), "eml_bias_act_bwd")
    return dx, sums


def spade_bwd(g, out, x, mean, inv, gb, bias_gamma, M, C, lrelu):
    """(d_gb like gb, d_xhat like x, sums (4,C) float64 = [d bias_gamma, d bias_beta, sum d_xhat, sum d_xhat*xhat])."""
    g = g.contiguous()
    d_gb = torch.zeros_like(gb)
    d_xhat = torch.zeros_like(x)
    sums = torch.zeros(4, C, dtype=torch.float64, device=x.device)
    _lib.check(_fn("eml_spade_bwd")(_lib.ptr(g), g.shape[-1], _lib.ptr(out), out.shape[-1], _lib.ptr(x), x.shape[-1], _lib.ptr(mean.contiguous()),
                                    _lib.ptr(inv.contiguous()), _lib.ptr(gb), gb.shape[-1], _lib.ptr(bias_gamma), _lib.ptr(d_gb), _lib.ptr(d_xhat),
                                    d_xhat.shape[-1], M, C, int(lrelu), _lib.ptr(sums), _st()), "eml_spade_bwd")
    return d_gb, d_xhat, sums


def bn_free_bwd(d_xhat, x, mean, inv, sums2, count, M, C):
    """dx like d_xhat: batch-statistic mode with sums2 (2,C) float64, running-statistic mode with sums2 None."""
    dx = torch.zeros_like(d_xhat)
    if sums2 is not None:
        sums2 = sums2.contiguous()
        mean = mean.contiguous()
    _lib.check(_fn("eml_bn_free_bwd")(_lib.ptr(d_xhat), d_xhat.shape[-1], _lib.ptr(x) if sums2 is not None else None,
                                      x.shape[-1] if sums2 is not None else 0, _lib.ptr(mean) if sums2 is not None else None,
                                      _lib.ptr(inv.contiguous()), _lib.ptr(sums2), float(count), _lib.ptr(dx), dx.shape[-1], M, C, _st()),
               "eml_bn_free_bwd")
    return dx


def instance_norm_bwd(g, out, raw, B, HW, C, lrelu, eps=1e-5):
    g = g.contiguous()
    sums = torch.zeros(B, 4, C, dtype=torch.float64, device=out.device)
    dx = torch.zeros_like(raw)
    _lib.check(_fn("eml_instance_norm_bwd")(_lib.ptr(g), g.shape[-1], _lib.ptr(out), out.shape[-1], _lib.ptr(raw), raw.shape[-1], B, HW, C,
                                            float(eps), int(lrelu), _lib.ptr(sums), _lib.ptr(dx), dx.shape[-1], _st()), "eml_instance_norm_bwd")
    return dx


def upsample2_bwd(g, B, H, W, C):
    """dx (B,H,W,pitch of g) = sum over the 2x2 blocks of g (B,2H,2W,pitch)."""
    g = g.contiguous()
    dx = torch.zeros(B, H, W, g.shape[-1], dtype=torch.float32, device=g.device)
    _lib.check(_fn("eml_upsample2_bwd")(_lib.ptr(g), g.shape[-1], _lib.ptr(dx), dx.shape[-1], B, H, W, C, _st()), "eml_upsample2_bwd")
    return dx


def tanh_nchw_bwd(g_nchw, out_nchw, scale, B, HW, C, pitch, want_sums):
    """(d_raw (B*HW rows, pitch) NHWC, bias sums float64 or None) from the NCHW gradient / output of tanh_to_nchw."""
    g_nchw = g_nchw.contiguous().float()
    d_raw = torch.zeros(B, HW, pitch, dtype=torch.float32, device=g_nchw.device)
    sums = torch.zeros(C, dtype=torch.float64, device=g_nchw.device) if want_sums else None
    _lib.check(_fn("eml_tanh_nchw_bwd")(_lib.ptr(g_nchw), _lib.ptr(out_nchw), float(scale), _lib.ptr(d_raw), pitch, B, HW, C, _lib.ptr(sums),
                                        _st()), "eml_tanh_nchw_bwd")
    return d_raw, sums


def pool2d_bwd(g, x, B, H, W, C, mode):
    """dx like x (B,H,W,pitch): adjoint of pool(x, mode) given g on the pooled grid."""
    g = g.contiguous()
    dx = torch.zeros_like(x)
    _lib.check(_fn("eml_pool2d_bwd")(_lib.ptr(g), g.shape[-1], _lib.ptr(x), x.shape[-1], _lib.ptr(dx), dx.shape[-1], H, W, C, B, int(mode),
                                     _st()), "eml_pool2d_bwd")
    return dx


def loss_seed(mode, a, M, C, coef, coef_dev, b=None, mask=None):
    """da like a = coef * coef_dev[0] * d(loss_sum(mode))/da (coef_dev: 1-element float32 device tensor or None)."""
    da = torch.zeros_like(a)
    _lib.check(_fn("eml_loss_seed")(_lib.ptr(a), a.shape[-1], _lib.ptr(b), b.shape[-1] if b is not None else 0, _lib.ptr(mask), M, C, int(mode),
                                    float(coef), _lib.ptr(coef_dev), _lib.ptr(da), da.shape[-1], _st()), "eml_loss_seed")
    return da


def im2col_t(x, B, H, W, C, lut_, bias_in, act, split=True):
    """(At_hi, At_lo or None): the im2col operand TRANSPOSED, (9*up4(C), Mp) bf16 with Mp = round_up(B*out_pixels, 64), for the
    weight-gradient GEMM (K = pixels) -- no fp32 matrix, no transpose pass."""
    idx, wgt, ho, wo = lut_
    Cp = _up4(C)
    M = B * ho * wo
    Mp = (M + 63) // 64 * 64
    if (bias_in is not None or act) and C == Cp and x.shape[-1] % 4 == 0:
        # the input transform once per value, then the tiled (shared-memory transposed) form of the gather
        x, bias_in, act = bias_act(x, bias_in, int(act), B * H * W, C), None, 0
    hi = torch.empty(9 * Cp, Mp, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi) if split else None
    if Mp > M:                                                     # the K padding of the weight-gradient GEMM
        hi[:, M:].zero_()
        if split:
            lo[:, M:].zero_()
    _lib.check(_fn("eml_im2col_lut_bf16_t")(_lib.ptr(x), x.shape[-1], C, Cp, _lib.ptr(idx), _lib.ptr(wgt), _lib.ptr(bias_in), int(act),
                                            _lib.ptr(hi), _lib.ptr(lo), Mp, B, ho * wo, H * W, _st()), "eml_im2col_lut_bf16_t")
    return hi, lo
