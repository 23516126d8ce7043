"""GenProjector SPADE / SphereNet generator -- drop-in for ``GenProjector/models/networks`` (inference path), sm_100a.

Classes keep the reference's names, constructor arguments and ``state_dict`` keys:

  SphereConv2D(in_c, out_c, stride=1, bias=True, mode='bilinear')      spherenet/sphere_cnn.py:87-124
  SPADE(config_text, norm_nc, label_nc)                                normalization.py:68-115
  SPADEResnetBlock(fin, fout, opt)                                     architecture.py:22-69
  ConvEncoder(opt)                                                     generator.py:90-126
  SPADEGenerator(opt).forward(input, crop) -> (B,3,128,256) in [0,50]  generator.py:17-88

Spectral-normalised layers carry ``weight_orig / weight_u / weight_v`` exactly like ``torch.nn.utils.spectral_norm`` so
reference checkpoints load; in eval mode the weight is ``weight_orig / (u . W v)`` (no power iteration), folded into the
packed bf16 hi/lo weight images at pack time.  Every convolution is LUT gather (``eml_im2col_lut``) + tcgen05 GEMM
(``eml_conv_forward``); normalisation / modulation / resize / tanh are the small kernels of ``csrc/spade_ops.cu``.

Forward values in both modes: eval (what ``GenProjector/test.py:21-39`` runs: running-statistics BatchNorm inside SPADE, stored
spectral-norm vectors) and train (batch-statistic (Sync)BatchNorm with running-stat update, one power iteration per spectral
convolution per forward).  Backward is not built: outputs carry no autograd graph.
"""
import math
import os
import re
from functools import lru_cache

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib
from ._lib import ConvParams

_NHIDDEN = 128
_MAX_N = 256          # output channels per GEMM launch (TMEM columns of the generic implicit-GEMM kernel)


def _up4(n):
    return (n + 3) & ~3


# ----------------------------------------------------------------------------------------------------- sampling tables
@lru_cache(maxsize=None)
def _sphere_coords(h, w, stride):
    """(Ho,Wo,3,3,2) float64 (row, col) sample positions of SphereConv2D (sphere_cnn.py:11-58): gnomonic projection of a
    3x3 tangent patch with d_phi = pi/h, d_theta = 2pi/w around every `stride`-th pixel; centre forced to the pixel itself
    (:57), longitude wrapped modulo w (:55)."""
    dphi, dth = math.pi / h, 2 * math.pi / w
    tx, ty = math.tan(dth), math.tan(dphi)
    sy = ty / math.cos(dth)
    px = np.array([[-tx, 0.0, tx]] * 3)                                  # get_xy(): x offsets per column
    px[1, 1] = 1.0
    py = np.array([[sy, ty, sy], [0.0, 1.0, 0.0], [-sy, -ty, -sy]])
    rr = np.arange(0, h, stride, dtype=np.float64).reshape(-1, 1, 1, 1)
    cc = np.arange(0, w, stride, dtype=np.float64).reshape(1, -1, 1, 1)
    phi = -((rr + 0.5) / h * math.pi - math.pi / 2)
    theta = (cc + 0.5) / w * 2 * math.pi - math.pi
    rho = np.sqrt(px ** 2 + py ** 2)
    nu = np.arctan(rho)
    nphi = np.arcsin(np.cos(nu) * np.sin(phi) + py * np.sin(nu) * np.cos(phi) / rho)
    nth = theta + np.arctan(px * np.sin(nu) / (rho * np.cos(phi) * np.cos(nu) - py * np.sin(phi) * np.sin(nu)))
    r = (-nphi + math.pi / 2) * h / math.pi - 0.5
    c = ((nth + math.pi) * w / 2 / math.pi - 0.5 + w) % w
    r, c = np.broadcast_arrays(r, c)
    out = np.stack((r, c), -1).copy()
    out[:, :, 1, 1, 0] = np.arange(0, h, stride).reshape(-1, 1)
    out[:, :, 1, 1, 1] = np.arange(0, w, stride).reshape(1, -1)
    return out


@lru_cache(maxsize=None)
def _sphere_lut(h, w, stride):
    """4 bilinear taps per (output pixel, filter tap) reproducing grid_sample(bilinear, zeros, align_corners=False) on the
    grid the reference builds (sphere_cnn.py:75-84): normalised g = 2p/size - 1 in fp32, un-normalised ((g+1)*size-1)/2."""
    co = _sphere_coords(h, w, stride)
    gy = (co[..., 0] * 2 / h - 1).astype(np.float32)
    gx = (co[..., 1] * 2 / w - 1).astype(np.float32)
    iy = ((gy + np.float32(1)) * np.float32(h) - np.float32(1)) / np.float32(2)
    ix = ((gx + np.float32(1)) * np.float32(w) - np.float32(1)) / np.float32(2)
    y0 = np.floor(iy); x0 = np.floor(ix)
    wy1 = (iy - y0).astype(np.float32); wx1 = (ix - x0).astype(np.float32)
    wy0 = (y0 + np.float32(1) - iy).astype(np.float32); wx0 = (x0 + np.float32(1) - ix).astype(np.float32)
    y0 = y0.astype(np.int64); x0 = x0.astype(np.int64)
    idx = np.full(co.shape[:4] + (4,), -1, np.int32)
    wgt = np.zeros(co.shape[:4] + (4,), np.float32)
    for t, (yy, xx, ww) in enumerate(((y0, x0, wy0 * wx0), (y0, x0 + 1, wy0 * wx1), (y0 + 1, x0, wy1 * wx0), (y0 + 1, x0 + 1, wy1 * wx1))):
        ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
        idx[..., t] = np.where(ok, yy * w + xx, -1)
        wgt[..., t] = np.where(ok, ww, 0)
    ho, wo = co.shape[:2]
    return idx.reshape(ho * wo, 9, 4), wgt.reshape(ho * wo, 9, 4), ho, wo


@lru_cache(maxsize=None)
def _conv_lut(h, w, stride=2):
    """Regular 3x3, padding 1 convolution of stride 1 (VGG19) or 2 (ConvEncoder, generator.py:100-104) as a one-tap table."""
    ho, wo = (h + stride - 1) // stride, (w + stride - 1) // stride
    yo, xo = np.meshgrid(np.arange(ho), np.arange(wo), indexing="ij")
    idx = np.full((ho, wo, 9, 4), -1, np.int32)
    wgt = np.zeros((ho, wo, 9, 4), np.float32)
    for ky in range(3):
        for kx in range(3):
            yy, xx = stride * yo + ky - 1, stride * xo + kx - 1
            ok = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
            idx[:, :, ky * 3 + kx, 0] = np.where(ok, yy * w + xx, -1)
            wgt[:, :, ky * 3 + kx, 0] = ok
    return idx.reshape(ho * wo, 9, 4), wgt.reshape(ho * wo, 9, 4), ho, wo


class _Luts:
    def __init__(self):
        self.cache = {}

    def get(self, kind, h, w, stride, device):
        key = (kind, h, w, stride, str(device))
        if key not in self.cache:
            idx, wgt, ho, wo = _sphere_lut(h, w, stride) if kind == "sphere" else _conv_lut(h, w, stride)
            self.cache[key] = (torch.from_numpy(idx).to(device), torch.from_numpy(wgt).to(device), ho, wo)
        return self.cache[key]


_LUTS = _Luts()


# ----------------------------------------------------------------------------------------------------- low-level ops
class _PackedConv:
    """A 3x3 (sphere or regular) convolution's weights as GEMM operands: K = 9*Cp, output sliced into <= 256 channels."""

    def __init__(self, weight, precision):
        lib = _lib.load()
        O, C = weight.shape[0], weight.shape[1]
        self.C, self.Cp, self.O = C, _up4(C), O
        K = 9 * self.Cp
        if precision != "fp32":
            K = (K + 63) & ~63                       # the TMA GEMM walks K in 64-wide chunks; pad columns are zero on both sides
        wk = torch.zeros(O, K, dtype=torch.float32, device=weight.device)
        wk[:, :9 * self.Cp].view(O, 9, self.Cp)[:, :, :C] = weight.detach().float().permute(0, 2, 3, 1).reshape(O, 9, C)   # (O, tap, c)
        self.wk = wk
        self.K = K
        self.slices = []
        st = _lib.stream_ptr()
        # equal slices live in ONE buffer at a fixed stride, so eml_gemm_bf16_slices can run them all in one launch
        self.pack_all, self.slice_bytes = None, 0
        if precision != "fp32" and O > _MAX_N and O % _MAX_N == 0:
            self.slice_bytes = lib.eml_conv_wpack_bytes(_MAX_N, K, 1)
            self.pack_all = torch.empty(self.slice_bytes * (O // _MAX_N), dtype=torch.uint8, device=weight.device)
            _lib.check(lib.eml_gemm_pack_slices(_lib.ptr(self.wk), _lib.ptr(self.pack_all), O // _MAX_N, _MAX_N, K, self.slice_bytes, st),
                       "eml_gemm_pack_slices")
        for n0 in range(0, O, _MAX_N):
            n = min(_MAX_N, O - n0)
            w_s = self.wk[n0:n0 + n]                   # a row range of a contiguous matrix: contiguous
            pack = None
            if precision != "fp32":
                if self.pack_all is not None:
                    pack = self.pack_all[(n0 // _MAX_N) * self.slice_bytes:(n0 // _MAX_N + 1) * self.slice_bytes]
                else:
                    pack = torch.empty(lib.eml_conv_wpack_bytes(n, K, 1), dtype=torch.uint8, device=weight.device)
                    _lib.check(lib.eml_conv_pack_weights(_lib.ptr(w_s), _lib.ptr(pack), n, K, 1, st), "eml_conv_pack_weights")
            self.slices.append((n0, n, w_s, pack))


def _gemm(A, M, pc, out, out_pitch, choff, precision):
    """out[:, choff:choff+O] = A (M,K) @ W^T via eml_conv_forward (1x1 mode), sliced over the output channels."""
    lib = _lib.load()
    if precision == "fp32" and pc.K * 16 * 4 > 200 * 1024:
        # the SIMT conv kernel keeps a 16-channel weight panel in shared memory (K <= 3200); wider reductions (ngf=64 generator
        # blocks, VGG conv4/5) take the tiled fp32 GEMM and are copied into the output slab
        for n0, n, w_s, _ in pc.slices:
            tmp = torch.empty(M, n, dtype=torch.float32, device=A.device)
            _lib.check(lib.eml_linear_fp32(_lib.ptr(A), _lib.ptr(w_s), None, _lib.ptr(tmp), M, n, pc.K, _lib.stream_ptr()),
                       "eml_linear_fp32(GEMM %dx%dx%d)" % (M, n, pc.K))
            out.view(M, out_pitch)[:, choff + n0:choff + n0 + n] = tmp
        return
    for n0, n, w_s, pack in pc.slices:
        p = ConvParams()
        p.in_ = A.data_ptr(); p.scale = None; p.shift = None
        p.w_oihw = w_s.data_ptr(); p.wpack = pack.data_ptr() if pack is not None else None
        p.out = out.data_ptr(); p.stats = None; p.stats_stride = 0
        p.B, p.H, p.W = 1, 1, M
        p.C_in, p.in_pitch = pc.K, pc.K
        p.C_out, p.out_pitch, p.out_choff = n, out_pitch, choff + n0
        p.mode, p.relu, p.precision = _lib.EML_CONV_1x1, 0, _lib.PRECISIONS[precision]
        _lib.check(lib.eml_conv_forward(p, _lib.stream_ptr()), "eml_conv_forward(GEMM %dx%dx%d)" % (M, n, pc.K))


def _im2col(x, B, H, W, C, lut, bias, act):
    """x NHWC (B,H,W,pitch) -> A (B*Ho*Wo, 9*Cp) fp32."""
    lib = _lib.load()
    idx, wgt, ho, wo = lut
    Cp = _up4(C)
    A = torch.empty(B * ho * wo, 9 * Cp, dtype=torch.float32, device=x.device)
    _lib.check(lib.eml_im2col_lut(_lib.ptr(x), x.shape[-1], C, Cp, _lib.ptr(idx), _lib.ptr(wgt), _lib.ptr(bias), act, _lib.ptr(A),
                                  B, ho * wo, H * W, _lib.stream_ptr()), "eml_im2col_lut")
    return A, ho, wo


def _conv_raw(x, B, H, W, pc, lut, bias_in, act, precision):
    """Bias-free 3x3 (sphere or regular) convolution on an NHWC tensor: (B,Ho,Wo,up4(O)) = W * S(act(x + bias_in)).
    fp32: fp32 im2col + the SIMT GEMM mode; bf16 / bf16x3: bf16 hi(/lo) im2col + the TMA-fed tcgen05 GEMM per <=256-channel slice."""
    lib = _lib.load()
    idx, wgt, ho, wo = lut
    M = B * ho * wo
    out = torch.empty(B, ho, wo, _up4(pc.O), dtype=torch.float32, device=x.device)
    if precision == "fp32":
        A, _, _ = _im2col(x, B, H, W, pc.C, lut, bias_in, act)
        _gemm(A, M, pc, out, out.shape[-1], 0, precision)
        return out
    split = precision == "bf16x3"
    st = _lib.stream_ptr()
    if (bias_in is not None or act) and pc.C == pc.Cp and x.shape[-1] % 4 == 0:
        # the input transform once per value (not once per filter tap x bilinear tap): the gather then takes its plain fast path
        xt = torch.empty(B, H, W, pc.Cp, dtype=torch.float32, device=x.device)
        _lib.check(lib.eml_bias_act(_lib.ptr(x), x.shape[-1], _lib.ptr(bias_in), act, _lib.ptr(xt), pc.Cp, B * H * W, pc.C, st), "eml_bias_act")
        x, bias_in, act = xt, None, 0
    a_hi = torch.empty(M, pc.K, dtype=torch.bfloat16, device=x.device)
    a_lo = torch.empty(M, pc.K, dtype=torch.bfloat16, device=x.device) if split else None
    _lib.check(lib.eml_im2col_lut_bf16(_lib.ptr(x), x.shape[-1], pc.C, pc.Cp, _lib.ptr(idx), _lib.ptr(wgt), _lib.ptr(bias_in), act,
                                       _lib.ptr(a_hi), _lib.ptr(a_lo), pc.K, B, ho * wo, H * W, st), "eml_im2col_lut_bf16")
    if pc.pack_all is not None:
        _lib.check(lib.eml_gemm_bf16_slices(_lib.ptr(a_hi), _lib.ptr(a_lo), M, pc.K, _lib.ptr(pc.pack_all), pc.slice_bytes, len(pc.slices), _MAX_N,
                                            None, _lib.ptr(out), out.shape[-1], 0, _lib.PRECISIONS[precision], 1, st),
                   "eml_gemm_bf16_slices(%dx%dx%d)" % (M, pc.O, pc.K))
        return out
    for n0, n, w_s, pack in pc.slices:
        _lib.check(lib.eml_gemm_bf16(_lib.ptr(a_hi), _lib.ptr(a_lo), M, pc.K, _lib.ptr(pack), n, None, _lib.ptr(out), out.shape[-1], n0,
                                     _lib.PRECISIONS[precision], st), "eml_gemm_bf16(%dx%dx%d)" % (M, n, pc.K))
    return out


def _sphere_conv_raw(x, B, H, W, pc, stride, bias_in, act, precision):
    return _conv_raw(x, B, H, W, pc, _LUTS.get("sphere", H, W, stride, x.device), bias_in, act, precision)


# ----------------------------------------------------------------------------------------------------- modules
class SphereConv2D(nn.Module):
    """Drop-in for sphere_cnn.SphereConv2D (3x3 only, bilinear).  NCHW in / NCHW out like the reference."""

    def __init__(self, in_c, out_c, stride=1, bias=True, mode="bilinear"):
        super().__init__()
        if mode != "bilinear":
            raise ValueError("only mode='bilinear' is implemented (the mode EMLight uses)")
        self.in_c, self.out_c, self.stride, self.mode = in_c, out_c, stride, mode
        self.weight = Parameter(torch.Tensor(out_c, in_c, 3, 3))
        if bias:
            self.bias = Parameter(torch.Tensor(out_c))
        else:
            self.register_parameter("bias", None)
        self.precision = "bf16x3"
        self.autograd = False            # opt-in: forward recorded on a tape so that .backward() reaches weight, bias and input (gp_train.py)
        self.reset_parameters()
        self._pc = None

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=np.sqrt(5))
        if self.bias is not None:
            self.bias.data.zero_()

    def effective_weight(self):
        return self.weight

    def _tracked(self):
        return (self.weight,)

    def packed(self, precision):
        key = (precision,) + tuple((t.data_ptr(), t._version, str(t.device)) for t in self._tracked())
        if self._pc is None or self._pc[0] != key:
            with torch.no_grad():
                self._pc = (key, _PackedConv(self.effective_weight(), precision))
        return self._pc[1]

    @_lib.on_tensor_device
    def forward(self, x):
        _lib.require_cuda(x)
        if self.autograd and torch.is_grad_enabled():
            from . import gp_train
            return gp_train.sphere_conv_module(self, x)
        with torch.no_grad():
            return self._forward_values(x)

    def _forward_values(self, x):
        lib = _lib.load()
        B, C, H, W = x.shape
        xn = x.float().permute(0, 2, 3, 1).contiguous()
        pc = self.packed(self.precision)
        raw = _sphere_conv_raw(xn, B, H, W, pc, self.stride, None, 0, self.precision)
        ho, wo = raw.shape[1], raw.shape[2]
        out = torch.empty_like(raw)
        _lib.check(lib.eml_bias_residual(_lib.ptr(raw), raw.shape[-1], _lib.ptr(self.bias), None, 0, None, _lib.ptr(out), out.shape[-1],
                                         B * ho * wo, self.out_c, _lib.stream_ptr()), "eml_bias_residual")
        return out[..., :self.out_c].permute(0, 3, 1, 2).contiguous()


_TORCH_SPECTRAL = os.environ.get("EML_TORCH_SPECTRAL") == "1"     # A/B: the tensor-op formulation


def _sn_kernel_ok(wm):
    return wm.is_cuda and wm.dtype == torch.float32 and wm.is_contiguous() and not _TORCH_SPECTRAL


def _sn_sigma_kernel(module, wm, update):
    """sigma (0-dim tensor) of the (O, K) weight `wm` through eml_spectral_norm: one C-ABI call (four small kernels) instead of ~12
    tensor ops per wrapped convolution and forward; with `update` the power iteration rewrites module.weight_u / weight_v in place."""
    O, K = wm.shape
    scratch = torch.empty(O + K + 1, dtype=torch.float32, device=wm.device)
    sigma = scratch[O + K:]
    with torch.no_grad():
        _lib.check(_lib.load().eml_spectral_norm(_lib.ptr(wm), O, K, _lib.ptr(module.weight_u), _lib.ptr(module.weight_v), int(bool(update)),
                                                 1e-12, _lib.ptr(scratch), _lib.ptr(sigma), _lib.stream_ptr()), "eml_spectral_norm")
        if update:
            torch.autograd.graph.increment_version(module.weight_u)
            torch.autograd.graph.increment_version(module.weight_v)
    return sigma[0]


def _spectral_sigma(module, w):
    """sigma of torch.nn.utils.spectral_norm: in training mode ONE power iteration first (u, v updated in place, eps 1e-12), in eval
    mode the stored vectors (GenProjector/models/networks/architecture.py:37-40, normalization.py:29 wrap their convs with it)."""
    wm = w.reshape(w.shape[0], -1)
    if _sn_kernel_ok(wm):
        return _sn_sigma_kernel(module, wm, module.training)
    if module.training:
        v = nn.functional.normalize(torch.mv(wm.t(), module.weight_u), dim=0, eps=1e-12)
        u = nn.functional.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
        module.weight_v.copy_(v)
        module.weight_u.copy_(u)
    return torch.dot(module.weight_u, torch.mv(wm, module.weight_v))


class _SpectralSphereConv2D(SphereConv2D):
    """SphereConv2D under torch.nn.utils.spectral_norm naming: weight_orig (parameter), weight_u / weight_v (buffers)."""

    def __init__(self, in_c, out_c, stride=1, bias=True):
        super().__init__(in_c, out_c, stride=stride, bias=bias)
        w = self.weight
        del self._parameters["weight"]
        self.register_parameter("weight_orig", Parameter(w.data))
        self.register_buffer("weight_u", nn.functional.normalize(torch.randn(out_c), dim=0))
        self.register_buffer("weight_v", nn.functional.normalize(torch.randn(in_c * 9), dim=0))

    def effective_weight(self):
        w = self.weight_orig
        return w / _spectral_sigma(self, w)

    def _tracked(self):
        return (self.weight_orig, self.weight_u, self.weight_v)

    def packed(self, precision):
        if self.training:                          # the power iteration changes u, v every forward: nothing to cache
            with torch.no_grad():
                return _PackedConv(self.effective_weight(), precision)
        return super().packed(precision)


class _SpectralConv2d(nn.Module):
    """nn.Conv2d(3x3, stride 2, padding 1, no bias) under spectral_norm naming (ConvEncoder layers, generator.py:100-104)."""

    def __init__(self, in_c, out_c):
        super().__init__()
        self.in_channels, self.out_channels = in_c, out_c
        w = torch.empty(out_c, in_c, 3, 3)
        nn.init.kaiming_uniform_(w, a=math.sqrt(5))
        self.register_parameter("bias", None)
        self.weight_orig = Parameter(w)
        self.register_buffer("weight_u", nn.functional.normalize(torch.randn(out_c), dim=0))
        self.register_buffer("weight_v", nn.functional.normalize(torch.randn(in_c * 9), dim=0))
        self._pc = None

    def packed(self, precision):
        w = self.weight_orig
        if self.training:
            with torch.no_grad():
                return _PackedConv(w / _spectral_sigma(self, w), precision)
        key = (w.data_ptr(), w._version, self.weight_u._version, self.weight_v._version, precision, str(w.device))
        if self._pc is None or self._pc[0] != key:
            with torch.no_grad():
                self._pc = (key, _PackedConv(w / _spectral_sigma(self, w), precision))
        return self._pc[1]


class _ParamFreeBN(nn.Module):
    """SynchronizedBatchNorm2d(affine=False) as SPADE uses it: buffers only (normalization.py:80)."""

    def __init__(self, c, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = c, eps, momentum
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class SPADE(nn.Module):
    def __init__(self, config_text, norm_nc, label_nc):
        super().__init__()
        parsed = re.search(r"spade(\D+)(\d)x\d", config_text)
        if parsed is None or parsed.group(1) not in ("syncbatch", "batch"):
            raise ValueError("only the spade(sync)batch3x3 configuration EMLight uses is implemented, got %r" % config_text)
        self.norm_nc, self.label_nc = norm_nc, label_nc
        self.param_free_norm = _ParamFreeBN(norm_nc)
        self.mlp_shared = nn.Sequential(SphereConv2D(label_nc, _NHIDDEN), nn.ReLU())
        self.mlp_gamma = SphereConv2D(_NHIDDEN, norm_nc)
        self.mlp_beta = SphereConv2D(_NHIDDEN, norm_nc)
        self._gb = None

    def _packed_gb(self, precision):
        g, b = self.mlp_gamma.weight, self.mlp_beta.weight
        key = (g.data_ptr(), g._version, b.data_ptr(), b._version, precision, str(g.device))
        if self._gb is None or self._gb[0] != key:
            with torch.no_grad():
                self._gb = (key, _PackedConv(torch.cat([g, b], 0), precision))       # one GEMM: gamma | beta
        return self._gb[1]

    @torch.no_grad()
    def apply_nhwc(self, x, B, H, W, seg, x_bias, lrelu, precision):
        """x (B,H,W,pitch) raw conv output whose bias `x_bias` (or None) has not been added yet; seg (B,H,W,4) resized guide."""
        lib = _lib.load()
        C = self.norm_nc
        bn = self.param_free_norm
        if self.training:
            # batch statistics over (B, H, W) of x + x_bias (normalization.py:80,104; SynchronizedBatchNorm2d = one all-reduce of the
            # 2C sums when several processes share the batch), then the running-statistics update of nn.BatchNorm (momentum 0.1)
            M = B * H * W
            sums = torch.zeros(2, C, dtype=torch.float64, device=x.device)
            _lib.check(lib.eml_channel_stats(_lib.ptr(x), x.shape[-1], M, C, _lib.ptr(sums), _lib.stream_ptr()), "eml_channel_stats")
            n = float(M)
            if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
                from .parallel import global_count
                torch.distributed.all_reduce(sums)
                n = float(global_count(B, x.device) * H * W)                                # shards may differ by one sample
            m_raw = sums[0] / n
            var = (sums[1] / n - m_raw * m_raw).clamp_min(0.0)
            mean = m_raw.float()                                                           # the bias shifts the mean and cancels
            inv = torch.rsqrt(var.float() + bn.eps)
            m_full = mean if x_bias is None else mean + x_bias
            bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * m_full)
            bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * (var * (n / max(n - 1.0, 1.0))).float())
            # (num_batches_tracked stays untouched: the reference's SynchronizedBatchNorm2d never increments it)
        else:
            mean = bn.running_mean if x_bias is None else bn.running_mean - x_bias        # (x + b - m) = x - (m - b)
            inv = torch.rsqrt(bn.running_var + bn.eps)
        shared = self.mlp_shared[0]
        actv = _sphere_conv_raw(seg, B, H, W, shared.packed(precision), 1, None, 0, precision)
        gb = _sphere_conv_raw(actv, B, H, W, self._packed_gb(precision), 1, shared.bias, 1, precision)      # relu(actv + bias)
        out = torch.empty(B, H, W, _up4(C), dtype=torch.float32, device=x.device)
        _lib.check(lib.eml_spade_modulate(_lib.ptr(x), x.shape[-1], _lib.ptr(mean.contiguous()), _lib.ptr(inv.contiguous()), _lib.ptr(gb),
                                          gb.shape[-1], _lib.ptr(self.mlp_gamma.bias), _lib.ptr(self.mlp_beta.bias), _lib.ptr(out),
                                          out.shape[-1], B * H * W, C, int(lrelu), _lib.stream_ptr()), "eml_spade_modulate")
        return out


class SPADEResnetBlock(nn.Module):
    def __init__(self, fin, fout, opt):
        super().__init__()
        self.learned_shortcut = fin != fout
        fmiddle = min(fin, fout)
        spectral = "spectral" in opt.norm_G
        conv = _SpectralSphereConv2D if spectral else SphereConv2D
        self.conv_0 = conv(fin, fmiddle)
        self.conv_1 = conv(fmiddle, fout)
        if self.learned_shortcut:
            self.conv_s = conv(fin, fout)
        cfg = opt.norm_G.replace("spectral", "")
        self.norm_0 = SPADE(cfg, fin, opt.semantic_nc)
        self.norm_1 = SPADE(cfg, fmiddle, opt.semantic_nc)
        if self.learned_shortcut:
            self.norm_s = SPADE(cfg, fin, opt.semantic_nc)
        self.fin, self.fout, self.fmiddle = fin, fout, fmiddle

    @torch.no_grad()
    def apply_nhwc(self, x, B, H, W, seg, precision):
        """x (B,H,W,pitch) finished activations -> (B,H,W,up4(fout))."""
        lib = _lib.load()
        r, r_bias = x, None
        if self.learned_shortcut:
            s = self.norm_s.apply_nhwc(x, B, H, W, seg, None, False, precision)
            r = _sphere_conv_raw(s, B, H, W, self.conv_s.packed(precision), 1, None, 0, precision)
            r_bias = self.conv_s.bias
        h = self.norm_0.apply_nhwc(x, B, H, W, seg, None, True, precision)
        d0 = _sphere_conv_raw(h, B, H, W, self.conv_0.packed(precision), 1, None, 0, precision)
        h = self.norm_1.apply_nhwc(d0, B, H, W, seg, self.conv_0.bias, True, precision)
        d1 = _sphere_conv_raw(h, B, H, W, self.conv_1.packed(precision), 1, None, 0, precision)
        out = torch.empty(B, H, W, _up4(self.fout), dtype=torch.float32, device=x.device)
        _lib.check(lib.eml_bias_residual(_lib.ptr(d1), d1.shape[-1], _lib.ptr(self.conv_1.bias), _lib.ptr(r), r.shape[-1], _lib.ptr(r_bias),
                                         _lib.ptr(out), out.shape[-1], B * H * W, self.fout, _lib.stream_ptr()), "eml_bias_residual")
        return out


class ConvEncoder(nn.Module):
    def __init__(self, opt):
        super().__init__()
        ndf = opt.ngf
        if getattr(opt, "norm_E", "spectralinstance") != "spectralinstance":
            raise ValueError("only norm_E='spectralinstance' (the reference default) is implemented")
        chans = [3, ndf, ndf * 2, ndf * 4, ndf * 8, ndf * 8]
        for i in range(1, 6):
            setattr(self, "layer%d" % i, nn.Sequential(_SpectralConv2d(chans[i - 1], chans[i]), nn.InstanceNorm2d(chans[i], affine=False)))
        self.fc = nn.Linear(ndf * 8 * 4 * 4, 16 * ndf * 2 * 1)
        self.actvn = nn.LeakyReLU(0.2, False)
        self.opt = opt
        self._fc = None

    @torch.no_grad()
    def encode(self, crop, precision):
        """crop (B,3,Hc,Wc) NCHW -> z (B, 32*ngf)."""
        lib = _lib.load()
        B = crop.shape[0]
        dev = crop.device
        st = _lib.stream_ptr()
        x = torch.empty(B, 128, 128, 4, dtype=torch.float32, device=dev)
        _lib.check(lib.eml_resize_bilinear_nchw(_lib.ptr(crop.contiguous().float()), crop.shape[2], crop.shape[3], _lib.ptr(x), 4, 128, 128, 3, B, st),
                   "eml_resize_bilinear_nchw")
        H = W = 128
        C = 3
        for i in range(1, 6):
            conv = getattr(self, "layer%d" % i)[0]
            pc = conv.packed(precision)
            raw = _conv_raw(x, B, H, W, pc, _LUTS.get("conv", H, W, 2, dev), None, 0, precision)   # LeakyReLU already applied by the norm below
            ho, wo = raw.shape[1], raw.shape[2]
            x = torch.empty_like(raw)
            _lib.check(lib.eml_instance_norm(_lib.ptr(raw), raw.shape[-1], _lib.ptr(x), x.shape[-1], B, ho * wo, pc.O, 1e-5, 1, st),
                       "eml_instance_norm")
            H, W, C = ho, wo, pc.O
        # fc consumes the NCHW flattening (generator.py:124); ours is NHWC -> permute the weight columns once
        key = (self.fc.weight.data_ptr(), self.fc.weight._version, str(dev))
        if self._fc is None or self._fc[0] != key:
            wf = self.fc.weight.detach().float().view(self.fc.out_features, C, H * W).permute(0, 2, 1).reshape(self.fc.out_features, -1).contiguous()
            self._fc = (key, wf)
        flat = x[..., :C].reshape(B, -1) if x.shape[-1] != C else x.reshape(B, -1)
        z = torch.empty(B, self.fc.out_features, dtype=torch.float32, device=dev)
        _lib.check(lib.eml_linear_fp32(_lib.ptr(flat.contiguous()), _lib.ptr(self._fc[1]), _lib.ptr(self.fc.bias), _lib.ptr(z), B,
                                       self.fc.out_features, flat.shape[1], st), "eml_linear_fp32(netE.fc)")
        return z


class SPADEGenerator(nn.Module):
    """opt needs: ngf, norm_G ('spectralspadesyncbatch3x3'), semantic_nc (3), num_upsampling_layers ('normal'), crop_size, aspect_ratio."""

    def __init__(self, opt, precision="bf16x3"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        if getattr(opt, "num_upsampling_layers", "normal") != "normal":
            raise ValueError("only num_upsampling_layers='normal' (the reference default) is implemented")
        self.opt = opt
        self.precision = precision
        self.use_cuda_graph = False      # replay the whole forward (hundreds of launches) as one CUDA graph per input shape
        self.autograd = False            # standalone use: set True to record the train-mode forward on a tape so that .backward() works
                                         # (gp_train.py); Pix2PixModel drives the tape itself and does not need this
        self._graphs = {}
        nf = opt.ngf
        self.sw = opt.crop_size // 32
        self.sh = round(self.sw / opt.aspect_ratio)
        self.head_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_0 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.G_middle_1 = SPADEResnetBlock(16 * nf, 16 * nf, opt)
        self.up_0 = SPADEResnetBlock(16 * nf, 8 * nf, opt)
        self.up_1 = SPADEResnetBlock(8 * nf, 4 * nf, opt)
        self.up_2 = SPADEResnetBlock(4 * nf, 2 * nf, opt)
        self.up_3 = SPADEResnetBlock(2 * nf, 1 * nf, opt)
        self.up = nn.Upsample(scale_factor=2)
        self.sphere_conv1 = SphereConv2D(nf, 3, stride=1)
        self.netE = ConvEncoder(opt)

    @_lib.on_tensor_device
    def forward(self, input, crop):
        # training mode = the reference's train-mode FORWARD (batch-statistic BatchNorm inside SPADE with running-stat update,
        # one spectral-norm power iteration per wrapped convolution); the output carries no autograd graph -- backward is not built
        _lib.require_cuda(input, crop)
        if self.autograd and self.training and torch.is_grad_enabled():
            from . import gp_train
            return gp_train.run_with_tape(lambda tape: (gp_train.generator(tape, self, input, crop, True),), list(self.parameters()))[0]
        with torch.no_grad():
            if self.use_cuda_graph and not self.training:
                from .graphs import graphed_call
                return graphed_call(self._graphs, self._graph_key(), self._run, (input, crop))
            return self._run(input, crop)

    def _graph_key(self):
        return (self.precision,) + tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))

    def _run(self, guide, crop):
        lib = _lib.load()
        prec = self.precision
        B = guide.shape[0]
        dev = guide.device
        st = _lib.stream_ptr()
        guide = guide.contiguous().float()
        gh, gw = guide.shape[2], guide.shape[3]
        segs = {}

        def seg(h, w):
            if (h, w) not in segs:
                s = torch.zeros(B, h, w, 4, dtype=torch.float32, device=dev)
                _lib.check(lib.eml_resize_nearest(_lib.ptr(guide), 0, gh, gw, _lib.ptr(s), 4, h, w, guide.shape[1], B, 1, st), "eml_resize_nearest(guide)")
                segs[(h, w)] = s
            return segs[(h, w)]

        def upsample(x, H, W, C):
            out = torch.empty(B, 2 * H, 2 * W, x.shape[-1], dtype=torch.float32, device=dev)
            _lib.check(lib.eml_resize_nearest(_lib.ptr(x), x.shape[-1], H, W, _lib.ptr(out), out.shape[-1], 2 * H, 2 * W, C, B, 0, st), "eml_resize_nearest(x2)")
            return out

        z = self.netE.encode(crop, prec)                                   # (B, 32 ngf) == (B, 16 ngf, 1, 2) NCHW
        C = 16 * self.opt.ngf
        H, W = self.sh, self.sw
        x = torch.empty(B, H, W, C, dtype=torch.float32, device=dev)
        _lib.check(lib.eml_resize_nearest(_lib.ptr(z), 0, 1, 2, _lib.ptr(x), C, H, W, C, B, 1, st), "eml_resize_nearest(latent)")
        for name in ("head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"):
            blk = getattr(self, name)
            x = blk.apply_nhwc(x, B, H, W, seg(H, W), prec)
            C = blk.fout
            if name not in ("G_middle_0", "up_3"):
                x = upsample(x, H, W, C)
                H, W = 2 * H, 2 * W
        pc = self.sphere_conv1.packed(prec)
        raw = _sphere_conv_raw(x, B, H, W, pc, 1, None, 2, prec)           # SphereConv(LeakyReLU(x))
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
        _lib.check(lib.eml_tanh_to_nchw(_lib.ptr(raw), raw.shape[-1], _lib.ptr(self.sphere_conv1.bias), _lib.ptr(out), B, H * W, 3, 25.0, st),
                   "eml_tanh_to_nchw")
        return out


# ===================================================================================== discriminator, losses, model wrapper (G6-G8)
def _nchw_to_nhwc(x, pitch):
    """(B,C,H,W) -> zero-padded NHWC (B,H,W,pitch) through the nearest-resize kernel at identity scale."""
    lib = _lib.load()
    B, C, H, W = x.shape
    out = torch.zeros(B, H, W, pitch, dtype=torch.float32, device=x.device)
    _lib.check(lib.eml_resize_nearest(_lib.ptr(x.contiguous().float()), 0, H, W, _lib.ptr(out), pitch, H, W, C, B, 1, _lib.stream_ptr()),
               "eml_resize_nearest(NCHW->NHWC)")
    return out


def _to_nchw(x, C):
    return x[..., :C].permute(0, 3, 1, 2).contiguous()


def _bias_act(raw, bias, act, M, C):
    out = torch.empty_like(raw) if raw.shape[-1] == C else torch.zeros_like(raw)
    _lib.check(_lib.load().eml_bias_act(_lib.ptr(raw), raw.shape[-1], _lib.ptr(bias), act, _lib.ptr(out), out.shape[-1], M, C, _lib.stream_ptr()),
               "eml_bias_act")
    return out


def _pool(x, B, H, W, C, mode):
    ho, wo = ((H + 1) // 2, (W + 1) // 2) if mode == 0 else (H // 2, W // 2)
    out = torch.zeros(B, ho, wo, x.shape[-1], dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().eml_pool2d(_lib.ptr(x), x.shape[-1], H, W, _lib.ptr(out), out.shape[-1], C, B, mode, _lib.stream_ptr()), "eml_pool2d")
    return out, ho, wo


_RED_SUM, _RED_HINGE_REAL, _RED_HINGE_FAKE, _RED_L1, _RED_L1_MASKED, _RED_COS = range(6)


def _reduce(acc, mode, a, M, C, a_pitch, b=None, b_pitch=0, mask=None):
    _lib.check(_lib.load().eml_loss_reduce(_lib.ptr(a), a_pitch, _lib.ptr(b), b_pitch, _lib.ptr(mask), M, C, mode, _lib.ptr(acc), _lib.stream_ptr()),
               "eml_loss_reduce(mode %d)" % mode)


def _reduced_mean(mode, a, M, C, a_pitch, count, b=None, b_pitch=0, mask=None, sign=1.0):
    acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    _reduce(acc, mode, a, M, C, a_pitch, b, b_pitch, mask)
    return (acc * (sign / count)).float().reshape(())


class NLayerDiscriminator(nn.Module):
    """Drop-in for discriminator.NLayerDiscriminator (discriminator.py:69-125): SphereConv PatchGAN, ``model0..model{n}`` groups with
    the reference's parameter names.  opt needs: ndf, n_layers_D, norm_D ('spectralinstance'), label_nc, output_nc, no_ganFeat_loss."""

    def __init__(self, opt, precision="bf16x3"):
        super().__init__()
        if getattr(opt, "norm_D", "spectralinstance") != "spectralinstance":
            raise ValueError("only norm_D='spectralinstance' (the reference default) is implemented")
        self.opt = opt
        self.precision = precision
        nf = opt.ndf
        self.input_nc = opt.label_nc + opt.output_nc
        self.n_layers = opt.n_layers_D
        self.model0 = nn.Sequential(SphereConv2D(self.input_nc, nf, stride=2), nn.LeakyReLU(0.2, False))
        self.strides = [2]
        for n in range(1, opt.n_layers_D):
            nf_prev, nf = nf, min(nf * 2, 512)
            stride = 1 if n == opt.n_layers_D - 1 else 2
            self.strides.append(stride)
            setattr(self, "model%d" % n, nn.Sequential(nn.Sequential(_SpectralSphereConv2D(nf_prev, nf, stride=stride, bias=False),
                                                                     nn.InstanceNorm2d(nf, affine=False)), nn.LeakyReLU(0.2, False)))
        setattr(self, "model%d" % opt.n_layers_D, nn.Sequential(SphereConv2D(nf, 3, stride=1)))

    def features_nhwc(self, x, B, H, W):
        """x NHWC (B,H,W,pitch>=up4(input_nc)) -> [(NHWC tensor, h, w, C)] * (n_layers+1)."""
        lib = _lib.load()
        prec = self.precision
        st = _lib.stream_ptr()
        outs = []
        conv = self.model0[0]
        raw = _sphere_conv_raw(x, B, H, W, conv.packed(prec), 2, None, 0, prec)
        H, W = raw.shape[1], raw.shape[2]
        x = _bias_act(raw, conv.bias, 2, B * H * W, conv.out_c)
        outs.append((x, H, W, conv.out_c))
        for n in range(1, self.n_layers):
            conv = getattr(self, "model%d" % n)[0][0]
            raw = _sphere_conv_raw(x, B, H, W, conv.packed(prec), conv.stride, None, 0, prec)
            H, W = raw.shape[1], raw.shape[2]
            x = torch.empty_like(raw)
            _lib.check(lib.eml_instance_norm(_lib.ptr(raw), raw.shape[-1], _lib.ptr(x), x.shape[-1], B, H * W, conv.out_c, 1e-5, 1, st), "eml_instance_norm")
            outs.append((x, H, W, conv.out_c))
        conv = getattr(self, "model%d" % self.n_layers)[0]
        raw = _sphere_conv_raw(x, B, H, W, conv.packed(prec), 1, None, 0, prec)
        outs.append((_bias_act(raw, conv.bias, 0, B * H * W, 3), H, W, 3))
        return outs

    @_lib.on_tensor_device
    @torch.no_grad()
    def forward(self, input):
        _lib.require_cuda(input)
        B, C, H, W = input.shape
        outs = [_to_nchw(t, c) for t, _, _, c in self.features_nhwc(_nchw_to_nhwc(input, _up4(C)), B, H, W)]
        return outs if not self.opt.no_ganFeat_loss else outs[-1]


class MultiscaleDiscriminator(nn.Module):
    """Drop-in for discriminator.MultiscaleDiscriminator (discriminator.py:16-65): ``discriminator_{i}`` on the input avg-pooled i times.
    Forward only (stored spectral-norm vectors, as in eval mode); returns the reference's list of lists of NCHW tensors."""

    def __init__(self, opt, precision="bf16x3"):
        super().__init__()
        if getattr(opt, "netD_subarch", "n_layer") != "n_layer":
            raise ValueError("unrecognized discriminator subarchitecture %s" % opt.netD_subarch)
        self.opt = opt
        for i in range(opt.num_D):
            self.add_module("discriminator_%d" % i, NLayerDiscriminator(opt, precision))

    @property
    def precision(self):
        return self.discriminator_0.precision

    @precision.setter
    def precision(self, p):
        for d in self.children():
            d.precision = p

    def features_nhwc(self, x, B, H, W):
        result = []
        C = self.discriminator_0.input_nc
        for D in self.children():
            result.append(D.features_nhwc(x, B, H, W))
            x, H, W = _pool(x, B, H, W, C, 0)
        return result

    @_lib.on_tensor_device
    @torch.no_grad()
    def forward(self, input):
        _lib.require_cuda(input)
        B, C, H, W = input.shape
        feats = self.features_nhwc(_nchw_to_nhwc(input, _up4(C)), B, H, W)
        result = [[_to_nchw(t, c) for t, _, _, c in fl] for fl in feats]
        return result if not self.opt.no_ganFeat_loss else [[r[-1]] for r in result]


class GANLoss(nn.Module):
    """Drop-in for loss.GANLoss (loss.py:15-98), hinge mode (the reference default, train_options.py): scalar per call, the mean over
    the multiscale list.  Values only -- there is no autograd graph behind the returned tensors."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor, opt=None):
        super().__init__()
        if gan_mode not in ("ls", "original", "w", "hinge"):
            raise ValueError("Unexpected gan_mode {}".format(gan_mode))
        if gan_mode != "hinge":
            raise NotImplementedError("emlight_b200.GANLoss: only gan_mode='hinge' (the reference default) is implemented")
        self.gan_mode, self.opt = gan_mode, opt

    @_lib.on_tensor_device
    @torch.no_grad()
    def loss(self, input, target_is_real, for_discriminator=True):
        _lib.require_cuda(input)
        x = input.contiguous().float()
        if for_discriminator:
            mode = _RED_HINGE_REAL if target_is_real else _RED_HINGE_FAKE
        else:
            assert target_is_real, "The generator's hinge loss must be aiming for real"
            mode = _RED_SUM
        return _reduced_mean(mode, x, x.numel(), 1, 1, x.numel(), sign=-1.0)

    def __call__(self, input, target_is_real, for_discriminator=True):
        if isinstance(input, list):
            loss = 0
            for pred_i in input:
                if isinstance(pred_i, list):
                    pred_i = pred_i[-1]
                loss = loss + self.loss(pred_i, target_is_real, for_discriminator)
            return loss / len(input)
        return self.loss(input, target_is_real, for_discriminator)


_VGG_SLICES = ((0, "R"), (2, "R", "P", 5, "R"), (7, "R", "P", 10, "R"), (12, "R", 14, "R", 16, "R", "P", 19, "R"),
               (21, "R", 23, "R", 25, "R", "P", 28, "R"))
_VGG_CH = {0: (3, 64), 2: (64, 64), 5: (64, 128), 7: (128, 128), 10: (128, 256), 12: (256, 256), 14: (256, 256), 16: (256, 256),
           19: (256, 512), 21: (512, 512), 23: (512, 512), 25: (512, 512), 28: (512, 512)}


class VGG19(nn.Module):
    """Drop-in for architecture.VGG19 (architecture.py:92-122): torchvision vgg19.features[0:30] cut into ``slice1..slice5`` with the
    torchvision layer indices as module names, so a reference / torchvision state_dict loads.  There is no network here for the
    ImageNet checkpoint: weights are whatever is loaded (He-initialised by default).  forward(X NCHW) -> 5 relu feature maps."""

    def __init__(self, requires_grad=False, precision="bf16x3"):
        super().__init__()
        self.precision = precision
        for s, ops in enumerate(_VGG_SLICES):
            seq = nn.Sequential()
            idx = None
            for op in ops:
                if op == "R":
                    seq.add_module(str(idx + 1), nn.ReLU(inplace=True))
                elif op == "P":
                    seq.add_module(str(idx + 2), nn.MaxPool2d(2, 2))
                    idx += 1
                else:
                    idx = op
                    seq.add_module(str(idx), nn.Conv2d(*_VGG_CH[idx], kernel_size=3, padding=1))
            setattr(self, "slice%d" % (s + 1), seq)
        for prm in self.parameters():
            prm.requires_grad = requires_grad
        self._pc = {}

    def _packed(self, conv, precision):
        w = conv.weight
        key = (w.data_ptr(), w._version, precision, str(w.device))
        hit = self._pc.get(id(conv))
        if hit is None or hit[0] != key:
            with torch.no_grad():
                hit = (key, _PackedConv(w, precision))
            self._pc[id(conv)] = hit
        return hit[1]

    def features_nhwc(self, x, B, H, W):
        """x NHWC (B,H,W,4) -> [(NHWC tensor, h, w, C)] * 5."""
        prec = self.precision
        outs = []
        C = 3
        for s in range(5):
            for m in getattr(self, "slice%d" % (s + 1)):
                if isinstance(m, nn.Conv2d):
                    raw = _conv_raw(x, B, H, W, self._packed(m, prec), _LUTS.get("conv", H, W, 1, x.device), None, 0, prec)
                    C = m.out_channels
                    x = _bias_act(raw, m.bias, 1, B * H * W, C)           # conv bias + the ReLU that follows it
                elif isinstance(m, nn.MaxPool2d):
                    x, H, W = _pool(x, B, H, W, C, 1)
            outs.append((x, H, W, C))
        return outs

    @_lib.on_tensor_device
    @torch.no_grad()
    def forward(self, X):
        _lib.require_cuda(X)
        B, _, H, W = X.shape
        return [_to_nchw(t, c) for t, _, _, c in self.features_nhwc(_nchw_to_nhwc(X, 4), B, H, W)]


def _find_vgg19_weights():
    """Where an ImageNet VGG19 checkpoint may live on an offline box: $EML_VGG19_WEIGHTS, then torch hub's cache (what
    torchvision.models.vgg19(pretrained=True) -- architecture.py:95 -- would have downloaded)."""
    import glob
    cand = [os.environ.get("EML_VGG19_WEIGHTS")]
    hub = os.path.join(os.environ.get("TORCH_HOME", os.path.join(os.path.expanduser("~"), ".cache", "torch")), "hub", "checkpoints")
    cand += sorted(glob.glob(os.path.join(hub, "vgg19-*.pth")))
    for c in cand:
        if c and os.path.exists(c):
            return c
    return None


class VGGLoss(nn.Module):
    """Drop-in for loss.VGGLoss (loss.py:102-114): sum_k w_k * L1(vgg_k(x), vgg_k(y)), w = 1/32, 1/16, 1/8, 1/4, 1.  x and y run as
    one 2B batch.  Value only here; the training tape (gp_train.py) differentiates it.

    Weights: the reference builds ``torchvision.models.vgg19(pretrained=True)`` (architecture.py:95).  ``weights`` = a path to that
    checkpoint (torchvision's ``vgg19-dcbb9e9d.pth`` layout, ``features.N.weight``) or a state_dict; default = $EML_VGG19_WEIGHTS or
    torch hub's cache.  When none is found the features stay randomly initialised -- fine for timing and parity tests, WRONG for real
    training (the perceptual term would be measured in a random network's feature space) -- so that case warns loudly, and raises
    when ``require_pretrained`` (or $EML_REQUIRE_VGG=1) is set."""

    def __init__(self, gpu_ids=None, precision="bf16x3", weights=None, require_pretrained=None):
        super().__init__()
        self.vgg = VGG19(precision=precision).cuda()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
        if require_pretrained is None:
            require_pretrained = os.environ.get("EML_REQUIRE_VGG") == "1"
        src = weights if weights is not None else _find_vgg19_weights()
        self.pretrained = False
        if src is not None:
            sd = torch.load(src, map_location="cpu") if isinstance(src, (str, bytes, os.PathLike)) else src
            self.load_torchvision_state_dict(sd)
        elif require_pretrained:
            raise RuntimeError("emlight_b200.VGGLoss: no ImageNet VGG19 checkpoint found (set EML_VGG19_WEIGHTS=/path/to/vgg19-dcbb9e9d.pth "
                               "or pass weights=...); refusing to train against randomly initialised features")
        else:
            import warnings
            warnings.warn("emlight_b200.VGGLoss: no ImageNet VGG19 checkpoint found (EML_VGG19_WEIGHTS / torch hub cache); the perceptual "
                          "loss uses RANDOMLY INITIALISED features -- acceptable for benchmarks and parity tests only, not for training "
                          "(reference: torchvision.models.vgg19(pretrained=True), architecture.py:95)", RuntimeWarning, stacklevel=2)

    def load_torchvision_state_dict(self, sd):
        """Load torchvision's vgg19 ``features.N.{weight,bias}`` (or the reference VGG19's ``sliceK.N.*``) into slice1..slice5."""
        own = self.vgg.state_dict()
        by_idx = {k.split(".", 1)[1]: k for k in own}                       # "N.weight" -> "sliceK.N.weight"
        new, used = {}, 0
        for k, v in sd.items():
            if k.startswith("features.") and k[len("features."):] in by_idx:
                new[by_idx[k[len("features."):]]] = v
                used += 1
            elif k in own:
                new[k] = v
                used += 1
        if used != len(own):
            raise RuntimeError("VGG19 checkpoint does not cover features[0:30]: matched %d of %d tensors" % (used, len(own)))
        self.vgg.load_state_dict(new)
        self.pretrained = True

    @_lib.on_tensor_device
    @torch.no_grad()
    def forward(self, x, y):
        _lib.require_cuda(x, y)
        B, _, H, W = x.shape
        feats = self.vgg.features_nhwc(_nchw_to_nhwc(torch.cat([x, y], 0), 4), 2 * B, H, W)
        loss = 0
        for wk, (t, h, w, c) in zip(self.weights, feats):
            M = B * h * w
            loss = loss + wk * _reduced_mean(_RED_L1, t[:B], M, c, t.shape[-1], M * c, b=t[B:], b_pitch=t.shape[-1])
        return loss


@torch.no_grad()
def feature_matching_loss(feats, B, mask):
    """pix2pix_model.py:101-117 on the NHWC feature lists of a [fake; real] 2B batch: L1 between the mask-re-weighted fake and real
    features (weight m + 50 (1-m)); the mask is nearest-resized from its previous size at every layer, like the reference (:111)."""
    lib = _lib.load()
    num_D = len(feats)
    m = mask.contiguous().float()                                         # (B,1,H,W) == NHWC with one channel
    mh, mw = m.shape[2], m.shape[3]
    loss = 0
    for fl in feats:
        for t, h, w, c in fl[:-1]:
            nm = torch.empty(B, h, w, 1, dtype=torch.float32, device=m.device)
            _lib.check(lib.eml_resize_nearest(_lib.ptr(m), 1, mh, mw, _lib.ptr(nm), 1, h, w, 1, B, 0, _lib.stream_ptr()), "eml_resize_nearest(mask)")
            m, mh, mw = nm, h, w
            M = B * h * w
            loss = loss + _reduced_mean(_RED_L1_MASKED, t[:B], M, c, t.shape[-1], M * c, b=t[B:], b_pitch=t.shape[-1], mask=m) / num_D
    return loss.reshape(1)


@torch.no_grad()
def cosine_loss(fake, real):
    """(1 - CosineSimilarity(dim=1, eps=1e-20)(fake, real)).mean()  (pix2pix_model.py:95,122)."""
    B, C, H, W = fake.shape
    a, b = _nchw_to_nhwc(fake, _up4(C)), _nchw_to_nhwc(real, _up4(C))
    return _reduced_mean(_RED_COS, a, B * H * W, C, a.shape[-1], B * H * W, b=b, b_pitch=b.shape[-1])


class Pix2PixModel(nn.Module):
    """Drop-in for pix2pix_model.Pix2PixModel (pix2pix_model.py:12-186): mode dispatch 'inference' / 'generator' / 'discriminator'.
    'inference' is the full product path.  'generator' and 'discriminator' return the reference's loss dictionaries (same keys and
    weights).  When the model was built for training (`opt.isTrain`) and gradients are enabled, each call returns tensors attached to one
    autograd node (emlight_b200/gp_train.py), so `sum(losses.values()).mean().backward()` + `create_optimizers()` train the networks
    like GenProjector/model_trainer.py does (`self.autograd`; verified on B200 against torch autograd through the CPU restatement of the reference,
    tests/test_gp_train_gpu.py); under `torch.no_grad()` -- or with `self.autograd = False` -- the same calls return plain values."""

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.autograd = bool(getattr(opt, "isTrain", False))   # loss dictionaries carry an autograd node (gp_train.py): the trainer's .backward() works
        self.netG, self.netD = self.initialize_networks(opt)
        if opt.isTrain:
            self.criterionGAN = GANLoss(opt.gan_mode, opt=opt)
            if not getattr(opt, "no_vgg_loss", False):
                self.criterionVGG = VGGLoss(getattr(opt, "gpu_ids", None))

    def initialize_networks(self, opt):
        """(netG, netD) -- pix2pix_model.py:80-88; the module-name shim overrides this with the reference's define_G / define_D flow."""
        return SPADEGenerator(opt).cuda().eval(), (MultiscaleDiscriminator(opt).cuda().eval() if opt.isTrain else None)

    def forward(self, data, mode):
        input, crop, real_image, map = data["input"].cuda(), data["crop"].cuda(), data["warped"].cuda(), data["map"].cuda()
        if mode == "generator":
            return self.compute_generator_loss(input, crop, real_image, map)
        if mode == "discriminator":
            return self.compute_discriminator_loss(input, crop, real_image)
        if mode == "inference":
            return self.generate_fake(input, crop)
        raise ValueError("|mode| is invalid")

    def create_optimizers(self, opt):
        G_params = list(self.netG.parameters())
        D_params = list(self.netD.parameters()) if opt.isTrain else []
        G_lr, D_lr = (opt.lr, opt.lr) if opt.no_TTUR else (opt.lr / 2, opt.lr * 2)
        return (torch.optim.Adam(G_params, lr=G_lr, betas=(opt.beta1, opt.beta2)),
                torch.optim.Adam(D_params, lr=D_lr, betas=(opt.beta1, opt.beta2)))

    def generate_fake(self, input, crop):
        return self.netG(input, crop)

    @torch.no_grad()
    def _discriminate_nhwc(self, input, fake_image, real_image):
        both = torch.cat([torch.cat([input, fake_image], 1), torch.cat([input, real_image], 1)], 0)
        B2, C, H, W = both.shape
        return self.netD.features_nhwc(_nchw_to_nhwc(both, _up4(C)), B2, H, W)

    def discriminate(self, input, fake_image, real_image):
        feats = self._discriminate_nhwc(input, fake_image, real_image)
        B = input.shape[0]
        return ([[_to_nchw(t[:B], c) for t, _, _, c in fl] for fl in feats], [[_to_nchw(t[B:], c) for t, _, _, c in fl] for fl in feats])

    def compute_generator_loss(self, input, crop, real_image, map):
        if self.autograd and torch.is_grad_enabled():
            from . import gp_train

            def runner(tape):
                fake = gp_train.generator(tape, self.netG, input, crop, self.netG.training)
                return tuple(gp_train.generator_losses(tape, self, fake, input, real_image, map)) + (fake,)

            outs = gp_train.run_with_tape(runner, list(self.netG.parameters()))
            keys = ["GAN"] + ([] if self.opt.no_ganFeat_loss else ["GAN_Feat"]) + ["VGG", "COS"]
            return dict(zip(keys, outs[:-1])), outs[-1]
        with torch.no_grad():
            return self._generator_loss_values(input, crop, real_image, map)

    def _generator_loss_values(self, input, crop, real_image, map):
        fake_image = self.generate_fake(input, crop)
        B = input.shape[0]
        feats = self._discriminate_nhwc(input, fake_image, real_image)
        G_losses = {"GAN": self.criterionGAN([_to_nchw(fl[-1][0][:B], 3) for fl in feats], True, for_discriminator=False)}
        if not self.opt.no_ganFeat_loss:
            G_losses["GAN_Feat"] = feature_matching_loss(feats, B, map)
        G_losses["VGG"] = self.criterionVGG(fake_image, real_image) * 5
        G_losses["COS"] = cosine_loss(fake_image, real_image) * 5
        return G_losses, fake_image

    def compute_discriminator_loss(self, input, crop, real_image):
        if self.autograd and torch.is_grad_enabled():
            from . import gp_train
            with torch.no_grad():
                fake = self.generate_fake(input, crop).detach()
            outs = gp_train.run_with_tape(lambda tape: tuple(gp_train.discriminator_losses(tape, self, fake, input, real_image)),
                                          list(self.netD.parameters()))
            return {"D_Fake": outs[0], "D_real": outs[1]}
        with torch.no_grad():
            return self._discriminator_loss_values(input, crop, real_image)

    def _discriminator_loss_values(self, input, crop, real_image):
        fake_image = self.generate_fake(input, crop)
        B = input.shape[0]
        feats = self._discriminate_nhwc(input, fake_image, real_image)
        return {"D_Fake": self.criterionGAN([_to_nchw(fl[-1][0][:B], 3) for fl in feats], False, for_discriminator=True),
                "D_real": self.criterionGAN([_to_nchw(fl[-1][0][B:], 3) for fl in feats], True, for_discriminator=True)}

    def use_gpu(self):
        return True
