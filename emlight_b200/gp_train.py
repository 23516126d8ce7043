"""GenProjector TRAINING: forward with a tape and the matching backward for the generator, the discriminator, VGG and the loss terms.

What `GenProjector/trainers/model_trainer.py` runs per iteration is `Pix2PixModel(data, mode='generator')` -> sum of the loss dict ->
`.backward()` -> `optimizer_G.step()`, then the same with `mode='discriminator'` (pix2pix_model.py:92-141).  PyTorch autograd cannot
see through the C ABI, so the forward here records every primitive it launches on a tape and one `torch.autograd.Function` node per
call (`run_with_tape`) replays the tape in reverse when autograd reaches it; parameter gradients come back as the node's input grads,
so `loss.backward()`, `optimizer.step()` and the gradient all-reduce of `parallel.py` work unchanged.

Split of work in this first version:
* all contractions -- forward convolutions, data gradients `dA = dY Wk`, weight gradients `dWk^T = A^T dY` (K = pixels, split-K) --
  run on the TMA-fed tcgen05 GEMM through `gp_ops.mm_nt` / `gp_ops.conv_raw`; the im2col operand of the weight gradient is
  recomputed instead of stored, written directly TRANSPOSED and split into bf16 hi/lo (`eml_im2col_lut_bf16_t`), so neither an
  fp32 im2col matrix nor a transpose pass exists;
* the gather's adjoint (col2im through the sampling table), activation / bias, SPADE modulation, parameter-free BatchNorm and
  InstanceNorm adjoints are the kernels of `csrc/gp_bwd.cu` (`gp_ops.col2im`, `act_bwd`, `bias_act_bwd`, `spade_bwd`, `bn_free_bwd`,
  `instance_norm_bwd`);
* nearest-upsample, tanh, the two pooling adjoints and the loss seeds (hinge / L1 / masked L1 / cosine; the upstream scalar gradient
  is read on the device) are kernels of the same file (`upsample2_bwd`, `tanh_nchw_bwd`, `pool2d_bwd`, `loss_seed`);
* what is left in torch tensor code is bookkeeping on the same buffers: batch slicing, NCHW <-> NHWC permutes at the image
  boundary, the latent's 2 -> 8 column expansion, the spectral-norm quotient on the weight matrices and operand transposes.
Semantics follow the reference's modules in `.train()`: batch-statistic (Sync)BatchNorm inside SPADE with the running-stat update and
one all-reduce of the per-channel sums when several processes share the batch (forward AND backward), one spectral-norm power
iteration per wrapped convolution per forward with `u`, `v` treated as constants in the backward (torch.nn.utils.spectral_norm).

STATUS: the algebra is checked on CPU against torch autograd of the reference restatement (`tests/test_gp_train_cpu.py`: torch
stand-ins for the forward primitives, the adjoint kernels through their host-emulation build) and on B200 with the real kernels
(`tests/test_gp_train_gpu.py`: every generator / discriminator parameter gradient against torch autograd through the CPU restatement of the reference; with fp32 forward
contractions the bf16x3 backward agrees to 5e-4 worst / 1.5e-5 median per tensor).  `Pix2PixModel` built with `opt.isTrain` uses it by
default; a standalone `SPADEGenerator` / `SphereConv2D` opts in with `.autograd = True`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import gp_ops as ops
from .genprojector import _RED_COS, _RED_HINGE_FAKE, _RED_HINGE_REAL, _RED_L1, _RED_L1_MASKED, _RED_SUM, _sn_kernel_ok, _sn_sigma_kernel, _up4



# ============================================================================================================== tape
class Tape:
    """Reverse-mode bookkeeping: `steps` are closures run last-to-first; gradients are keyed by the identity of the forward tensor."""

    def __init__(self):
        self.steps = []
        self._grads = {}
        self._keep = []            # keeps every tensor that was used as a key alive (ids must not be recycled)
        self.param_grads = {}

    def record(self, fn):
        self.steps.append(fn)

    def add(self, t, g):
        key = id(t)
        if key in self._grads:
            self._grads[key] = self._grads[key] + g
        else:
            self._grads[key] = g
            self._keep.append(t)

    def take(self, t):
        return self._grads.pop(id(t), None)

    def add_param(self, p, g):
        if p is None or not p.requires_grad:
            return
        g = g.reshape(p.shape).to(p.dtype)
        self.param_grads[p] = self.param_grads[p] + g if p in self.param_grads else g

    def backward(self):
        for fn in reversed(self.steps):
            fn()
        self.steps = []
        self._grads = {}
        self._keep = []


class _TapeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, runner, *params):
        tape = Tape()
        outs = runner(tape)
        ctx.tape, ctx.outs, ctx.params = tape, outs, params
        return tuple(o.clone() for o in outs)

    @staticmethod
    def backward(ctx, *grads):
        tape = ctx.tape
        if tape is None:
            raise RuntimeError("emlight_b200: this GenProjector graph was already back-propagated (the tape is freed after one backward)")
        for o, g in zip(ctx.outs, grads):
            if g is not None:
                tape.add(o, g.to(o.dtype).expand_as(o))
        tape.backward()
        ctx.tape = None
        return (None,) + tuple(tape.param_grads.get(p) for p in ctx.params)


def run_with_tape(runner, params):
    """outs = runner(tape) as ONE autograd node whose inputs are `params` (the tensors that receive gradients)."""
    params = [p for p in params if p.requires_grad]
    return _TapeFn.apply(runner, *params)


# ============================================================================================================== helpers
def _pad_c(t, pitch):
    """(..., C) -> (..., pitch) zero padded."""
    C = t.shape[-1]
    return t if C == pitch else F.pad(t, (0, pitch - C))


def spectral_weight(module, update):
    """(W_eff, sigma, u, v) of a spectral-normalised conv: ONE power iteration when `update` (training forward; buffers written in place),
    the stored vectors otherwise -- torch.nn.utils.spectral_norm as used at architecture.py:37-40 / normalization.py:29."""
    w = module.weight_orig.detach()
    wm = w.reshape(w.shape[0], -1)
    with torch.no_grad():
        if _sn_kernel_ok(wm):
            sigma = _sn_sigma_kernel(module, wm, update)
            return w / sigma, sigma, module.weight_u.clone(), module.weight_v.clone()
        if update:
            v = F.normalize(torch.mv(wm.t(), module.weight_u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
            module.weight_v.copy_(v)
            module.weight_u.copy_(u)
        u, v = module.weight_u.clone(), module.weight_v.clone()
        sigma = torch.dot(u, torch.mv(wm, v))
        return w / sigma, sigma, u, v


def _spectral_backward(tape, module, w_eff, sigma, u, v, dw_eff):
    """d/dW_orig of W_orig / (u^T W_orig v) with u, v constant: (dW_eff - <dW_eff, W_eff> u v^T) / sigma."""
    inner = (dw_eff * w_eff).sum()
    uv = torch.outer(u, v).reshape(w_eff.shape)
    tape.add_param(module.weight_orig, (dw_eff - inner * uv) / sigma)


# ============================================================================================================== primitives with adjoints
# Diagnostics only (tests/debug_gp_bwd.py): {"fwd": p, "bwd": p} runs the forward / backward contractions at another precision than the
# module's, to separate activation-mask flips caused by forward rounding from the accuracy of the backward GEMMs themselves.
PRECISION_OVERRIDE = {}


def conv(tape, x, B, H, W, C, w_eff, lut, bias_in, bias_in_param, act, precision, on_dw, need_dx=True):
    """raw (B,ho,wo,up4(O)) = Wk * S(act(x[..., :C] + bias_in)) -- SphereConv2D / 3x3 conv without its own bias (the consumer adds it).
    `on_dw(dW (O,C,3,3))` receives the weight gradient (None: not needed); `bias_in_param` is the parameter behind `bias_in`."""
    fprec = PRECISION_OVERRIDE.get("fwd", precision)
    pc = ops.PackedConv(w_eff, fprec)
    raw = ops.conv_raw(x, B, H, W, pc, lut, bias_in, act, fprec)
    idx, wgt, ho, wo = lut
    O, Cp = pc.O, pc.Cp
    precision = PRECISION_OVERRIDE.get("bwd", precision)

    def bwd():
        g = tape.take(raw)
        if g is None:
            return
        M = B * ho * wo
        g2 = g[..., :O].reshape(M, O).contiguous()
        if on_dw is not None:
            if precision == "fp32":
                A = ops.im2col(x, B, H, W, C, lut, bias_in, act)                              # (M, 9 Cp), recomputed
                dwk_t = ops.mm_nt(A.t().contiguous(), g2.t().contiguous(), precision)        # (9 Cp, O), K = pixels
                del A
            else:                                                                             # operand recomputed straight into A^T bf16 hi/lo
                at_hi, at_lo = ops.im2col_t(x, B, H, W, C, lut, bias_in, act, precision == "bf16x3")
                dwk_t = ops.mm_nt_split(at_hi, at_lo, 9 * Cp, M, g2.t().contiguous(), precision)
                del at_hi, at_lo
            on_dw(dwk_t.reshape(9, Cp, O)[:, :C, :].permute(2, 1, 0).reshape(O, C, 3, 3))
        need_b = bias_in_param is not None and bias_in_param.requires_grad
        if not (need_dx or need_b):
            return
        wk_t = pc.wk[:, :9 * Cp].t().contiguous()                                             # (9 Cp, O)
        dA = ops.mm_nt(g2, wk_t, precision)                                                   # (M, 9 Cp)
        dx = ops.col2im(dA, Cp, lut, B, H * W).reshape(B, H, W, Cp)                          # adjoint of the 4-tap gather
        del dA
        sums = ops.act_bwd(dx, x, bias_in, act, B * H * W, C, need_b)                         # dx *= act'(x + bias_in), sum per channel
        if need_b:
            tape.add_param(bias_in_param, sums.float())
        if need_dx:
            tape.add(x, dx if x.shape[-1] == Cp else _pad_c(dx[..., :C], x.shape[-1]))

    tape.record(bwd)
    return raw, ho, wo


def bias_act(tape, raw, bias, act, B, H, W, C):
    out = ops.bias_act(raw, bias, act, B * H * W, C)

    def bwd():
        g = tape.take(out)
        if g is None:
            return
        need_b = bias is not None and bias.requires_grad
        dx, sums = ops.bias_act_bwd(g, out, act, B * H * W, C, need_b)
        if need_b:
            tape.add_param(bias, sums.float())
        tape.add(raw, dx)

    tape.record(bwd)
    return out


def instance_norm(tape, raw, B, H, W, C, lrelu=True, eps=1e-5):
    out = ops.instance_norm(raw, B, H * W, C, lrelu)

    def bwd():
        g = tape.take(out)
        if g is None:
            return
        tape.add(raw, ops.instance_norm_bwd(g, out, raw, B, H * W, C, lrelu, eps))

    tape.record(bwd)
    return out


def _world():
    d = torch.distributed
    return d.get_world_size() if d.is_available() and d.is_initialized() else 1


def spade(tape, mod, x, B, H, W, seg, x_bias, lrelu, precision, training):
    """SPADE on a raw conv output x whose bias `x_bias` (a Parameter or None) has not been added (genprojector.SPADE.apply_nhwc)."""
    C = mod.norm_nc
    bn = mod.param_free_norm
    M = B * H * W
    n = float(M)
    if training:
        sums = ops.channel_sums(x, M, C)
        if _world() > 1:
            from .parallel import global_count
            torch.distributed.all_reduce(sums)
            n = float(global_count(B, x.device) * H * W)            # true global count: shards may differ by one sample
        m_raw = sums[0] / n
        var = (sums[1] / n - m_raw * m_raw).clamp_min(0.0)
        mean = m_raw.float()
        inv = torch.rsqrt(var.float() + bn.eps)
        with torch.no_grad():
            m_full = mean if x_bias is None else mean + x_bias.detach()
            bn.running_mean.mul_(1 - bn.momentum).add_(bn.momentum * m_full)
            bn.running_var.mul_(1 - bn.momentum).add_(bn.momentum * (var * (n / max(n - 1.0, 1.0))).float())
    else:
        mean = bn.running_mean if x_bias is None else bn.running_mean - x_bias.detach()
        inv = torch.rsqrt(bn.running_var + bn.eps)
    shared, gamma, beta = mod.mlp_shared[0], mod.mlp_gamma, mod.mlp_beta
    dev = x.device
    sl = ops.lut("sphere", H, W, 1, dev)

    def dw_shared(dw):
        tape.add_param(shared.weight, dw)

    def dw_gb(dw):
        tape.add_param(gamma.weight, dw[:C])
        tape.add_param(beta.weight, dw[C:])

    actv, _, _ = conv(tape, seg, B, H, W, mod.label_nc, shared.weight.detach(), sl, None, None, 0, precision, dw_shared, need_dx=False)
    gb, _, _ = conv(tape, actv, B, H, W, shared.out_c, torch.cat([gamma.weight.detach(), beta.weight.detach()], 0), sl,
                    shared.bias.detach(), shared.bias, 1, precision, dw_gb, need_dx=True)
    out = ops.spade_modulate(x, mean, inv, gb, gamma.bias.detach(), beta.bias.detach(), M, C, lrelu)

    def bwd():
        g = tape.take(out)
        if g is None:
            return
        d_gb, d_xhat, sums = ops.spade_bwd(g, out, x, mean, inv, gb, gamma.bias.detach(), M, C, lrelu)
        tape.add(gb, d_gb)
        tape.add_param(gamma.bias, sums[0].float())
        tape.add_param(beta.bias, sums[1].float())
        if training:
            s2 = sums[2:4].clone()
            if _world() > 1:
                torch.distributed.all_reduce(s2)
            dx = ops.bn_free_bwd(d_xhat, x, mean, inv, s2, n, M, C)
            if x_bias is not None:                                  # a shift in front of a batch-statistic BatchNorm cancels exactly
                tape.add_param(x_bias, torch.zeros_like(x_bias))
        else:
            dx = ops.bn_free_bwd(d_xhat, None, None, inv, None, 0.0, M, C)
            if x_bias is not None:
                tape.add_param(x_bias, (inv.double() * sums[2]).float())
        tape.add(x, dx)

    tape.record(bwd)
    return out


def bias_residual(tape, a, bias_a, r, bias_r, B, H, W, C):
    out = ops.bias_residual(a, bias_a.detach() if bias_a is not None else None, r, bias_r.detach() if bias_r is not None else None,
                            B * H * W, C)

    def bwd():
        g = tape.take(out)
        if g is None:
            return
        gc = g[..., :C]
        s = ops.channel_sums(g.contiguous(), B * H * W, C)[0].float() if (bias_a is not None or bias_r is not None) else None
        if bias_a is not None:
            tape.add_param(bias_a, s)
        if bias_r is not None:
            tape.add_param(bias_r, s)
        tape.add(a, _pad_c(gc, a.shape[-1]))
        if r is not None:
            tape.add(r, _pad_c(gc, r.shape[-1]))

    tape.record(bwd)
    return out


def upsample2(tape, x, B, H, W, C):
    out = ops.resize_nearest(x, x.shape[-1], H, W, 2 * H, 2 * W, C, B, 0, x.shape[-1])

    def bwd():
        g = tape.take(out)
        if g is not None:
            tape.add(x, ops.upsample2_bwd(g, B, H, W, C))

    tape.record(bwd)
    return out


# ============================================================================================================== generator
def _sn_conv(tape, mod, x, B, H, W, C, lut, precision, training, need_dx=True, bias_in=None, bias_in_param=None, act=0):
    """A (possibly spectral-normalised) SphereConv2D / encoder conv without its own bias."""
    if hasattr(mod, "weight_orig"):
        w_eff, sigma, u, v = spectral_weight(mod, training)

        def on_dw(dw):
            _spectral_backward(tape, mod, w_eff, sigma, u, v, dw)
    else:
        w_eff = mod.weight.detach()

        def on_dw(dw):
            tape.add_param(mod.weight, dw)
    return conv(tape, x, B, H, W, C, w_eff, lut, bias_in, bias_in_param, act, precision, on_dw, need_dx)


def resnet_block(tape, blk, x, B, H, W, seg, precision, training):
    dev = x.device
    sl = ops.lut("sphere", H, W, 1, dev)
    r, r_bias = x, None
    if blk.learned_shortcut:
        s = spade(tape, blk.norm_s, x, B, H, W, seg, None, False, precision, training)
        r, _, _ = _sn_conv(tape, blk.conv_s, s, B, H, W, blk.fin, sl, precision, training)
        r_bias = blk.conv_s.bias
    h = spade(tape, blk.norm_0, x, B, H, W, seg, None, True, precision, training)
    d0, _, _ = _sn_conv(tape, blk.conv_0, h, B, H, W, blk.fin, sl, precision, training)
    h = spade(tape, blk.norm_1, d0, B, H, W, seg, blk.conv_0.bias, True, precision, training)
    d1, _, _ = _sn_conv(tape, blk.conv_1, h, B, H, W, blk.fmiddle, sl, precision, training)
    return bias_residual(tape, d1, blk.conv_1.bias, r, r_bias, B, H, W, blk.fout)


def encoder(tape, enc, crop, precision, training):
    B = crop.shape[0]
    dev = crop.device
    x = ops.resize_bilinear_nchw(crop, 128, 128)
    H = W = 128
    C = 3
    for i in range(1, 6):
        cv = getattr(enc, "layer%d" % i)[0]
        raw, ho, wo = _sn_conv(tape, cv, x, B, H, W, C, ops.lut("conv", H, W, 2, dev), precision, training, need_dx=i > 1)
        H, W, C = ho, wo, cv.out_channels
        x = instance_norm(tape, raw, B, H, W, C, lrelu=True)
    fc = enc.fc
    # fc consumes the NCHW flattening (generator.py:124); the activations are NHWC -> permute the weight columns
    wf = fc.weight.detach().float().view(fc.out_features, C, H * W).permute(0, 2, 1).reshape(fc.out_features, -1).contiguous()
    flat = x[..., :C].reshape(B, -1).contiguous()
    z = ops.linear(flat, wf, fc.bias.detach())

    def bwd():
        g = tape.take(z)
        if g is None:
            return
        tape.add_param(fc.bias, g.sum(0))
        dwf = ops.mm_nt(g.t().contiguous(), flat.t().contiguous(), precision)              # (out, H*W*C) in NHWC column order
        tape.add_param(fc.weight, dwf.reshape(fc.out_features, H * W, C).permute(0, 2, 1).reshape(fc.out_features, -1))
        dflat = ops.mm_nt(g, wf.t().contiguous(), precision)                                # (B, H*W*C)
        tape.add(x, _pad_c(dflat.reshape(B, H, W, C), x.shape[-1]))

    tape.record(bwd)
    return z


def generator(tape, G, guide, crop, training=True):
    """SPADEGenerator.forward (generator.py:65-88) on the tape: returns fake (B,3,128,256) NCHW in [0,50]."""
    prec = G.precision
    B = guide.shape[0]
    guide = guide.contiguous().float()
    gh, gw = guide.shape[2], guide.shape[3]
    segs = {}

    def seg(h, w):
        if (h, w) not in segs:
            segs[(h, w)] = ops.resize_nearest(guide, 0, gh, gw, h, w, guide.shape[1], B, 1, 4)
        return segs[(h, w)]

    z = encoder(tape, G.netE, crop, prec, training)
    C = 16 * G.opt.ngf
    H, W = G.sh, G.sw
    x = ops.resize_nearest(z, 0, 1, 2, H, W, C, B, 1, C)

    def bwd_latent(x=x, H=H, W=W, C=C):
        g = tape.take(x)
        if g is not None:                                                                   # z viewed (B,C,1,2): column j feeds w in [j W/2, (j+1) W/2)
            gz = g.reshape(B, H, 2, W // 2, C).sum((1, 3))                                  # (B, 2, C)
            tape.add(z, gz.permute(0, 2, 1).reshape(B, 2 * C))

    tape.record(bwd_latent)
    for name in ("head_0", "G_middle_0", "G_middle_1", "up_0", "up_1", "up_2", "up_3"):
        blk = getattr(G, name)
        x = resnet_block(tape, blk, x, B, H, W, seg(H, W), prec, training)
        C = blk.fout
        if name not in ("G_middle_0", "up_3"):
            x = upsample2(tape, x, B, H, W, C)
            H, W = 2 * H, 2 * W
    last = G.sphere_conv1
    raw, _, _ = _sn_conv(tape, last, x, B, H, W, C, ops.lut("sphere", H, W, 1, x.device), prec, training, act=2)
    out = ops.tanh_to_nchw(raw, last.bias.detach(), B, H, W, 3, 25.0)

    def bwd_out():
        g = tape.take(out)
        if g is None:
            return
        d_raw, sums = ops.tanh_nchw_bwd(g, out, 25.0, B, H * W, 3, raw.shape[-1], last.bias.requires_grad)
        if sums is not None:
            tape.add_param(last.bias, sums.float())
        tape.add(raw, d_raw.reshape(raw.shape))

    tape.record(bwd_out)
    return out


# ============================================================================================================== discriminator / VGG
def _pool_with_grad(tape, x, B, H, W, C, mode):
    out, ho, wo = ops.pool(x, B, H, W, C, mode)

    def bwd():
        g = tape.take(out)
        if g is None:
            return
        tape.add(x, ops.pool2d_bwd(g, x, B, H, W, C, mode))

    tape.record(bwd)
    return out, ho, wo


def nlayer_discriminator(tape, D, x, B, H, W, training, need_dx, want_dw):
    """NLayerDiscriminator.features_nhwc on the tape (discriminator.py:113-123)."""
    prec = D.precision
    dev = x.device
    outs = []

    def sn(mod, x, H, W, C, stride, ndx):
        lut = ops.lut("sphere", H, W, stride, dev)
        if not want_dw:
            w_eff = spectral_weight(mod, training)[0] if hasattr(mod, "weight_orig") else mod.weight.detach()
            return conv(tape, x, B, H, W, C, w_eff, lut, None, None, 0, prec, None, ndx)
        return _sn_conv(tape, mod, x, B, H, W, C, lut, prec, training, need_dx=ndx)

    cv = D.model0[0]
    raw, H, W = sn(cv, x, H, W, D.input_nc, 2, need_dx)
    x = bias_act(tape, raw, cv.bias if want_dw else cv.bias.detach(), 2, B, H, W, cv.out_c)
    outs.append((x, H, W, cv.out_c))
    for n in range(1, D.n_layers):
        C = cv.out_c
        cv = getattr(D, "model%d" % n)[0][0]
        raw, H, W = sn(cv, x, H, W, C, cv.stride, True)
        x = instance_norm(tape, raw, B, H, W, cv.out_c, lrelu=True)
        outs.append((x, H, W, cv.out_c))
    C = cv.out_c
    cv = getattr(D, "model%d" % D.n_layers)[0]
    raw, H, W = sn(cv, x, H, W, C, 1, True)
    outs.append((bias_act(tape, raw, cv.bias if want_dw else cv.bias.detach(), 0, B, H, W, 3), H, W, 3))
    return outs


def multiscale_discriminator(tape, netD, x, B, H, W, training, need_dx, want_dw):
    result = []
    C = netD.discriminator_0.input_nc
    kids = list(netD.children())
    for i, D in enumerate(kids):
        result.append(nlayer_discriminator(tape, D, x, B, H, W, training, need_dx, want_dw))
        if i + 1 < len(kids):
            x, H, W = _pool_with_grad(tape, x, B, H, W, C, 0) if need_dx else ops.pool(x, B, H, W, C, 0)
    return result


def vgg_features(tape, vgg, x, B, H, W):
    """VGG19.features_nhwc on the tape: data gradients only (the loss network is frozen, architecture.py:118-120)."""
    prec = vgg.precision
    outs = []
    C = 3
    for s in range(5):
        for m in getattr(vgg, "slice%d" % (s + 1)):
            if isinstance(m, nn.Conv2d):
                raw, _, _ = conv(tape, x, B, H, W, C, m.weight.detach(), ops.lut("conv", H, W, 1, x.device), None, None, 0, prec, None, True)
                C = m.out_channels
                x = bias_act(tape, raw, m.bias.detach(), 1, B, H, W, C)
            elif isinstance(m, nn.MaxPool2d):
                x, H, W = _pool_with_grad(tape, x, B, H, W, C, 1)
        outs.append((x, H, W, C))
    return outs


# ============================================================================================================== losses
def _mean_loss(tape, mode, a, M, C, count, b=None, mask=None, sign=1.0, scale=1.0):
    """scalar = scale * sign / count * eml_loss_reduce(mode); its adjoint seeds d/da through eml_loss_seed with the upstream scalar
    gradient read on the device (no host synchronisation)."""
    coef = sign * scale / count
    val = (ops.loss_sum(mode, a, M, C, a.shape[-1], b, b.shape[-1] if b is not None else 0, mask) * coef).float().reshape(())

    def bwd():
        g = tape.take(val)
        if g is None:
            return
        tape.add(a, ops.loss_seed(mode, a, M, C, coef, g.reshape(1).float().contiguous(), b, mask))

    tape.record(bwd)
    return val


def _sum(tape, vals, shape=()):
    """Sum of taped scalars as a new taped tensor of `shape` (the loss dictionaries' entries)."""
    out = torch.stack([v.reshape(()) for v in vals]).sum().reshape(shape)

    def bwd():
        g = tape.take(out)
        if g is not None:
            for v in vals:
                tape.add(v, g.reshape(v.shape))

    tape.record(bwd)
    return out


def _slice_rows(tape, t, lo, hi):
    """t[lo:hi] along the batch as its own tape tensor (contiguous view of the NHWC buffer)."""
    v = t[lo:hi]

    def bwd():
        g = tape.take(v)
        if g is not None:
            full = torch.zeros_like(t)
            full[lo:hi] = g
            tape.add(t, full)

    tape.record(bwd)
    return v


def generator_losses(tape, model, fake, guide, real, mask):
    """compute_generator_loss (pix2pix_model.py:92-128) given the taped fake image: [GAN, GAN_Feat, VGG, COS] scalars."""
    opt = model.opt
    netD = model.netD
    training = netD.training
    B, _, H, W = fake.shape
    guide, real = guide.contiguous().float(), real.contiguous().float()
    both = torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0)
    C6 = both.shape[1]
    gc = guide.shape[1]
    x = ops.nchw_to_nhwc(both, _up4(C6))

    def bwd_cat():
        g = tape.take(x)
        if g is not None:
            tape.add(fake, g[:B, :, :, gc:C6].permute(0, 3, 1, 2))

    tape.record(bwd_cat)
    feats = multiscale_discriminator(tape, netD, x, 2 * B, H, W, training, need_dx=True, want_dw=False)
    num_D = len(feats)
    losses = []
    gan = []
    for fl in feats:                                                     # GANLoss hinge, generator side: -mean(D(fake)) per scale
        t, h, w, c = fl[-1]
        tf = _slice_rows(tape, t, 0, B)
        n = B * h * w * c
        gan.append(_mean_loss(tape, _RED_SUM, tf, B * h * w, c, n, sign=-1.0, scale=1.0 / num_D))
    losses.append(_sum(tape, gan))
    if not opt.no_ganFeat_loss:
        m = mask.contiguous().float()
        mh, mw = m.shape[2], m.shape[3]
        fm = []
        for fl in feats:
            for t, h, w, c in fl[:-1]:
                m = ops.resize_nearest(m, 1, mh, mw, h, w, 1, B, 0, 1)   # re-interpolated from its PREVIOUS size (pix2pix_model.py:111)
                mh, mw = h, w
                tf = _slice_rows(tape, t, 0, B)
                fm.append(_mean_loss(tape, _RED_L1_MASKED, tf, B * h * w, c, B * h * w * c, b=t[B:], mask=m, scale=1.0 / num_D))
        losses.append(_sum(tape, fm, (1,)))
    vin = torch.cat([fake, real], 0)
    xv = ops.nchw_to_nhwc(vin, 4)

    def bwd_vin():
        g = tape.take(xv)
        if g is not None:
            tape.add(fake, g[:B, :, :, :3].permute(0, 3, 1, 2))

    tape.record(bwd_vin)
    vf = vgg_features(tape, model.criterionVGG.vgg, xv, 2 * B, H, W)
    vl = []
    for wk, (t, h, w, c) in zip(model.criterionVGG.weights, vf):
        tf = _slice_rows(tape, t, 0, B)
        vl.append(_mean_loss(tape, _RED_L1, tf, B * h * w, c, B * h * w * c, b=t[B:], scale=5.0 * wk))
    losses.append(_sum(tape, vl))
    a = ops.nchw_to_nhwc(fake, 4)
    bq = ops.nchw_to_nhwc(real, 4)

    def bwd_a():
        g = tape.take(a)
        if g is not None:
            tape.add(fake, g[..., :3].permute(0, 3, 1, 2))

    tape.record(bwd_a)

    losses.append(_mean_loss(tape, _RED_COS, a, B * H * W, 3, B * H * W, b=bq, scale=5.0))
    return losses


def discriminator_losses(tape, model, fake, guide, real):
    """compute_discriminator_loss (pix2pix_model.py:130-141) with a detached fake image: [D_Fake, D_real]."""
    netD = model.netD
    B, _, H, W = fake.shape
    guide, real = guide.contiguous().float(), real.contiguous().float()
    both = torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0)
    x = ops.nchw_to_nhwc(both, _up4(both.shape[1]))
    feats = multiscale_discriminator(tape, netD, x, 2 * B, H, W, netD.training, need_dx=False, want_dw=True)
    num_D = len(feats)
    d_fake, d_real = [], []
    for fl in feats:
        t, h, w, c = fl[-1]
        n = B * h * w * c
        tf, tr = _slice_rows(tape, t, 0, B), _slice_rows(tape, t, B, 2 * B)
        # hinge: D_Fake = -mean(min(-x-1, 0)) -> d/dx = [x > -1] / n ; D_real = -mean(min(x-1, 0)) -> d/dx = -[x < 1] / n
        d_fake.append(_mean_loss(tape, _RED_HINGE_FAKE, tf, B * h * w, c, n, sign=-1.0, scale=1.0 / num_D))
        d_real.append(_mean_loss(tape, _RED_HINGE_REAL, tr, B * h * w, c, n, sign=-1.0, scale=1.0 / num_D))
    return [_sum(tape, d_fake), _sum(tape, d_real)]


def sphere_conv_module(mod, x):
    """`SphereConv2D.forward` (sphere_cnn.py:111-124) with autograd: NCHW in / NCHW out, gradients for weight, bias and -- when it
    requires them -- the input.  One tape node per call (opt-in: `SphereConv2D.autograd = True`)."""
    B, C, H, W = x.shape
    xd = x.detach().float()
    box = {}

    def runner(tape):
        xn = ops.nchw_to_nhwc(xd, _up4(C))
        box["xn"], box["tape"] = xn, tape
        lut = ops.lut("sphere", H, W, mod.stride, xd.device)
        raw, ho, wo = _sn_conv(tape, mod, xn, B, H, W, C, lut, mod.precision, mod.training, need_dx=x.requires_grad)
        out = bias_act(tape, raw, mod.bias, 0, B, ho, wo, mod.out_c) if mod.bias is not None else raw
        y = out[..., :mod.out_c].permute(0, 3, 1, 2).contiguous()

        def bwd():
            g = tape.take(y)
            if g is not None:
                tape.add(out, _pad_c(g.permute(0, 2, 3, 1), out.shape[-1]).contiguous())

        tape.record(bwd)
        if x.requires_grad:                                       # runs last in the reverse sweep: hand the input gradient to autograd
            def bwd_x():
                g = tape.take(xn)
                if g is not None:
                    tape.param_grads[x] = g[..., :C].permute(0, 3, 1, 2).contiguous().to(x.dtype)
            tape.steps.insert(0, bwd_x)
        return (y,)

    params = [p for p in mod.parameters()] + ([x] if x.requires_grad else [])
    return run_with_tape(runner, params)[0]
