"""EMLight regression network -- drop-in for ``RegressionNetwork/DenseNet.py`` running on sm_100a kernels.

Same constructor signature, same module / parameter / buffer names (``features.denseblock{b}.denselayer{l}.
{norm1,conv1,norm2,conv2}``, ``features.transition{b}.{norm,conv}``, ``features.last_norm{b}``, ``fc*``) and the same
default initialisation order as DenseNet.py:82-133, so ``state_dict()`` / ``load_state_dict()`` interoperate with
reference checkpoints and ``torch.manual_seed(s); DenseNet()`` yields the reference's initial weights.  ``forward(x)``
returns the dict of DenseNet.py:153-157.  The nn.Conv2d / nn.BatchNorm2d / nn.Linear children only *own* the
parameters; forward never calls them -- it drives the C ABI of include/emlight_b200.h:

  stem   conv0+norm0+relu        -> eml_stem_forward                      (NCHW image -> NHWC concat buffer)
  layer  norm1+relu+conv1        -> eml_conv_forward(1x1, relu=1)         reads the concat buffer in place
         norm2+conv2 (NO relu,   -> eml_conv_forward(3x3, relu=0)         writes its 12 channels in place
         DenseNet.py:41-43)
  trans  norm+relu+conv+avgpool  -> eml_conv_forward(POOL2)               pooling commuted in front of the 1x1 conv
  last_norm{b}                   -> folded into every consumer's scale/shift (affine o affine), never materialised
  head   relu+avgpool4+fc+4 fc   -> eml_head_pool, eml_linear_fp32 x2     (fc columns permuted NCHW->NHWC at pack time)

BatchNorm follows ``self.training`` like the reference: batch statistics (accumulated by the producing kernels'
epilogues, folded by eml_bn_fold) + running-stat update in train mode -- which is what the reference's test.py
actually runs -- running statistics in eval mode.

Backward (training-mode BatchNorm, the mode train.py runs) is `_backward`: both data-gradient convolutions are the same
tcgen05 implicit-GEMM kernels with transposed / flipped weights, BatchNorm backward and the weight gradients are the kernels of
csrc/bwd_ops.cu and csrc/dense_bwd1.cu (fused conv1 backward), the 48-channel bottlenecks are kept by the training forward when memory
allows (else recomputed), the tiny fc / head GEMMs use torch.matmul (cuBLAS).
"""
import math
import os
from collections import OrderedDict
from ctypes import c_void_p

import torch
import torch.nn as nn

from . import _lib
from ._lib import ConvParams, DenseLayerParams

_EPS = 1e-5


def _up4(n):
    return (n + 3) & ~3


def _up8(n):
    """Slab pixel pitch in floats: records that are a multiple of 32 bytes keep every pixel's channels on the same sector grid,
    which lets the dense-layer kernel write its 12 new channels as two FULL 32-byte sectors (csrc/dense_layer.cu, wide store)."""
    return (n + 7) & ~7


class _Transition(nn.Sequential):
    def __init__(self, c_in, c_out):
        super().__init__()
        self.add_module("norm", nn.BatchNorm2d(c_in))
        self.add_module("relu", nn.ReLU(inplace=True))
        self.add_module("conv", nn.Conv2d(c_in, c_out, kernel_size=1, stride=1, bias=False))
        self.add_module("pool", nn.AvgPool2d(kernel_size=2, stride=2))


class _DenseLayer(nn.Sequential):
    def __init__(self, c_in, growth_rate, bn_size, drop_rate):
        super().__init__()
        if bn_size <= 0:
            raise ValueError("emlight_b200.DenseNet implements the bottleneck (bn_size > 0) variant the reference uses")
        if drop_rate:
            raise ValueError("drop_rate != 0 is not implemented (the reference trains with drop_rate=0)")
        inter = 4 * growth_rate                       # DenseNet.py:37: hard-wired, bn_size only gates it
        self.add_module("norm1", nn.BatchNorm2d(c_in))
        self.add_module("relu1", nn.ReLU(inplace=True))
        self.add_module("conv1", nn.Conv2d(c_in, inter, kernel_size=1, stride=1, bias=False))
        self.add_module("norm2", nn.BatchNorm2d(inter))
        self.add_module("conv2", nn.Conv2d(inter, growth_rate, kernel_size=3, padding=1, bias=False))
        self.drop_rate = drop_rate


class _DenseBlock(nn.Sequential):
    def __init__(self, num_layers, c_in, bn_size, growth_rate, drop_rate):
        super().__init__()
        for i in range(num_layers):
            self.add_module("denselayer%d" % (i + 1), _DenseLayer(c_in + i * growth_rate, growth_rate, bn_size, drop_rate))


class _DenseNetFn(torch.autograd.Function):
    """Whole-network autograd node: forward runs the kernel sequence, backward runs `DenseNet._backward` (training-mode BN only)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        outs = module._run(x)
        ctx.module = module
        ctx.fwd_id = module._fwd_id
        ctx.n_params = len(params)
        ctx.save_for_backward(x)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        module = ctx.module
        if not module._last_train:
            raise NotImplementedError("emlight_b200.DenseNet.backward is implemented for training-mode BatchNorm (module.train()), "
                                      "the mode RegressionNetwork/train.py runs; eval-mode backward is not implemented")
        if ctx.fwd_id != module._fwd_id:
            raise RuntimeError("emlight_b200.DenseNet: backward() must follow its own forward() (activations are kept in a single "
                               "workspace and were overwritten by a later forward)")
        (x,) = ctx.saved_tensors
        named = module._backward(x, grads)
        out = [named.get(n) for n, p in module.named_parameters() if p.requires_grad]
        assert len(out) == ctx.n_params
        return (None, None) + tuple(out)


class DenseNet(nn.Module):
    def __init__(self, growth_rate=12, block_config=(16, 16, 16), compression=0.5, num_init_features=24, bn_size=4,
                 drop_rate=0, avgpool_size=4, *, n_anchors=96, precision="bf16x3"):
        super().__init__()
        if precision not in _lib.PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(_lib.PRECISIONS))
        self.avgpool_size = avgpool_size
        self.precision = precision
        self.growth_rate = growth_rate
        self.block_config = tuple(block_config)
        self.features = nn.Sequential(OrderedDict([
            ("conv0", nn.Conv2d(3, num_init_features, kernel_size=3, stride=1, padding=1, bias=False)),
            ("norm0", nn.BatchNorm2d(num_init_features)),
            ("relu0", nn.ReLU(inplace=True)),
        ]))
        c = num_init_features
        self._plan = []                                # (block, c_in, c_out, c_tr)
        for i, n in enumerate(self.block_config):
            self.features.add_module("denseblock%d" % (i + 1), _DenseBlock(n, c, bn_size, growth_rate, drop_rate))
            c_out = c + n * growth_rate
            c_tr = int(math.floor(c_out * compression))
            # DenseNet.py:110: `i != len(block_config)` is always true -> a transition follows EVERY block
            self.features.add_module("transition%d" % (i + 1), _Transition(c_out, c_tr))
            self.features.add_module("last_norm%d" % (i + 1), nn.BatchNorm2d(c_tr))
            self._plan.append((i + 1, c, c_out, c_tr))
            c = c_tr
        self.fc = nn.Linear(8208, 1024)
        self.fc_dist = nn.Linear(1024, n_anchors)
        self.fc_intensity = nn.Linear(1024, 1)
        self.fc_rgb_ratio = nn.Linear(1024, 3)
        self.fc_ambient = nn.Linear(1024, 3)
        self.sigmoid = nn.Sigmoid()                    # parameter-free members the reference also registers
        self.tanh = nn.Tanh()
        self.softmax = nn.Softmax(dim=1)
        self.relu = nn.ReLU()
        self._cache = None                             # packed weights + folded eval-mode BN, keyed on param versions
        self._ws = {}                                  # activation workspaces keyed on (B,H,W,device)
        self.launch_log = None                         # set to a list to collect (kernel family, name, shape info, start, end events)
        self._fwd_id = 0                               # bumped by every forward; backward checks it still owns the workspace
        self._last_train = False
        self.use_cuda_graph = False                    # eval mode only: replay the 104-launch forward as one CUDA graph per input shape
        self._graphs = {}
        self.fuse_dense_layers = True                  # eval mode: one composite-filter kernel per dense layer (csrc/dense_layer.cu)
        self._grad_sink = None                         # parallel.FlatAdam.sink: receives finished gradients DURING the backward (per block)

    # ------------------------------------------------------------------ reference-facing API
    @_lib.on_tensor_device
    def forward(self, x):
        _lib.require_cuda(x)
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError("expected input (B,3,H,W), got %s" % (tuple(x.shape),))
        params = [p for p in self.parameters() if p.requires_grad]
        self._for_backward = bool(torch.is_grad_enabled() and (x.requires_grad or params))
        if self._for_backward:
            d, i, r, a = _DenseNetFn.apply(self, x, *params)
        elif self.use_cuda_graph and not self.training and self.launch_log is None:
            from .graphs import graphed_call
            d, i, r, a = graphed_call(self._graphs, self._state_key(x.device), lambda t: tuple(self._run(t)), (x,))
        else:
            d, i, r, a = self._run(x)
        return {"distribution": d, "intensity": i, "rgb_ratio": r, "ambient": a}

    # ------------------------------------------------------------------ packing
    def _bns(self):
        """[(name, module)] of every BatchNorm in execution order."""
        out = [("norm0", self.features.norm0)]
        for b, _, _, _ in self._plan:
            blk = getattr(self.features, "denseblock%d" % b)
            for l, layer in enumerate(blk.children()):
                out.append(("b%d.l%d.norm1" % (b, l), layer.norm1))
                out.append(("b%d.l%d.norm2" % (b, l), layer.norm2))
            out.append(("t%d.norm" % b, getattr(self.features, "transition%d" % b).norm))
            out.append(("ln%d" % b, getattr(self.features, "last_norm%d" % b)))
        return out

    def _convs(self):
        out = []
        for b, _, _, _ in self._plan:
            blk = getattr(self.features, "denseblock%d" % b)
            for l, layer in enumerate(blk.children()):
                out.append(("b%d.l%d.conv1" % (b, l), layer.conv1, 1))
                out.append(("b%d.l%d.conv2" % (b, l), layer.conv2, 9))
            out.append(("t%d.conv" % b, getattr(self.features, "transition%d" % b).conv, 1))
        return out

    def _state_key(self, device):
        vs = tuple((p.data_ptr(), p._version) for p in self.parameters())
        bs = tuple((b.data_ptr(), b._version) for b in self.buffers())
        return (str(device), self.precision, self.fuse_dense_layers, vs, bs if not self.training else None)

    @torch.no_grad()
    def _pack(self, device):
        """(Re)build the derived cache: packed bf16 hi/lo weight images, permuted fc weight, eval-mode BN folds."""
        lib = _lib.load()
        key = self._state_key(device)
        if self._cache is not None and self._cache["key"] == key:
            return self._cache
        st = _lib.stream_ptr()
        c = {"key": key, "wpack": {}, "w": {}}
        for name, conv, taps in self._convs():
            w = conv.weight.detach().contiguous().float()
            c["w"][name] = w
            if True:                                    # packed even in fp32 mode: the backward may run at another precision than the forward
                nbytes = lib.eml_conv_wpack_bytes(w.shape[0], w.shape[1], taps)      # (tests isolate forward-rounding effects that way)
                buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
                _lib.check(lib.eml_conv_pack_weights(_lib.ptr(w), _lib.ptr(buf), w.shape[0], w.shape[1], taps, st),
                           "eml_conv_pack_weights")
                c["wpack"][name] = buf
        c["w0"] = self.features.conv0.weight.detach().contiguous().float()
        # fc consumes the pooled map flattened in NCHW order (DenseNet.py:138); ours is NHWC -> permute its columns once
        c_tr = self._plan[-1][3]
        P = self.fc.in_features // c_tr
        c["fc_w"] = self.fc.weight.detach().float().view(-1, c_tr, P).permute(0, 2, 1).reshape(self.fc.out_features, -1).contiguous()
        c["fc_b"] = self.fc.bias.detach().float().contiguous()
        if self.precision != "fp32":
            # fc as a TMA-fed tcgen05 GEMM (eml_gemm_bf16): weights packed in slices of <= 256 output features.  (Narrower slices do not
            # help: the launches serialise on the stream and each is bound by its 8208-deep K loop -- 16 x 64 features took 1.25 ms
            # against 0.62 ms for 4 x 256, profiles/r02_profile_train_b64_wgrad3x3_v2.txt.)
            c["fc_pack"] = []
            K = self.fc.in_features
            step = int(os.environ.get("EML_FC_SLICE", "256"))
            nfc = self.fc.out_features
            if (nfc % step == 0 and nfc > step and step % 16 == 0 and os.environ.get("EML_FC_SERIAL") != "1" and
                    os.environ.get("EML_FC_SPLITK") != "1"):
                # all slices in one buffer at a fixed stride: ONE launch runs them side by side (eml_gemm_bf16_slices; same bits)
                sb = lib.eml_conv_wpack_bytes(step, K, 1)
                buf = torch.empty(sb * (nfc // step), dtype=torch.uint8, device=device)
                _lib.check(lib.eml_gemm_pack_slices(_lib.ptr(c["fc_w"]), _lib.ptr(buf), nfc // step, step, K, sb, st), "eml_gemm_pack_slices(fc)")
                c["fc_slices"] = (buf, sb, nfc // step, step)
            for n0 in range(0, nfc if "fc_slices" not in c else 0, step):
                rows = min(step, nfc - n0)
                buf = torch.empty(lib.eml_conv_wpack_bytes(rows, K, 1), dtype=torch.uint8, device=device)
                _lib.check(lib.eml_conv_pack_weights(_lib.ptr(c["fc_w"][n0:n0 + rows]), _lib.ptr(buf), rows, K, 1, st), "eml_conv_pack_weights(fc)")
                c["fc_pack"].append((n0, rows, buf))
        heads = (self.fc_dist, self.fc_intensity, self.fc_rgb_ratio, self.fc_ambient)
        c["head_w"] = torch.cat([h.weight.detach().float() for h in heads], 0).contiguous()
        c["head_b"] = torch.cat([h.bias.detach().float() for h in heads], 0).contiguous()
        c["head_split"] = [h.out_features for h in heads]
        # affine table: one [scale | shift] pair (each padded to 4 floats) per BatchNorm
        offs, total = {}, 0
        for name, bn in self._bns():
            cp = _up4(bn.num_features)
            offs[name] = (total, cp)
            total += 2 * cp
        c["aff_offs"] = offs
        c["aff"] = torch.zeros(total, dtype=torch.float32, device=device)
        # identity pre-affine rows (scale 1, shift 0) for channels that are stored post-activation
        cmax = _up4(max(p[2] for p in self._plan))
        c["pre"] = torch.zeros(len(self._plan), 2, cmax, dtype=torch.float32, device=device)
        c["pre"][:, 0, :] = 1.0
        if not self.training:
            self._fold_eval(c)
        self._cache = c
        return c

    def _aff(self, c, name):
        off, cp = c["aff_offs"][name]
        return c["aff"][off:off + cp], c["aff"][off + cp:off + 2 * cp]

    def _fold(self, c, name, bn, stats=None, stride=0, count=0.0, pre=None, mean_var=None):
        lib = _lib.load()
        sc, sh = self._aff(c, name)
        ps, pt = (None, None) if pre is None else pre
        mv_m, mv_v = (None, None) if mean_var is None else mean_var
        _lib.check(lib.eml_bn_fold(_lib.ptr(stats), stride, float(count),
                                   _lib.ptr(bn.running_mean), _lib.ptr(bn.running_var), _lib.ptr(bn.weight),
                                   _lib.ptr(bn.bias), _lib.ptr(ps), _lib.ptr(pt), _lib.ptr(sc), _lib.ptr(sh),
                                   _lib.ptr(mv_m), _lib.ptr(mv_v), bn.num_features, _EPS, _lib.stream_ptr()),
                   "eml_bn_fold(%s)" % name)

    def _fold_eval(self, c):
        """Eval mode: every BN's scale/shift from running statistics; last_norm{b} composed into block b+1's consumers."""
        f = self.features
        self._fold(c, "norm0", f.norm0)
        pre = None
        for bi, (b, c_in, c_out, c_tr) in enumerate(self._plan):
            blk = getattr(f, "denseblock%d" % b)
            for l, layer in enumerate(blk.children()):
                self._fold(c, "b%d.l%d.norm1" % (b, l), layer.norm1, pre=pre)
                self._fold(c, "b%d.l%d.norm2" % (b, l), layer.norm2)
            tr = getattr(f, "transition%d" % b)
            self._fold(c, "t%d.norm" % b, tr.norm, pre=pre)
            ln = getattr(f, "last_norm%d" % b)
            self._fold(c, "ln%d" % b, ln)
            if bi + 1 < len(self._plan):
                a, s = self._aff(c, "ln%d" % b)
                c["pre"][bi + 1, 0, :c_tr] = a[:c_tr]
                c["pre"][bi + 1, 1, :c_tr] = s[:c_tr]
                pre = (c["pre"][bi + 1, 0], c["pre"][bi + 1, 1])
        if self.fuse_dense_layers and self.precision != "fp32":
            self._compose_layers(c)

    def _compose_layers(self, c):
        """Eval mode: conv2 o norm2 o conv1 is linear (no ReLU between them, DenseNet.py:41-43) -> one 3x3 filter per layer,
        Weff[(dy,dx,o), ci] = sum_b W2[o,b,dy,dx] * scale2[b] * W1[b,ci], plus the bias table of the norm2 shift seen through
        the taps that fall inside the image (include/emlight_b200.h: eml_dense_layer_forward).  fp64 on the device, once per
        parameter version."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        g = self.growth_rate
        c["fused"] = {}
        dev = c["aff"].device
        nlayers = sum(self.block_config)
        bias_all = torch.empty(nlayers, 9 * g, dtype=torch.float32, device=dev)
        i = 0
        for b, c_in, _, _ in self._plan:
            blk = getattr(self.features, "denseblock%d" % b)
            for l, layer in enumerate(blk.children()):
                ci = c_in + l * g
                if not any(lib.eml_dense_layer_supported(2, wq, ci, g, _lib.PRECISIONS[self.precision]) for wq in (64, 128, 256)):
                    i += 1                              # no geometry fuses this layer (block 3's last one, C_in = 330: more resident
                    continue                            # weights than shared memory holds): it keeps the two-kernel path
                s2, t2 = self._aff(c, "b%d.l%d.norm2" % (b, l))
                nb = layer.conv1.out_channels
                buf = torch.empty(lib.eml_dense_layer_wpack_bytes(ci), dtype=torch.uint8, device=dev)
                # one launch per layer: fp64 composition written straight into the packed operand image (csrc/dense_layer.cu)
                _lib.check(lib.eml_dense_layer_compose(_lib.ptr(c["w"]["b%d.l%d.conv1" % (b, l)]), _lib.ptr(c["w"]["b%d.l%d.conv2" % (b, l)]),
                                                       _lib.ptr(s2), _lib.ptr(t2), nb, ci, g, _lib.ptr(buf), _lib.ptr(bias_all[i]), st),
                           "eml_dense_layer_compose(b%d.l%d)" % (b, l))
                c["fused"][(b, l)] = (buf, bias_all[i])
                i += 1

    def _dense_layer(self, c, key, slab, pitch, h, w, B, ci, plane_pixels=0):
        lib = _lib.load()
        buf, bias9 = c["fused"][key]
        sc, sh = self._aff(c, "b%d.l%d.norm1" % key)
        p = DenseLayerParams()
        p.in_ = slab.data_ptr(); p.scale = sc.data_ptr(); p.shift = sh.data_ptr()
        p.wpack = buf.data_ptr(); p.bias9 = bias9.data_ptr(); p.out = slab.data_ptr()
        p.B, p.H, p.W, p.C_in, p.in_pitch = B, h, w, ci, pitch
        p.growth, p.out_pitch, p.out_choff = self.growth_rate, pitch, ci
        p.precision = _lib.PRECISIONS[self.precision]
        p.plane_pixels = plane_pixels                      # > 0: `slab` is the channel-plane form (G, B*h*w, 32), see _planes_ok
        if self.launch_log is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(lib.eml_dense_layer_forward(p, _lib.stream_ptr()), "eml_dense_layer_forward(b%d.l%d)" % key)
        if self.launch_log is not None:
            e1.record()
            g, nb = self.growth_rate, 4 * self.growth_rate
            M = B * h * w
            # algorithmic bytes: the slab channels read once + the 12 new channels written once; flops of the reference layer
            self.launch_log.append(("dense_layer", "b%d.l%d" % key, 4 * M * (ci + g), 2 * M * (ci * nb + 9 * nb * g), e0, e1))

    def _planes_ok(self, c, ws, B):
        """Block 1 can keep its slab as channel planes when every one of its layers runs as the one-kernel dense layer and transition 1 on
        the TMA pipeline (eval mode, tensor-core precision, W in {128, 256}); EML_DENSE_PLANES=0 keeps the NHWC records."""
        if os.environ.get("EML_DENSE_PLANES") == "0" or self.precision == "fp32":
            return False
        lib = _lib.load()
        b, c_in, c_out, c_tr = self._plan[0]
        h, w = ws["geom"][0]
        prec = _lib.PRECISIONS[self.precision]
        nl = (c_out - c_in) // self.growth_rate
        if w not in (128, 256) or any((b, l) not in c.get("fused", ()) or
                                      not lib.eml_dense_layer_supported(h, w, c_in + l * self.growth_rate, self.growth_rate, prec) for l in range(nl)):
            return False
        return bool(lib.eml_transition_planes_supported(h, w, c_out, c_tr, prec)) and ("t%d.conv" % b) in c["wpack"]

    # ------------------------------------------------------------------ workspaces
    def _workspace(self, B, H, W, device):
        key = (B, H, W, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        if H % 32 or W % 32:
            raise ValueError("input height/width must be multiples of 32 (three /2 transitions and a /4 pool)")
        c_last = self._plan[-1][3]
        if c_last * (H // 32) * (W // 32) != self.fc.in_features:
            raise ValueError("input %dx%d gives %d pooled features but fc expects %d (the reference network is built "
                             "for 192x256 crops, SURVEY F2)" % (H, W, c_last * (H // 32) * (W // 32), self.fc.in_features))
        ws = {"slab": [], "geom": []}
        h, w = H, W
        for b, c_in, c_out, c_tr in self._plan:
            ws["slab"].append(torch.empty(B, h, w, _up8(c_out), dtype=torch.float32, device=device))
            if _up8(c_out) > c_out:
                ws["slab"][-1][..., c_out:].zero_()        # pitch padding: quad loads of the last channels reach it (masked, but defined)
            ws["geom"].append((h, w))
            h, w = h // 2, w // 2
        ws["bott"] = torch.empty(B, H, W, 4 * self.growth_rate, dtype=torch.float32, device=device)
        ws["t_last"] = torch.empty(B, h, w, _up4(c_last), dtype=torch.float32, device=device)
        ws["pooled"] = torch.empty(B, self.fc.in_features, dtype=torch.float32, device=device)
        ws["fc"] = torch.empty(B, self.fc.out_features, dtype=torch.float32, device=device)
        # batch-statistics accumulators (train mode): per slab (2, pitch) + per conv1 output (2, 48) + stem raw + last
        n = 2 * 32
        so = {"stem_raw": 0}
        for bi, (b, c_in, c_out, c_tr) in enumerate(self._plan):
            so["slab%d" % b] = n; n += 2 * _up8(c_out)
            for l in range(self.block_config[bi]):
                so["b%d.l%d.mid" % (b, l)] = n; n += 2 * 4 * self.growth_rate
        so["t_last"] = n; n += 2 * _up4(c_last)
        ws["stats"] = torch.zeros(n, dtype=torch.float64, device=device)
        ws["stats_offs"] = so
        nb = sum(bn.num_features for _, bn in self._bns())
        ws["bmean"] = torch.zeros(nb, dtype=torch.float32, device=device)
        ws["bvar"] = torch.zeros(nb, dtype=torch.float32, device=device)
        self._ws = {key: ws}                            # keep one geometry resident
        return ws

    # ------------------------------------------------------------------ execution
    def _conv(self, c, name, src, src_pitch, H, W, B, c_in, dst, dst_pitch, choff, c_out, mode, relu, aff, stats, stride, plane_pixels=0):
        lib = _lib.load()
        p = ConvParams()
        p.in_ = src.data_ptr()
        p.scale = aff[0].data_ptr()
        p.shift = aff[1].data_ptr()
        p.w_oihw = c["w"][name].data_ptr()
        p.wpack = c["wpack"][name].data_ptr() if name in c["wpack"] else None
        p.out = dst.data_ptr()
        p.stats = stats.data_ptr() if stats is not None else None
        p.stats_stride = stride
        p.B, p.H, p.W = B, H, W
        p.C_in, p.in_pitch = c_in, src_pitch
        p.C_out, p.out_pitch, p.out_choff = c_out, dst_pitch, choff
        p.mode, p.relu, p.precision = mode, relu, _lib.PRECISIONS[self.precision]
        p.plane_pixels = plane_pixels
        if self.launch_log is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        _lib.check(lib.eml_conv_forward(p, _lib.stream_ptr()), "eml_conv_forward(%s)" % name)
        if self.launch_log is not None:
            e1.record()
            m_out = B * H * W // (4 if mode == _lib.EML_CONV_POOL2 else 1)
            # algorithmic bytes: every input element read once, every output element written once (fp32)
            abytes = 4 * (B * H * W * c_in + m_out * c_out)
            flops = 2 * m_out * c_out * c_in * (9 if mode == _lib.EML_CONV_3x3 else 1)
            self.launch_log.append((("conv1x1", "conv3x3", "pool1x1")[mode], name, abytes, flops, e0, e1))

    @torch.no_grad()
    def _run(self, x):
        lib = _lib.load()
        dev = x.device
        B, _, H, W = x.shape
        x = x.contiguous().float()
        c = self._pack(dev)
        ws = self._workspace(B, H, W, dev)
        st = _lib.stream_ptr()
        train = self.training
        self._fwd_id += 1
        self._last_train = train
        f = self.features
        g = 4 * self.growth_rate
        so = ws["stats_offs"]
        stats = ws["stats"]
        bm_off = [0]
        updates = []                                    # (bn, mean_view, var_view, count) for the running-stat update

        def mv(bn, count):
            o = bm_off[0]
            bm_off[0] += bn.num_features
            m, v = ws["bmean"][o:o + bn.num_features], ws["bvar"][o:o + bn.num_features]
            updates.append((bn, m, v, count))
            return m, v

        if train:
            stats.zero_()
        # Training keeps every layer's 48-channel bottleneck (conv1 output) for the backward when that fits comfortably -- 4 * 48 B per
        # pixel and layer: 12.7 GB at B = 64 -- instead of recomputing it there (one more read of the layer's whole input slab per layer);
        # otherwise one buffer is reused and the backward recomputes.
        keep = None
        if train and getattr(self, "_for_backward", False):
            need = sum(B * hh * ww * g * 4 * n for (hh, ww), n in zip(ws["geom"], self.block_config))
            free, _total = torch.cuda.mem_get_info(dev)
            have = sum(t.numel() * 4 for t in ws.get("bott_keep", {}).values())
            if os.environ.get("EML_RECOMPUTE_BOTTLENECK") != "1" and need - have < 0.5 * free:
                keep = ws.setdefault("bott_keep", {})
        ws["bott_kept"] = keep is not None
        # ---- block 1 as channel planes (eval): (G, B*H*W, 32) instead of (B, H, W, pitch) -- §4.0 of DESIGN.md
        planes0 = 0
        if not train and self._planes_ok(c, ws, B):
            if "slab_planes0" not in ws:
                groups = (self._plan[0][2] + 31) // 32
                free, _total = torch.cuda.mem_get_info(dev)
                # a second copy of block 1's buffer (the NHWC one stays for training): only when memory is plentiful
                ws["slab_planes0"] = (torch.zeros(groups, B * H * W, 32, dtype=torch.float32, device=dev)     # zeroed: unwritten channels stay finite
                                      if groups * B * H * W * 128 < 0.5 * free else None)
            if ws["slab_planes0"] is not None:
                planes0 = B * H * W
        # ---- stem (DenseNet.py:89-92)
        slab = ws["slab_planes0"] if planes0 else ws["slab"][0]
        pitch = 32 if planes0 else slab.shape[3]
        s1 = stats[so["slab1"]:]
        if train:
            raw = stats[so["stem_raw"]:]
            _lib.check(lib.eml_stem_forward(_lib.ptr(x), _lib.ptr(c["w0"]), None, None, None, pitch, _lib.ptr(raw), None, 0,
                                            B, H, W, f.conv0.out_channels, 0, 1, st), "eml_stem_forward(stats)")
            self._fold(c, "norm0", f.norm0, stats=raw, stride=f.conv0.out_channels, count=B * H * W,
                       mean_var=mv(f.norm0, B * H * W))
        a0 = self._aff(c, "norm0")
        _lib.check(lib.eml_stem_forward(_lib.ptr(x), _lib.ptr(c["w0"]), _lib.ptr(a0[0]), _lib.ptr(a0[1]), _lib.ptr(slab), pitch,
                                        None, _lib.ptr(s1) if train else None, pitch, B, H, W, f.conv0.out_channels, 1, 1, st),
                   "eml_stem_forward")
        # ---- dense blocks + transitions
        pre = None
        for bi, (b, c_in, c_out, c_tr) in enumerate(self._plan):
            planes = planes0 if bi == 0 else 0
            slab = ws["slab_planes0"] if planes else ws["slab"][bi]
            pitch = 32 if planes else slab.shape[3]
            h, w = ws["geom"][bi]
            count = B * h * w
            sstat = stats[so["slab%d" % b]:]
            blk = getattr(f, "denseblock%d" % b)
            for l, layer in enumerate(blk.children()):
                ci = c_in + l * self.growth_rate
                n1, n2 = "b%d.l%d.norm1" % (b, l), "b%d.l%d.norm2" % (b, l)
                mid = stats[so["b%d.l%d.mid" % (b, l)]:]
                if (not train and (b, l) in c.get("fused", ()) and (w != 64 or B % 2 == 0) and    # W = 64 tiles hold a row of two images
                        lib.eml_dense_layer_supported(h, w, ci, self.growth_rate, _lib.PRECISIONS[self.precision])):
                    self._dense_layer(c, (b, l), slab, pitch, h, w, B, ci, planes)
                    continue
                assert not planes, "channel-plane slab: every layer of the block must take the one-kernel path (_planes_ok)"
                if train:
                    self._fold(c, n1, layer.norm1, stats=sstat, stride=pitch, count=count, pre=pre, mean_var=mv(layer.norm1, count))
                bott = ws["bott"]
                if keep is not None:
                    bott = keep.get((b, l))
                    if bott is None or bott.shape[0] != B:
                        bott = keep[(b, l)] = torch.empty(B, h, w, g, dtype=torch.float32, device=dev)
                self._conv(c, "b%d.l%d.conv1" % (b, l), slab, pitch, h, w, B, ci, bott, g, 0, g, _lib.EML_CONV_1x1, 1,
                           self._aff(c, n1), mid if train else None, g)
                if train:
                    self._fold(c, n2, layer.norm2, stats=mid, stride=g, count=count, mean_var=mv(layer.norm2, count))
                self._conv(c, "b%d.l%d.conv2" % (b, l), bott, g, h, w, B, g, slab, pitch, ci, self.growth_rate,
                           _lib.EML_CONV_3x3, 0, self._aff(c, n2), sstat[ci:] if train else None, pitch)
            tr = getattr(f, "transition%d" % b)
            tn = "t%d.norm" % b
            if train:
                self._fold(c, tn, tr.norm, stats=sstat, stride=pitch, count=count, pre=pre, mean_var=mv(tr.norm, count))
            last = bi + 1 == len(self._plan)
            dst = ws["t_last"] if last else ws["slab"][bi + 1]
            dpitch = dst.shape[3]
            dstat = stats[so["t_last"]:] if last else stats[so["slab%d" % (b + 1)]:]
            self._conv(c, "t%d.conv" % b, slab, pitch, h, w, B, c_out, dst, dpitch, 0, c_tr, _lib.EML_CONV_POOL2, 1,
                       self._aff(c, tn), dstat if train else None, dpitch, planes)
            ln = getattr(f, "last_norm%d" % b)
            if train:
                cnt2 = B * (h // 2) * (w // 2)
                self._fold(c, "ln%d" % b, ln, stats=dstat, stride=dpitch, count=cnt2, mean_var=mv(ln, cnt2))
                if not last:
                    a, s = self._aff(c, "ln%d" % b)
                    c["pre"][bi + 1, 0, :c_tr] = a[:c_tr]
                    c["pre"][bi + 1, 1, :c_tr] = s[:c_tr]
            if not last:
                pre = (c["pre"][bi + 1, 0], c["pre"][bi + 1, 1])
        # ---- head (DenseNet.py:136-150): last_norm affine + relu + avgpool -> fc -> 4 heads
        hl, wl = ws["t_last"].shape[1], ws["t_last"].shape[2]
        a, s = self._aff(c, "ln%d" % self._plan[-1][0])
        _lib.check(lib.eml_head_pool(_lib.ptr(ws["t_last"]), ws["t_last"].shape[3], _lib.ptr(a), _lib.ptr(s), _lib.ptr(ws["pooled"]),
                                     B, hl, wl, self._plan[-1][3], self.avgpool_size, st), "eml_head_pool")
        if "fc_pack" in c and B >= 32:
            # (B, 8208) x (8208, 1024): tensor cores (1.44 ms -> ~0.2 ms at B = 256); small batches keep the weight-streaming SIMT kernel
            K = self.fc.in_features
            Kp = (K + 63) // 64 * 64
            if "fc_a" not in ws:
                ws["fc_a"] = torch.empty(2, B, Kp, dtype=torch.bfloat16, device=dev)
            a_hi, a_lo = ws["fc_a"][0], (ws["fc_a"][1] if self.precision == "bf16x3" else None)
            _lib.check(lib.eml_split_bf16(_lib.ptr(ws["pooled"]), B, K, K, _lib.ptr(a_hi), _lib.ptr(a_lo), Kp, st), "eml_split_bf16(pooled)")
            splitk = os.environ.get("EML_FC_SPLITK") == "1"          # experiment (off by default): M = B rows give only ceil(B/128) row
            if splitk:                                               # tiles per 256-feature slice, i.e. 8 CTAs at B = 256; split K over the SMs
                ws["fc"].zero_()                                     # (float atomics: no longer bit-reproducible run to run)
                ks = max(1, min(Kp // 64, 148 // ((B + 127) // 128)))
            if "fc_slices" in c and not splitk:
                buf, sb, nsl, step = c["fc_slices"]
                _lib.check(lib.eml_gemm_bf16_slices(_lib.ptr(a_hi), _lib.ptr(a_lo), B, Kp, _lib.ptr(buf), sb, nsl, step, _lib.ptr(c["fc_b"]),
                                                    _lib.ptr(ws["fc"]), self.fc.out_features, 0, _lib.PRECISIONS[self.precision], 1, st),
                           "eml_gemm_bf16_slices(fc)")
            for n0, rows, buf in c["fc_pack"]:
                if splitk:
                    _lib.check(lib.eml_gemm_bf16_splitk(_lib.ptr(a_hi), _lib.ptr(a_lo), B, Kp, _lib.ptr(buf), rows, _lib.ptr(c["fc_b"][n0:n0 + rows]),
                                                        _lib.ptr(ws["fc"]), self.fc.out_features, n0, _lib.PRECISIONS[self.precision], ks, st),
                               "eml_gemm_bf16_splitk(fc)")
                    continue
                _lib.check(lib.eml_gemm_bf16(_lib.ptr(a_hi), _lib.ptr(a_lo), B, Kp, _lib.ptr(buf), rows, _lib.ptr(c["fc_b"][n0:n0 + rows]),
                                             _lib.ptr(ws["fc"]), self.fc.out_features, n0, _lib.PRECISIONS[self.precision], st), "eml_gemm_bf16(fc)")
        else:
            _lib.check(lib.eml_linear_fp32(_lib.ptr(ws["pooled"]), _lib.ptr(c["fc_w"]), _lib.ptr(c["fc_b"]), _lib.ptr(ws["fc"]),
                                           B, self.fc.out_features, self.fc.in_features, st), "eml_linear_fp32(fc)")
        nh = c["head_w"].shape[0]
        heads = torch.empty(B, nh, dtype=torch.float32, device=dev)
        _lib.check(lib.eml_linear_fp32(_lib.ptr(ws["fc"]), _lib.ptr(c["head_w"]), _lib.ptr(c["head_b"]), _lib.ptr(heads),
                                       B, nh, self.fc.out_features, st), "eml_linear_fp32(heads)")
        if train:
            self._update_running_stats(updates)
        outs, o = [], 0
        for n in c["head_split"]:
            outs.append(heads[:, o:o + n])
            o += n
        return outs

    @torch.no_grad()
    def _update_running_stats(self, updates):
        """running = (1-m)*running + m*batch, unbiased variance, num_batches_tracked += 1 (nn.BatchNorm2d semantics)."""
        for bn, mean, var, count in updates:
            if not bn.track_running_stats or bn.running_mean is None:
                continue
            bn.num_batches_tracked += 1
            m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - m).add_(mean, alpha=m)
            bn.running_var.mul_(1 - m).add_(var, alpha=m * count / max(count - 1, 1))

    # ------------------------------------------------------------------ backward (training-mode BN)
    def _gemm_bwd(self, src_ptr, src_pitch, B, H, W, c_in, w_oihw, dst, mode):
        """dst[..., :N] = conv(src, w) with no prologue affine, N sliced into <= 256 output channels (dgrad GEMMs)."""
        lib = _lib.load()
        st = _lib.stream_ptr()
        N = w_oihw.shape[0]
        taps = w_oihw.shape[2] * w_oihw.shape[3]
        for n0 in range(0, N, 256):
            n = min(256, N - n0)
            w_s = w_oihw[n0:n0 + n].contiguous()
            pack = None
            if self.precision != "fp32":
                pack = torch.empty(lib.eml_conv_wpack_bytes(n, c_in, taps), dtype=torch.uint8, device=dst.device)
                _lib.check(lib.eml_conv_pack_weights(_lib.ptr(w_s), _lib.ptr(pack), n, c_in, taps, st), "eml_conv_pack_weights")
            p = ConvParams()
            p.in_ = src_ptr; p.scale = None; p.shift = None
            p.w_oihw = w_s.data_ptr(); p.wpack = pack.data_ptr() if pack is not None else None
            p.out = dst.data_ptr(); p.stats = None; p.stats_stride = 0
            p.B, p.H, p.W = B, H, W
            p.C_in, p.in_pitch = c_in, src_pitch
            p.C_out, p.out_pitch, p.out_choff = n, dst.shape[-1], n0
            p.mode, p.relu, p.precision = mode, 0, _lib.PRECISIONS[self.precision]
            _lib.check(lib.eml_conv_forward(p, st), "eml_conv_forward(dgrad)")

    @torch.no_grad()
    def _backward(self, x, grads):
        lib = _lib.load()
        st = _lib.stream_ptr()
        dev = x.device
        B, _, H, W = x.shape
        x = x.contiguous().float()
        c = self._cache
        ws = self._ws[(B, H, W, str(dev))]
        f = self.features
        g = 4 * self.growth_rate
        gr = self.growth_rate
        offs, o = {}, 0
        for name, bn in self._bns():
            offs[name] = (o, bn.num_features)
            o += bn.num_features
        mean_all = ws["bmean"]
        inv_all = torch.rsqrt(ws["bvar"] + _EPS)
        cmax = _up4(max(p[2] for p in self._plan))
        sums = torch.zeros(2 * cmax, dtype=torch.float64, device=dev)
        out = {}

        def bn_bwd(name, bn, prefix, grad, g_pitch, xs, x_pitch, pre, relu, pool, h, w, M, C, dst, dst_pitch, accumulate):
            """BatchNorm(+ReLU) backward: fills d gamma / d beta, writes or accumulates the input gradient."""
            o0, _ = offs[name]
            mean, inv = mean_all[o0:o0 + C], inv_all[o0:o0 + C]
            pa, pb = (None, None) if pre is None else pre
            sums.zero_()
            _lib.check(lib.eml_bn_bwd_reduce(grad, g_pitch, xs, x_pitch, _lib.ptr(pa), _lib.ptr(pb), _lib.ptr(mean), _lib.ptr(inv),
                                             _lib.ptr(bn.weight), _lib.ptr(bn.bias), relu, pool, h, w, M, C, _lib.ptr(sums), cmax, st),
                       "eml_bn_bwd_reduce(%s)" % name)
            _lib.check(lib.eml_bn_bwd_apply(grad, g_pitch, xs, x_pitch, _lib.ptr(pa), _lib.ptr(pb), _lib.ptr(mean), _lib.ptr(inv),
                                            _lib.ptr(bn.weight), _lib.ptr(bn.bias), relu, pool, h, w, M, C, _lib.ptr(sums), cmax,
                                            dst, dst_pitch, accumulate, 0, st), "eml_bn_bwd_apply(%s)" % name)
            out[prefix + ".weight"] = sums[cmax:cmax + C].float()
            out[prefix + ".bias"] = sums[:C].float()

        # ---- heads and fc: plain small GEMMs (torch.matmul / cuBLAS)
        heads = (("fc_dist", self.fc_dist), ("fc_intensity", self.fc_intensity), ("fc_rgb_ratio", self.fc_rgb_ratio), ("fc_ambient", self.fc_ambient))
        dh = torch.cat([(gd if gd is not None else torch.zeros(B, m.out_features, device=dev)).float().reshape(B, -1)
                        for gd, (_, m) in zip(grads, heads)], 1)
        fc_out, pooled = ws["fc"], ws["pooled"]
        dwh = dh.t() @ fc_out
        o = 0
        for n, m in heads:
            out[n + ".weight"] = dwh[o:o + m.out_features].contiguous()
            out[n + ".bias"] = dh[:, o:o + m.out_features].sum(0)
            o += m.out_features
        dfc = dh @ c["head_w"]
        c_last = self._plan[-1][3]
        P = self.fc.in_features // c_last
        out["fc.weight"] = (dfc.t() @ pooled).view(-1, P, c_last).permute(0, 2, 1).reshape(self.fc.out_features, -1).contiguous()
        out["fc.bias"] = dfc.sum(0)
        sink = self._grad_sink
        if sink is not None:                                                     # fc + heads = 90 % of the gradient bytes, finished first:
            sink(out)                                                            # their all-reduce overlaps the whole convolutional backward
        dpool = dfc @ c["fc_w"]                                                  # (B, P*c_last) in (yo, xo, c) order
        hl, wl = ws["t_last"].shape[1], ws["t_last"].shape[2]
        k = self.avgpool_size
        dz = (dpool.view(B, hl // k, wl // k, c_last).repeat_interleave(k, 1).repeat_interleave(k, 2) / float(k * k)).contiguous()
        # relu + last_norm{last} backward -> gradient w.r.t. the stored raw transition output
        nb = len(self._plan)
        dt = torch.zeros(B, hl, wl, ws["t_last"].shape[3], dtype=torch.float32, device=dev)
        ln = getattr(f, "last_norm%d" % self._plan[-1][0])
        bn_bwd("ln%d" % self._plan[-1][0], ln, "features.last_norm%d" % self._plan[-1][0], _lib.ptr(dz), c_last, _lib.ptr(ws["t_last"]),
               ws["t_last"].shape[3], None, 1, 0, hl, wl, B * hl * wl, c_last, _lib.ptr(dt), dt.shape[3], 0)
        dbg = getattr(self, "_debug", None)
        if dbg is not None:
            dbg["dz"] = dz.clone(); dbg["dt_last"] = dt.clone(); dbg["t_last"] = ws["t_last"].clone()
            o0, _ = offs["ln%d" % self._plan[-1][0]]
            dbg["ln_mean"] = mean_all[o0:o0 + c_last].clone(); dbg["ln_inv"] = inv_all[o0:o0 + c_last].clone()

        for bi in range(nb - 1, -1, -1):
            b, c_in, c_out, c_tr = self._plan[bi]
            slab = ws["slab"][bi]
            pitch = slab.shape[3]
            h, w = ws["geom"][bi]
            M = B * h * w
            Mp = B * (h // 2) * (w // 2)
            pre = (c["pre"][bi, 0], c["pre"][bi, 1])
            dS = torch.zeros(B, h, w, pitch, dtype=torch.float32, device=dev)
            # ---- transition b: dp = dt . W_t ; dW_t ; BN(+ReLU, pooled gradient) backward into the slab gradient
            tr = getattr(f, "transition%d" % b)
            wt = tr.conv.weight.detach().float()                                  # (c_tr, c_out, 1, 1)
            dp = torch.empty(B, h // 2, w // 2, _up4(c_out), dtype=torch.float32, device=dev)
            if _up4(c_out) > c_out:
                dp[..., c_out:].zero_()                    # pitch padding: read as part of the last channel quad (masked, but defined)
            self._gemm_bwd(dt.data_ptr(), dt.shape[3], B, h // 2, w // 2, c_tr, wt.permute(1, 0, 2, 3).contiguous(), dp, _lib.EML_CONV_1x1)
            a_t = self._aff(c, "t%d.norm" % b)
            dwt = torch.zeros(c_tr, c_out, dtype=torch.float32, device=dev)
            if self.precision != "fp32" and (dt.shape[3] & 3) == 0:
                # pooled activation once, then the tensor-core wgrad (K = pooled pixels) per slice of <= 64 output channels
                ap = torch.empty(B, h // 2, w // 2, _up4(c_out), dtype=torch.float32, device=dev)
                _lib.check(lib.eml_pool_act(_lib.ptr(slab), pitch, _lib.ptr(a_t[0]), _lib.ptr(a_t[1]), B, h, w, c_out, _lib.ptr(ap), ap.shape[3], st),
                           "eml_pool_act(transition%d)" % b)
                for n0 in range(0, c_tr, 64):
                    nn_ = min(64, c_tr - n0)
                    _lib.check(lib.eml_wgrad_1x1(c_void_p(dt.data_ptr() + 4 * n0), dt.shape[3], nn_, _lib.ptr(ap), ap.shape[3], c_out, None, None,
                                                 0, 0, h // 2, w // 2, c_void_p(dwt.data_ptr() + 4 * n0 * c_out), Mp, _lib.PRECISIONS[self.precision], st),
                               "eml_wgrad_1x1(transition%d)" % b)
                del ap
            else:
                _lib.check(lib.eml_wgrad_1x1(_lib.ptr(dt), dt.shape[3], c_tr, _lib.ptr(slab), pitch, c_out, _lib.ptr(a_t[0]), _lib.ptr(a_t[1]),
                                             1, 1, h, w, _lib.ptr(dwt), Mp, _lib.PRECISIONS[self.precision], st), "eml_wgrad_1x1(transition%d)" % b)
            out["features.transition%d.conv.weight" % b] = dwt.view(c_tr, c_out, 1, 1)
            bn_bwd("t%d.norm" % b, tr.norm, "features.transition%d.norm" % b, _lib.ptr(dp), dp.shape[3], _lib.ptr(slab), pitch, pre, 1, 1,
                   h, w, M, c_out, _lib.ptr(dS), pitch, 1)
            del dp
            # ---- dense layers, last to first
            blk = getattr(f, "denseblock%d" % b)
            layers = list(blk.children())
            dN = torch.empty(B, h, w, g, dtype=torch.float32, device=dev)
            prec = _lib.PRECISIONS[self.precision]
            # fused conv1 backward (csrc/dense_bwd1.cu): dA never reaches memory, BatchNorm's mean terms are deferred as per-channel
            # affine coefficients (coefA + coefB * x) applied when a channel range's gradient is read
            fused1 = bool(lib.eml_dense_bwd1_supported(c_out, M, prec))
            if fused1:
                coef = torch.zeros(2, pitch, dtype=torch.float32, device=dev)
                vec = torch.empty(5, pitch, dtype=torch.float32, device=dev)
                w1pack = torch.empty(lib.eml_dense_bwd1_wpack_bytes(c_out), dtype=torch.uint8, device=dev)
            for l in range(len(layers) - 1, -1, -1):
                layer = layers[l]
                ci = c_in + l * gr
                pfx = "features.denseblock%d.denselayer%d" % (b, l + 1)
                n1, n2 = "b%d.l%d.norm1" % (b, l), "b%d.l%d.norm2" % (b, l)
                a1, a2 = self._aff(c, n1), self._aff(c, n2)
                if ws.get("bott_kept"):
                    bott = ws["bott_keep"][(b, l)]            # kept by the forward
                else:                                         # recompute the bottleneck (conv1 output) instead of having stored it
                    bott = ws["bott"]
                    self._conv(c, "b%d.l%d.conv1" % (b, l), slab, pitch, h, w, B, ci, bott, g, 0, g, _lib.EML_CONV_1x1, 1, a1, None, g)
                # gradient of this layer's 12 output channels; compacted because block 3's channel offsets (150 + 12 l) are not
                # 16-byte aligned and the gather uses float4 loads
                if fused1:
                    dy = torch.empty(B, h, w, 16, dtype=torch.float32, device=dev)
                    _lib.check(lib.eml_dense_bwd1_gather(_lib.ptr(dS), pitch, _lib.ptr(slab), pitch, _lib.ptr(coef[0]), _lib.ptr(coef[1]), ci, gr,
                                                         _lib.ptr(dy), 16, 16, M, st), "eml_dense_bwd1_gather")
                else:
                    dy = torch.zeros(B, h, w, 16, dtype=torch.float32, device=dev)    # 12 -> 16 channels: what the rolling 3x3 kernel stages
                    dy[..., :gr] = dS[..., ci:ci + gr]
                w2 = layer.conv2.weight.detach().float()                          # (12, 48, 3, 3)
                self._gemm_bwd(dy.data_ptr(), 16, B, h, w, gr, w2.permute(1, 0, 2, 3).flip(2, 3).contiguous(), dN, _lib.EML_CONV_3x3)
                dw2 = torch.zeros(gr, g, 3, 3, dtype=torch.float32, device=dev)
                _lib.check(lib.eml_wgrad_3x3(_lib.ptr(dy), 16, gr, _lib.ptr(bott), g, g, _lib.ptr(a2[0]), _lib.ptr(a2[1]), _lib.ptr(dw2),
                                             B, h, w, _lib.PRECISIONS[self.precision], st), "eml_wgrad_3x3")
                out[pfx + ".conv2.weight"] = dw2
                bn_bwd(n2, layer.norm2, pfx + ".norm2", _lib.ptr(dN), g, _lib.ptr(bott), g, None, 0, 0, h, w, M, g, _lib.ptr(dN), g, 0)
                w1 = layer.conv1.weight.detach().float()                          # (48, ci, 1, 1)
                dw1 = torch.zeros(g, ci, dtype=torch.float32, device=dev)
                if fused1:
                    o0, _ = offs[n1]
                    _lib.check(lib.eml_dense_bwd1_prep(_lib.ptr(a1[0]), _lib.ptr(a1[1]), _lib.ptr(pre[0]), _lib.ptr(pre[1]),
                                                       _lib.ptr(mean_all[o0:o0 + ci]), _lib.ptr(inv_all[o0:o0 + ci]), _lib.ptr(layer.norm1.weight),
                                                       ci, pitch, _lib.ptr(vec), st), "eml_dense_bwd1_prep")
                    _lib.check(lib.eml_dense_bwd1_pack(_lib.ptr(w1), _lib.ptr(w1pack), ci, st), "eml_dense_bwd1_pack")
                    sums.zero_()
                    _lib.check(lib.eml_dense_bwd1(_lib.ptr(dN), _lib.ptr(slab), pitch, _lib.ptr(dS), pitch, _lib.ptr(w1pack), _lib.ptr(vec), pitch,
                                                  ci, M, _lib.ptr(sums), cmax, prec, st), "eml_dense_bwd1(%s)" % pfx)
                    _lib.check(lib.eml_wgrad_1x1(_lib.ptr(dN), g, g, _lib.ptr(slab), pitch, ci, _lib.ptr(a1[0]), _lib.ptr(a1[1]), 1, 0, h, w,
                                                 _lib.ptr(dw1), M, _lib.PRECISIONS[self.precision], st), "eml_wgrad_1x1(conv1)")
                    dgb = torch.empty(2, ci, dtype=torch.float32, device=dev)
                    _lib.check(lib.eml_dense_bwd1_accum(_lib.ptr(sums), cmax, _lib.ptr(vec), pitch, float(M), ci, _lib.ptr(coef[0]),
                                                        _lib.ptr(coef[1]), _lib.ptr(dgb[0]), _lib.ptr(dgb[1]), st), "eml_dense_bwd1_accum")
                    out[pfx + ".norm1.weight"] = dgb[0]
                    out[pfx + ".norm1.bias"] = dgb[1]
                    out[pfx + ".conv1.weight"] = dw1.view(g, ci, 1, 1)
                    continue
                dA = torch.empty(B, h, w, _up4(ci), dtype=torch.float32, device=dev)
                self._gemm_bwd(dN.data_ptr(), g, B, h, w, g, w1.permute(1, 0, 2, 3).contiguous(), dA, _lib.EML_CONV_1x1)
                _lib.check(lib.eml_wgrad_1x1(_lib.ptr(dN), g, g, _lib.ptr(slab), pitch, ci, _lib.ptr(a1[0]), _lib.ptr(a1[1]), 1, 0, h, w,
                                             _lib.ptr(dw1), M, _lib.PRECISIONS[self.precision], st), "eml_wgrad_1x1(conv1)")
                out[pfx + ".conv1.weight"] = dw1.view(g, ci, 1, 1)
                bn_bwd(n1, layer.norm1, pfx + ".norm1", _lib.ptr(dA), dA.shape[3], _lib.ptr(slab), pitch, pre, 1, 0, h, w, M, ci,
                       _lib.ptr(dS), pitch, 1)
                del dA
            if fused1:                                                            # the block-input channels' deferred BatchNorm mean terms
                _lib.check(lib.eml_dense_bwd1_gather(_lib.ptr(dS), pitch, _lib.ptr(slab), pitch, _lib.ptr(coef[0]), _lib.ptr(coef[1]), 0, c_in,
                                                     _lib.ptr(dS), pitch, c_in, M, st), "eml_dense_bwd1_gather(block input)")
            # ---- block input
            if bi > 0:
                pb, _, _, ptr_c = self._plan[bi - 1]
                ln = getattr(f, "last_norm%d" % pb)
                dt = torch.zeros(B, h, w, pitch, dtype=torch.float32, device=dev)
                bn_bwd("ln%d" % pb, ln, "features.last_norm%d" % pb, _lib.ptr(dS), pitch, _lib.ptr(slab), pitch, None, 0, 0, h, w, M, ptr_c,
                       _lib.ptr(dt), pitch, 0)
            else:
                c0 = f.conv0.out_channels
                z0 = torch.empty(B, h, w, c0, dtype=torch.float32, device=dev)
                _lib.check(lib.eml_stem_forward(_lib.ptr(x), _lib.ptr(c["w0"]), None, None, _lib.ptr(z0), c0, None, None, 0, B, H, W, c0, 1, 0, st),
                           "eml_stem_forward(recompute)")
                dz0 = torch.empty_like(z0)
                bn_bwd("norm0", f.norm0, "features.norm0", _lib.ptr(dS), pitch, _lib.ptr(z0), c0, None, 1, 0, h, w, M, c0, _lib.ptr(dz0), c0, 0)
                dw0 = torch.zeros(c0, 3, 3, 3, dtype=torch.float32, device=dev)
                _lib.check(lib.eml_wgrad_stem(_lib.ptr(dz0), c0, c0, _lib.ptr(x), _lib.ptr(dw0), B, H, W, st), "eml_wgrad_stem")
                out["features.conv0.weight"] = dw0
            del dS
            if sink is not None:
                sink(out)
        return out
