"""CPU oracle for the EMLight regression network (test infrastructure, see oracle/__init__.py).

Restates ``RegressionNetwork/DenseNet.py`` of the reference as one flat function over a
``state_dict`` carrying the reference's parameter names:

* stem               DenseNet.py:88-93   conv0 3x3 s1 p1 (no bias) -> norm0 -> relu
* dense layer        DenseNet.py:26-55   cat[x, conv2(norm2(conv1(relu(norm1(x)))))]
                     NOTE: the reference registers norm1, relu1, conv1, norm2, conv2 (DenseNet.py:30-43):
                     there is NO ReLU between norm2 and conv2 (unlike torchvision / SURVEY 8a-D2);
                     conv2's zero padding applies to the norm2 output.
* dense block        DenseNet.py:58-65   16 layers, growth 12, bottleneck 48
* transition         DenseNet.py:14-21   norm -> relu -> conv1x1 (C -> C//2) -> avgpool2
                     DenseNet.py:110     (quirk: a transition follows *every* block)
* last_norm{i}       DenseNet.py:122     extra BatchNorm after each transition
* head               DenseNet.py:135-157 relu -> avgpool4 -> flatten (NCHW order) -> fc -> 4 linears

BatchNorm: eps 1e-5, momentum 0.1 (PyTorch defaults, DenseNet.py:17,30,41,91,122).
``training=True`` uses biased batch statistics (what test.py:46 actually runs).

The ``quant`` hook rounds the *operands of every convolution / linear* to a narrower
format while keeping fp32 accumulation -- used only to predict tensor-core error budgets.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

GROWTH = 12
BLOCKS = (16, 16, 16)
INIT_FEATURES = 24
BOTTLENECK = 4 * GROWTH
EPS = 1e-5


def layer_plan(growth=GROWTH, blocks=BLOCKS, init=INIT_FEATURES):
    """[(block_index, C_in_of_block, C_out_of_block, C_after_transition)] -- DenseNet.py:96-119."""
    plan, c = [], init
    for b, n in enumerate(blocks):
        c_out = c + n * growth
        c_tr = int(math.floor(c_out * 0.5))
        plan.append((b + 1, c, c_out, c_tr))
        c = c_tr
    return plan


def init_state_dict(seed=0, n_anchors=96, dtype=torch.float32):
    """Deterministic random parameters with the reference's names and shapes.

    Not the PyTorch default initialisers (those would need the reference class); the
    distributions are chosen so activations stay O(1) through 100 convolutions and so that
    BatchNorm has non-trivial affine parameters and running statistics.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def conv(name, o, i, k):
        fan_in = i * k * k
        sd[name + ".weight"] = (torch.randn(o, i, k, k, generator=g) * math.sqrt(2.0 / fan_in)).to(dtype)

    def bn(name, c):
        sd[name + ".weight"] = (0.5 + torch.rand(c, generator=g)).to(dtype)
        sd[name + ".bias"] = (0.2 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".running_mean"] = (0.2 * torch.randn(c, generator=g)).to(dtype)
        sd[name + ".running_var"] = (0.5 + torch.rand(c, generator=g)).to(dtype)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def lin(name, o, i):
        bound = 1.0 / math.sqrt(i)
        sd[name + ".weight"] = ((torch.rand(o, i, generator=g) * 2 - 1) * bound).to(dtype)
        sd[name + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * bound).to(dtype)

    conv("features.conv0", INIT_FEATURES, 3, 3)
    bn("features.norm0", INIT_FEATURES)
    for b, c_in, c_out, c_tr in layer_plan():
        for l in range(BLOCKS[b - 1]):
            p = "features.denseblock%d.denselayer%d" % (b, l + 1)
            c = c_in + l * GROWTH
            bn(p + ".norm1", c)
            conv(p + ".conv1", BOTTLENECK, c, 1)
            bn(p + ".norm2", BOTTLENECK)
            conv(p + ".conv2", GROWTH, BOTTLENECK, 3)
        bn("features.transition%d.norm" % b, c_out)
        conv("features.transition%d.conv" % b, c_tr, c_out, 1)
        bn("features.last_norm%d" % b, c_tr)
    lin("fc", 1024, 8208)
    lin("fc_dist", n_anchors, 1024)
    lin("fc_intensity", 1, 1024)
    lin("fc_rgb_ratio", 3, 1024)
    lin("fc_ambient", 3, 1024)
    return sd


def _bn(x, sd, name, training):
    if training:
        return F.batch_norm(x, None, None, sd[name + ".weight"], sd[name + ".bias"], True, 0.0, EPS)
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"],
                        sd[name + ".weight"], sd[name + ".bias"], False, 0.0, EPS)


def densenet_forward(sd, x, training=False, quant=None, taps=None):
    """x: (B,3,H,W) fp32 NCHW in [0,1]; returns the 4-key dict of DenseNet.py:153-157.

    ``taps``: optional dict filled with named intermediates (NCHW) for kernel-level tests.
    """
    q = (lambda t: t) if quant is None else quant

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    h = F.conv2d(q(x), q(sd["features.conv0.weight"]), padding=1)
    h = F.relu(_bn(h, sd, "features.norm0", training))
    tap("stem", h)
    for b, c_in, c_out, c_tr in layer_plan():
        for l in range(BLOCKS[b - 1]):
            p = "features.denseblock%d.denselayer%d" % (b, l + 1)
            t = F.relu(_bn(h, sd, p + ".norm1", training))
            t = F.conv2d(q(t), q(sd[p + ".conv1.weight"]))
            t = _bn(t, sd, p + ".norm2", training)          # no ReLU here (DenseNet.py:41-43)
            t = F.conv2d(q(t), q(sd[p + ".conv2.weight"]), padding=1)
            h = torch.cat([h, t], 1)
        tap("block%d" % b, h)
        p = "features.transition%d" % b
        t = F.relu(_bn(h, sd, p + ".norm", training))
        t = F.conv2d(q(t), q(sd[p + ".conv.weight"]))
        t = F.avg_pool2d(t, 2, 2)
        tap("trans%d" % b, t)
        h = _bn(t, sd, "features.last_norm%d" % b, training)
    tap("features", h)
    out = F.avg_pool2d(F.relu(h), 4).reshape(h.shape[0], -1)
    tap("pooled", out)
    out = F.linear(q(out), q(sd["fc.weight"]), sd["fc.bias"])
    tap("fc", out)
    return {
        "distribution": F.linear(q(out), q(sd["fc_dist.weight"]), sd["fc_dist.bias"]),
        "intensity": F.linear(q(out), q(sd["fc_intensity.weight"]), sd["fc_intensity.bias"]),
        "rgb_ratio": F.linear(q(out), q(sd["fc_rgb_ratio.weight"]), sd["fc_rgb_ratio.bias"]),
        "ambient": F.linear(q(out), q(sd["fc_ambient.weight"]), sd["fc_ambient.bias"]),
    }


def round_tf32(t):
    """Round-to-nearest-even to 10 explicit mantissa bits (what cvt.rna.tf32.f32 produces, ties aside)."""
    i = t.contiguous().view(torch.int32)
    i = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def round_bf16(t):
    return t.to(torch.bfloat16).to(torch.float32)
