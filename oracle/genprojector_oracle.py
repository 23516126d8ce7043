"""CPU oracle for the GenProjector SPADE/SphereConv generator (test infrastructure, see oracle/__init__.py).

Functional restatement (eval / inference mode, the path GenProjector/test.py:21-39 runs) of

* SphereConv2D            GenProjector/models/networks/spherenet/sphere_cnn.py:11-124
    tangent-plane 3x3 sampling pattern (get_xy :11-28, cal_index :31-58: gnomonic projection, centre forced to the pixel,
    longitude wrapped mod W), laid out as a (1,3H/s,3W/s,2) grid (:75-84), then grid_sample (bilinear, zeros padding,
    align_corners=False -- the torch>=1.3 default the in-container reference runs with) and conv2d(stride=3) (:122-123)
* SPADE                   models/networks/normalization.py:68-115   BN(no affine, running stats) * (1+gamma) + beta,
                          gamma/beta = SphereConv(ReLU(SphereConv(nearest-resized guide)))
* SPADEResnetBlock        models/networks/architecture.py:22-69     learned shortcut when fin != fout, LeakyReLU(0.2)
* spectral_norm (eval)    torch.nn.utils.spectral_norm: W = W_orig / (u^T W_mat v) with the stored u, v (no power iteration)
* ConvEncoder             models/networks/generator.py:90-126       bilinear 128^2, 5 x [conv3x3 s2 + InstanceNorm] with LeakyReLU, fc
* SPADEGenerator.forward  models/networks/generator.py:65-88        7 blocks / 5 nearest x2 upsamples, (tanh+1)*25

over a state_dict with the reference's parameter names.
"""
import math
from collections import OrderedDict
from functools import lru_cache

import numpy as np
import torch
import torch.nn.functional as F

BLOCKS = (("head_0", 16, 16), ("G_middle_0", 16, 16), ("G_middle_1", 16, 16), ("up_0", 16, 8), ("up_1", 8, 4),
          ("up_2", 4, 2), ("up_3", 2, 1))
NHIDDEN = 128


@lru_cache(None)
def sphere_sample_coords(h, w, stride=1):
    """(Ho, Wo, 3, 3, 2) float64 (row, col) sample positions in pixel units -- sphere_cnn.py:11-72 vectorised."""
    pi = np.pi
    dphi, dth = pi / h, 2 * pi / w
    t, s = math.tan(dth), 1 / math.cos(dth) * math.tan(dphi)
    xs = np.array([[-t, 0, t], [-t, 1, t], [-t, 0, t]], dtype=np.float64)
    ys = np.array([[s, math.tan(dphi), s], [0, 1, 0], [-s, -math.tan(dphi), -s]], dtype=np.float64)
    r = np.arange(0, h, stride, dtype=np.float64)[:, None, None, None]
    c = np.arange(0, w, stride, dtype=np.float64)[None, :, None, None]
    phi = -((r + 0.5) / h * pi - pi / 2)
    theta = (c + 0.5) / w * 2 * pi - pi
    x, y = xs[None, None], ys[None, None]
    rho = np.sqrt(x ** 2 + y ** 2)
    v = np.arctan(rho)
    new_phi = np.arcsin(np.cos(v) * np.sin(phi) + y * np.sin(v) * np.cos(phi) / rho)
    new_theta = theta + np.arctan(x * np.sin(v) / (rho * np.cos(phi) * np.cos(v) - y * np.sin(phi) * np.sin(v)))
    new_r = (-new_phi + pi / 2) * h / pi - 0.5
    new_c = ((new_theta + pi) * w / 2 / pi - 0.5 + w) % w
    out = np.stack(np.broadcast_arrays(new_r, new_c), -1)
    out[:, :, 1, 1, 0] = np.arange(0, h, stride)[:, None]
    out[:, :, 1, 1, 1] = np.arange(0, w, stride)[None, :]
    return out


def sphere_grid(h, w, stride=1):
    """The (1, 3Ho, 3Wo, 2) fp32 grid_sample grid of sphere_cnn.py:75-84 (x = col, y = row, normalised as 2p/size - 1)."""
    co = sphere_sample_coords(h, w, stride)
    gy = co[..., 0] * 2 / h - 1
    gx = co[..., 1] * 2 / w - 1
    g = np.stack((gx, gy), -1)                                    # (Ho, Wo, 3, 3, 2)
    ho, wo = g.shape[:2]
    g = g.transpose(0, 2, 1, 3, 4).reshape(1, ho * 3, wo * 3, 2)
    return torch.from_numpy(g.astype(np.float32))


_GRIDS = {}


def sphere_conv(x, weight, bias, stride=1):
    key = (x.shape[2], x.shape[3], stride, str(x.device), x.dtype)          # the reference module keeps its grid too (sphere_cnn.py:111-118)
    if key not in _GRIDS:
        _GRIDS[key] = sphere_grid(x.shape[2], x.shape[3], stride).to(device=x.device, dtype=x.dtype)   # fp32 coordinates (as the reference); cast only for the fp64 debug runs
    grid = _GRIDS[key].expand(x.shape[0], -1, -1, -1)
    s = F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    return F.conv2d(s, weight, bias, stride=3)


def sn_weight(sd, prefix, upd=None):
    """Spectral norm: W_orig / sigma, sigma = u . (W_mat v).  Eval mode (upd is None): the stored u, v.  Training mode (upd = dict that
    receives the new buffers): torch.nn.utils.spectral_norm's single power iteration first, v = normalize(W^T u), u = normalize(W v)."""
    w = sd[prefix + ".weight_orig"]
    u, v = sd[prefix + ".weight_u"], sd[prefix + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if upd is not None:
        with torch.no_grad():              # like torch's SpectralNorm.compute_weight: u, v are constants of the autograd graph
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=1e-12)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=1e-12)
        upd[prefix + ".weight_u"], upd[prefix + ".weight_v"] = u, v
    sigma = torch.dot(u, torch.mv(wm, v))
    return w / sigma


def spade(sd, p, x, guide, upd=None):
    rm, rv = sd[p + ".param_free_norm.running_mean"], sd[p + ".param_free_norm.running_var"]
    if upd is None:
        normalized = F.batch_norm(x, rm, rv, None, None, False, 0.0, 1e-5)
    else:                                  # training: batch statistics + running-stat update (normalization.py:80, momentum 0.1)
        rm, rv = rm.clone(), rv.clone()
        normalized = F.batch_norm(x, rm, rv, None, None, True, 0.1, 1e-5)
        upd[p + ".param_free_norm.running_mean"], upd[p + ".param_free_norm.running_var"] = rm, rv
    seg = F.interpolate(guide, size=x.shape[2:], mode="nearest")
    actv = F.relu(sphere_conv(seg, sd[p + ".mlp_shared.0.weight"], sd[p + ".mlp_shared.0.bias"]))
    gamma = sphere_conv(actv, sd[p + ".mlp_gamma.weight"], sd[p + ".mlp_gamma.bias"])
    beta = sphere_conv(actv, sd[p + ".mlp_beta.weight"], sd[p + ".mlp_beta.bias"])
    return normalized * (1 + gamma) + beta


def spade_block(sd, p, x, guide, learned, upd=None):
    x_s = x
    if learned:
        x_s = sphere_conv(spade(sd, p + ".norm_s", x, guide, upd), sn_weight(sd, p + ".conv_s", upd), sd[p + ".conv_s.bias"])
    dx = sphere_conv(F.leaky_relu(spade(sd, p + ".norm_0", x, guide, upd), 0.2), sn_weight(sd, p + ".conv_0", upd), sd[p + ".conv_0.bias"])
    dx = sphere_conv(F.leaky_relu(spade(sd, p + ".norm_1", dx, guide, upd), 0.2), sn_weight(sd, p + ".conv_1", upd), sd[p + ".conv_1.bias"])
    return x_s + dx


def encoder(sd, crop, upd=None):
    x = F.interpolate(crop, size=(128, 128), mode="bilinear")
    for i in range(1, 6):
        if i > 1:
            x = F.leaky_relu(x, 0.2)
        x = F.conv2d(x, sn_weight(sd, "netE.layer%d.0" % i, upd), None, stride=2, padding=1)
        x = F.instance_norm(x, eps=1e-5)
    x = F.leaky_relu(x, 0.2)
    return F.linear(x.reshape(x.shape[0], -1), sd["netE.fc.weight"], sd["netE.fc.bias"])


def generator_forward(sd, guide, crop, ngf=64, taps=None, upd=None):
    """guide (B,3,128,256), crop (B,3,Hc,Wc) -> (B,3,128,256) in [0,50].  upd = None: eval mode; upd = {}: the train-mode forward
    (batch-statistic BatchNorm in SPADE, one spectral-norm power iteration per wrapped conv), upd receives every updated buffer."""
    x = encoder(sd, crop, upd).view(-1, 16 * ngf, 1, 2)
    x = F.interpolate(x, size=(4, 8))
    if taps is not None:
        taps["latent"] = x
    for i, (name, fi, fo) in enumerate(BLOCKS):
        x = spade_block(sd, name, x, guide, fi != fo, upd)
        if taps is not None:
            taps[name] = x
        if name not in ("G_middle_0", "up_3"):
            x = F.interpolate(x, scale_factor=2)
    x = sphere_conv(F.leaky_relu(x, 0.2), sd["sphere_conv1.weight"], sd["sphere_conv1.bias"])
    return (torch.tanh(x) + 1) * 25


def init_generator_state_dict(seed=0, ngf=64):
    """Deterministic parameters with the reference's names / shapes (253 entries, 118.4 M values at ngf=64)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def conv(name, o, i, k=3, gain=1.0, bias=True, spectral=False):
        w = torch.randn(o, i, k, k, generator=g) * (gain / math.sqrt(i * k * k))
        if bias:
            sd[name + ".bias"] = 0.1 * torch.randn(o, generator=g)
        if spectral:
            sd[name + ".weight_orig"] = w
            sd[name + ".weight_u"] = F.normalize(torch.randn(o, generator=g), dim=0)
            sd[name + ".weight_v"] = F.normalize(torch.randn(i * k * k, generator=g), dim=0)
        else:
            sd[name + ".weight"] = w

    def spade_p(name, c):
        sd[name + ".param_free_norm.running_mean"] = 0.1 * torch.randn(c, generator=g)
        sd[name + ".param_free_norm.running_var"] = 0.5 + torch.rand(c, generator=g)
        sd[name + ".param_free_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        conv(name + ".mlp_shared.0", NHIDDEN, 3, gain=2.0)
        conv(name + ".mlp_gamma", c, NHIDDEN, gain=0.5)
        conv(name + ".mlp_beta", c, NHIDDEN, gain=0.5)

    for name, fi, fo in BLOCKS:
        fin, fout = fi * ngf, fo * ngf
        fmid = min(fin, fout)
        # spectral norm divides by u.Wv (|.| << the true sigma for random u, v): keep W_orig small enough that W/sigma is O(1)
        conv(name + ".conv_0", fmid, fin, spectral=True)
        conv(name + ".conv_1", fout, fmid, spectral=True)
        if fin != fout:
            conv(name + ".conv_s", fout, fin, spectral=True)
        spade_p(name + ".norm_0", fin)
        spade_p(name + ".norm_1", fmid)
        if fin != fout:
            spade_p(name + ".norm_s", fin)
    conv("sphere_conv1", 3, ngf, gain=0.15)                     # keep the final tanh out of saturation so errors stay visible
    chans = [3, ngf, 2 * ngf, 4 * ngf, 8 * ngf, 8 * ngf]
    for i in range(1, 6):
        conv("netE.layer%d.0" % i, chans[i], chans[i - 1], bias=False, spectral=True)
    sd["netE.fc.weight"] = torch.randn(16 * ngf * 2, 8 * ngf * 16, generator=g) / math.sqrt(8 * ngf * 16)
    sd["netE.fc.bias"] = 0.1 * torch.randn(16 * ngf * 2, generator=g)
    # make the spectral-norm sigma well-conditioned: align u, v with one power iteration of each weight
    for k in [k for k in sd if k.endswith(".weight_orig")]:
        w = sd[k].reshape(sd[k].shape[0], -1)
        v = F.normalize(torch.mv(w.t(), sd[k[:-5] + "_u"]), dim=0)
        u = F.normalize(torch.mv(w, v), dim=0)
        sd[k[:-5] + "_u"], sd[k[:-5] + "_v"] = u, v
    return sd


# ===================================================================================== discriminator + losses (G6-G8)
# * NLayerDiscriminator     models/networks/discriminator.py:69-125    SphereConv(6->ndf, s2)+LeakyReLU; 3 x [spectral SphereConv
#                           (s2, s2, s1) + InstanceNorm2d(affine=False) + LeakyReLU]; SphereConv(->3); all 5 outputs returned
# * MultiscaleDiscriminator models/networks/discriminator.py:16-65     D_i on the input avg-pooled i times (3x3 s2 p1, pad not counted)
# * GANLoss (hinge)         models/networks/loss.py:57-98              mean over the multiscale list of the per-scale scalar
# * VGG19 / VGGLoss         models/networks/architecture.py:92-122, loss.py:102-114   torchvision vgg19.features[0:30], 5 relu taps
# * generator / discriminator loss composition   models/pix2pix_model.py:92-141
VGG_CFG = ((0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256), (16, 256, 256),
           (19, 256, 512), (21, 512, 512), (23, 512, 512), (25, 512, 512), (28, 512, 512))
VGG_SLICES = ((0,), (2, "P", 5), (7, "P", 10), (12, 14, 16, "P", 19), (21, 23, 25, "P", 28))
VGG_WEIGHTS = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)


def d_channels(ndf=64, n_layers=4):
    ch = [ndf]
    for _ in range(1, n_layers):
        ch.append(min(ch[-1] * 2, 512))
    return ch


def nlayer_discriminator(sd, p, x, n_layers=4):
    """One PatchGAN: returns the n_layers+1 intermediate outputs (discriminator.py:113-123)."""
    outs = []
    x = F.leaky_relu(sphere_conv(x, sd[p + "model0.0.weight"], sd[p + "model0.0.bias"], stride=2), 0.2)
    outs.append(x)
    for n in range(1, n_layers):
        stride = 1 if n == n_layers - 1 else 2
        x = sphere_conv(x, sn_weight(sd, p + "model%d.0.0" % n), None, stride=stride)       # bias removed by the instance-norm wrapper
        x = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
        outs.append(x)
    outs.append(sphere_conv(x, sd[p + "model%d.0.weight" % n_layers], sd[p + "model%d.0.bias" % n_layers], stride=1))
    return outs


def multiscale_discriminator(sd, x, num_D=2, n_layers=4):
    result = []
    for i in range(num_D):
        result.append(nlayer_discriminator(sd, "discriminator_%d." % i, x, n_layers))
        x = F.avg_pool2d(x, kernel_size=3, stride=2, padding=[1, 1], count_include_pad=False)
    return result


def hinge_loss(preds, target_is_real, for_discriminator=True):
    """GANLoss('hinge').__call__ on a multiscale list (loss.py:65-98)."""
    total = 0.0
    for p in preds:
        p = p[-1] if isinstance(p, (list, tuple)) else p
        if not for_discriminator:
            l = -p.mean()
        elif target_is_real:
            l = -torch.clamp(p - 1, max=0).mean()
        else:
            l = -torch.clamp(-p - 1, max=0).mean()
        total = total + l
    return total / len(preds)


def feature_matching_loss(pred_fake, pred_real, mask):
    """pix2pix_model.py:101-117 -- note the mask is re-interpolated from its PREVIOUS size at every layer (:111)."""
    num_D = len(pred_fake)
    loss = 0.0
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            h, w = pred_fake[i][j].shape[2:]
            mask = F.interpolate(mask, size=(h, w))
            fw = pred_fake[i][j] * mask + pred_fake[i][j] * (1 - mask) * 50
            rw = pred_real[i][j] * mask + pred_real[i][j] * (1 - mask) * 50
            loss = loss + F.l1_loss(fw, rw) / num_D
    return loss


def vgg_features(sd, x, p="vgg."):
    outs = []
    for s, ops in enumerate(VGG_SLICES):
        for op in ops:
            if op == "P":
                x = F.max_pool2d(x, 2, 2)
            else:
                x = F.relu(F.conv2d(x, sd["%sslice%d.%d.weight" % (p, s + 1, op)], sd["%sslice%d.%d.bias" % (p, s + 1, op)], padding=1))
        outs.append(x)
    return outs


def vgg_loss(sd, x, y, p="vgg."):
    fx, fy = vgg_features(sd, x, p), vgg_features(sd, y, p)
    return sum(w * F.l1_loss(a, b) for w, a, b in zip(VGG_WEIGHTS, fx, fy))


def cosine_loss(fake, real):
    return (1 - F.cosine_similarity(fake, real, dim=1, eps=1e-20)).mean()


def split_pred(pred):
    """Pix2PixModel.divide_pred (pix2pix_model.py:164-178)."""
    return ([[t[:t.size(0) // 2] for t in p] for p in pred], [[t[t.size(0) // 2:] for t in p] for p in pred])


def generator_losses(sd_d, sd_vgg, guide, fake, real, mask, num_D=2, n_layers=4):
    """compute_generator_loss (pix2pix_model.py:92-128) given the generated image: {'GAN','GAN_Feat','VGG','COS'}."""
    pred = multiscale_discriminator(sd_d, torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0), num_D, n_layers)
    pf, pr = split_pred(pred)
    return {"GAN": hinge_loss(pf, True, False), "GAN_Feat": feature_matching_loss(pf, pr, mask),
            "VGG": vgg_loss(sd_vgg, fake, real) * 5, "COS": cosine_loss(fake, real) * 5}


def discriminator_losses(sd_d, guide, fake, real, num_D=2, n_layers=4):
    """compute_discriminator_loss (pix2pix_model.py:130-141): {'D_Fake','D_real'}."""
    pred = multiscale_discriminator(sd_d, torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0), num_D, n_layers)
    pf, pr = split_pred(pred)
    return {"D_Fake": hinge_loss(pf, False, True), "D_real": hinge_loss(pr, True, True)}


def init_discriminator_state_dict(seed=0, ndf=64, num_D=2, n_layers=4, input_nc=6):
    """Deterministic parameters with the reference's names (spectral layers: weight_orig/_u/_v, bias deleted by the norm wrapper)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    ch = d_channels(ndf, n_layers)
    for d in range(num_D):
        p = "discriminator_%d." % d
        sd[p + "model0.0.weight"] = torch.randn(ch[0], input_nc, 3, 3, generator=g) / math.sqrt(input_nc * 9)
        sd[p + "model0.0.bias"] = 0.1 * torch.randn(ch[0], generator=g)
        for n in range(1, n_layers):
            w = torch.randn(ch[n], ch[n - 1], 3, 3, generator=g) / math.sqrt(ch[n - 1] * 9)
            u = F.normalize(torch.randn(ch[n], generator=g), dim=0)
            wm = w.reshape(ch[n], -1)
            v = F.normalize(torch.mv(wm.t(), u), dim=0)
            u = F.normalize(torch.mv(wm, v), dim=0)
            sd[p + "model%d.0.0.weight_orig" % n] = w
            sd[p + "model%d.0.0.weight_u" % n] = u
            sd[p + "model%d.0.0.weight_v" % n] = v
        sd[p + "model%d.0.weight" % n_layers] = torch.randn(3, ch[-1], 3, 3, generator=g) / math.sqrt(ch[-1] * 9)
        sd[p + "model%d.0.bias" % n_layers] = 0.1 * torch.randn(3, generator=g)
    return sd


def init_vgg_state_dict(seed=0, p="vgg."):
    """He-initialised VGG19 feature weights under the reference's slice names (there is no network for the ImageNet checkpoint;
    parity is about the arithmetic, the production weights load through the same keys)."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    slice_of = {op: s + 1 for s, ops in enumerate(VGG_SLICES) for op in ops if op != "P"}
    for idx, ci, co in VGG_CFG:
        sd["%sslice%d.%d.weight" % (p, slice_of[idx], idx)] = torch.randn(co, ci, 3, 3, generator=g) * math.sqrt(2.0 / (ci * 9))
        sd["%sslice%d.%d.bias" % (p, slice_of[idx], idx)] = 0.05 * torch.randn(co, generator=g)
    return sd
