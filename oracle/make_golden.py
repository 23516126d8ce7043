"""Generate tests/golden/*.npz by running the REFERENCE's own modules (from /root/reference) on CPU.

Run in the build container only (``python -m oracle.make_golden``); /root/reference does not exist on
the GPU box, so the outputs are committed as small fixtures.  No reference source is copied: the
modules are imported in place and the two pure functions of the un-importable
``RegressionNetwork/util.py`` (merge-conflict markers) are exec'd from their exact line ranges.

Fixtures
  densenet.npz   DenseNet.DenseNet() with ``oracle.densenet_oracle.init_state_dict(seed)`` loaded, eval and
                 train-mode BN, inputs U[0,1) (2,3,192,256) from seed 1234 -> the 4 head outputs.
  sinkhorn.npz   geomloss.SamplesLoss("sinkhorn",p=2,blur=.025,batchsize=B)(x,y) and d(sum)/dx, N=96;
                 gmloss variant with a geometry vector, N=128.
  render.npz     convert_to_panorama / sphere_points outputs.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/RegressionNetwork"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def render_go(B):
    """Closed-form upstream gradient for the render backward fixtures (re-creatable in the tests)."""
    b, c, r, w = np.meshgrid(np.arange(B), np.arange(3), np.arange(128), np.arange(256), indexing="ij")
    return np.sin(0.37 * r + 0.11 * w + c + 2.0 * b).astype(np.float32)


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU-only host: .cuda() -> identity (appendix C.2)
    torch.set_num_threads(os.cpu_count())
    from oracle.densenet_oracle import init_state_dict
    os.makedirs(OUT, exist_ok=True)

    # ---------------- DenseNet (DenseNet.py imported unchanged) ----------------
    import DenseNet as RefDenseNet
    sd = init_state_dict(seed=0, n_anchors=96)
    x = torch.rand(2, 3, 192, 256, generator=torch.Generator().manual_seed(1234))
    gold = {"x_seed": np.int64(1234), "sd_seed": np.int64(0)}
    for mode in ("eval", "train"):
        m = RefDenseNet.DenseNet()
        m.load_state_dict(sd)
        m.train(mode == "train")
        with torch.no_grad():
            out = m(x)
        for k, v in out.items():
            gold["%s_%s" % (mode, k)] = v.numpy()
        if mode == "train":   # running statistics after one training-mode forward (momentum .1, unbiased var)
            gold["train_norm0_running_mean"] = m.features.norm0.running_mean.numpy().copy()
            gold["train_norm0_running_var"] = m.features.norm0.running_var.numpy().copy()
            gold["train_last_norm3_running_var"] = m.features.last_norm3.running_var.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "densenet.npz"), **gold)

    # ---------------- Sinkhorn (geomloss / gmloss imported unchanged) ----------------
    import importlib
    gold = {}
    for name, N, B in (("geomloss", 96, 4), ("gmloss", 128, 3)):
        mod = importlib.import_module(name)
        g = torch.Generator().manual_seed(77 + N)
        xs = (0.3 * torch.randn(B, N, 1, generator=g)).requires_grad_()
        ys = torch.softmax(3 * torch.randn(B, N, generator=g), 1).view(B, N, 1)
        if name == "geomloss":
            L = mod.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
            v = L(xs, ys)
        else:
            geometry = (0.5 + torch.rand(N, generator=g)).numpy()
            L = mod.SamplesLoss("sinkhorn", p=2, blur=.025)
            v = L(xs, ys, geometry)
            gold[name + "_geometry"] = geometry
            gold[name + "_M"] = L.distance.M[0].numpy()
        v.sum().backward()
        gold[name + "_x"] = xs.detach().numpy(); gold[name + "_y"] = ys.numpy()
        gold[name + "_loss"] = v.detach().numpy(); gold[name + "_grad"] = xs.grad.numpy()
        if name == "geomloss":
            gold["geomloss_M"] = L.distance.M[0].numpy()
    np.savez_compressed(os.path.join(OUT, "sinkhorn.npz"), **gold)

    # ---------------- render (util.py:222-245 and :286-299 exec'd in place) ----------------
    src = open(os.path.join(REF, "util.py")).read().split("\n")
    ns = {"torch": torch, "np": np}
    exec("\n".join(src[221:245]), ns)      # def convert_to_panorama
    exec("\n".join(src[285:299]), ns)      # def sphere_points
    gold = {}
    for N, B in ((96, 2), (128, 1)):
        g = torch.Generator().manual_seed(900 + N)
        pts = ns["sphere_points"](N)
        gold["points_%d" % N] = pts
        dirs = torch.from_numpy(pts).float().view(1, -1).repeat(B, 1)
        dirs = dirs + 0.01 * torch.randn(dirs.shape, generator=g)
        sizes = 0.0025 + 0.05 * torch.rand(B, N, generator=g)
        colors = torch.rand(B, 3 * N, generator=g)
        dirs.requires_grad_(); sizes.requires_grad_(); colors.requires_grad_()
        pano = ns["convert_to_panorama"](dirs, sizes, colors)
        go = torch.from_numpy(render_go(B))
        (pano * go).sum().backward()
        # keep the fixture small: every 2nd row/column of the panorama; grads use the closed-form render_go()
        gold["dirs_%d" % N] = dirs.detach().numpy(); gold["sizes_%d" % N] = sizes.detach().numpy()
        gold["colors_%d" % N] = colors.detach().numpy()
        gold["pano_%d" % N] = pano.detach().numpy()[:, :, ::2, ::2].astype(np.float32)
        gold["gdirs_%d" % N] = dirs.grad.numpy(); gold["gsizes_%d" % N] = sizes.grad.numpy()
        gold["gcolors_%d" % N] = colors.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "render.npz"), **gold)
    # ---------------- GenProjector generator (generator.py / architecture.py / normalization.py / sphere_cnn.py imported unchanged) -----
    import argparse
    import types
    for name in ("OpenEXR", "Imath", "imageio", "imageio.plugins", "imageio.plugins.freeimage", "vtk", "vtk.util", "vtk.util.numpy_support"):
        sys.modules.setdefault(name, types.ModuleType(name))             # IO modules the hot path never calls (appendix C.3)
    sys.modules["imageio.plugins.freeimage"].download = lambda: None
    sys.modules["imageio"].plugins = sys.modules["imageio.plugins"]
    sys.modules["imageio.plugins"].freeimage = sys.modules["imageio.plugins.freeimage"]
    sys.modules["vtk"].util = sys.modules["vtk.util"]
    sys.modules["vtk.util"].numpy_support = sys.modules["vtk.util.numpy_support"]
    sys.path.insert(0, "/root/reference/GenProjector")
    for k in [k for k in sys.modules if k == "util" or k.startswith("models")]:
        del sys.modules[k]
    from models.networks.generator import SPADEGenerator
    from models.networks.spherenet import SphereConv2D
    from oracle import genprojector_oracle as GO
    ngf = 16
    opt = argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                             num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
    G = SPADEGenerator(opt).eval()
    sdg = GO.init_generator_state_dict(seed=0, ngf=ngf)
    G.load_state_dict(sdg)
    gen = torch.Generator().manual_seed(3)
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(1, 3, 160, 160, generator=gen)
    with torch.no_grad():
        out = G(guide, crop)
        sc = SphereConv2D(5, 7, stride=2)
        xs = torch.randn(2, 5, 16, 32, generator=gen)
        ys = sc(xs)
    np.savez_compressed(os.path.join(OUT, "generator.npz"), ngf=np.int64(ngf), sd_seed=np.int64(0), in_seed=np.int64(3),
                        out=out.numpy()[:, :, ::2, ::2], out_mean=np.float64(out.double().mean()), out_std=np.float64(out.double().std()),
                        sc_weight=sc.weight.detach().numpy(), sc_bias=sc.bias.detach().numpy(), sc_x=xs.numpy(), sc_y=ys.detach().numpy())
    # ---------------- GenProjector discriminator + losses (discriminator.py / loss.py / architecture.py VGG19 / pix2pix_model.py unchanged) ----
    # Pix2PixModel.forward needs .cuda(); its loss-composition methods are called unbound on a stand-in ``self`` that carries the
    # reference's own netD / GANLoss / VGGLoss objects.  VGG19 weights: torchvision's ImageNet checkpoint is not reachable (no
    # network) -> vgg19(weights=None) loaded with GO.init_vgg_state_dict(0) under the reference's slice names.
    import torchvision
    _vgg19 = torchvision.models.vgg19
    torchvision.models.vgg19 = lambda pretrained=False, **kw: _vgg19(weights=None)
    from models.pix2pix_model import Pix2PixModel
    from models.networks.discriminator import MultiscaleDiscriminator
    from models.networks.loss import GANLoss, VGGLoss
    from models.networks.architecture import VGG19
    ndf = 16
    optd = argparse.Namespace(ndf=ndf, norm_D="spectralinstance", label_nc=3, output_nc=3, num_D=2, n_layers_D=4, netD_subarch="n_layer",
                              no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[])
    D = MultiscaleDiscriminator(optd).eval()
    D.load_state_dict(GO.init_discriminator_state_dict(0, ndf))
    vgg = VGG19().eval()
    vgg.load_state_dict(GO.init_vgg_state_dict(0, p=""))
    vl = VGGLoss.__new__(VGGLoss)                                         # VGGLoss.__init__ hard-codes .cuda(); forward is the reference's
    torch.nn.Module.__init__(vl)
    vl.vgg, vl.criterion, vl.weights = vgg, torch.nn.L1Loss(), [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]
    gen = torch.Generator().manual_seed(5)
    guide = torch.rand(1, 3, 128, 256, generator=gen) * 2
    fake = torch.rand(1, 3, 128, 256, generator=gen) * 50 * torch.rand(1, 1, 128, 256, generator=gen) ** 4
    real = torch.rand(1, 3, 128, 256, generator=gen) * 50 * torch.rand(1, 1, 128, 256, generator=gen) ** 4
    mask = (torch.rand(1, 1, 128, 256, generator=gen) > 0.3).float()
    m = types.SimpleNamespace(opt=optd, FloatTensor=torch.FloatTensor, netD=D, criterionVGG=vl, criterionFeat=torch.nn.L1Loss(),
                              criterionGAN=GANLoss("hinge", tensor=torch.FloatTensor, opt=optd), generate_fake=lambda i, c: fake)
    for n in ("discriminate", "divide_pred"):
        setattr(m, n, types.MethodType(getattr(Pix2PixModel, n), m))
    with torch.no_grad():
        gl, _ = Pix2PixModel.compute_generator_loss(m, guide, None, real, mask)
        dl = Pix2PixModel.compute_discriminator_loss(m, guide, None, real)
        feats = D(torch.cat([torch.cat([guide, fake], 1), torch.cat([guide, real], 1)], 0))
        vf = vgg(fake)
    gold = {"ndf": np.int64(ndf), "sd_seed": np.int64(0), "vgg_seed": np.int64(0), "in_seed": np.int64(5)}
    for k, v in list(gl.items()) + list(dl.items()):
        gold["loss_" + k] = np.float64(float(v))
    for i, fl in enumerate(feats):
        for j, f in enumerate(fl):
            gold["d%d_%d" % (i, j)] = f.numpy()[:, ::max(1, f.shape[1] // 8), ::2, ::2].astype(np.float32) if j < 4 else f.numpy()
    for j, f in enumerate(vf):
        gold["vgg_%d_absmean" % j] = np.float64(f.double().abs().mean())
        gold["vgg_%d" % j] = f.numpy()[:, ::max(1, f.shape[1] // 8), ::4, ::4].astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "discriminator.npz"), **gold)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
