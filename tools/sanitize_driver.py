"""Small-shape pass over the hot kernels for compute-sanitizer (tools/sanitize.sh).  Product code only -- no oracle, no checks beyond
finiteness: the sanitizer's report is the result.   python tools/sanitize_driver.py <what> [--poison]

  densenet_eval    B=2 eval forward: stem, 47 fused dense layers (pair mode in block 3), transitions, head, fc, heads + SG render
  densenet_train   B=2 train-mode forward + backward (stats epilogues, dgrad convs, tcgen05 wgrads, BN backward)
  sinkhorn         SamplesLoss forward + backward, B=4, N=128
  gemm             eml_gemm_bf16 / split-K through gp_ops.mm_nt at ragged sizes
  generator_train  SPADE generator ngf=4, B=2: train forward on the tape + backward (im2col_t, split-K GEMMs, col2im_csr, SPADE / BN adjoints)
  gan_step         Pix2PixModel G step + D step, ngf=ndf=4, B=1
--poison fills the caching allocator's free blocks with NaN first, so that a kernel reading memory it never wrote shows up as NaN."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))


def poison(dev):
    big = torch.full((1 << 29,), float("nan"), device=dev)            # 2 GiB of the large pool
    small = [torch.full((1 << 17,), float("nan"), device=dev) for _ in range(512)]   # 512 x 512 KiB of the small pool
    torch.cuda.synchronize()
    del big, small


def finite(name, t):
    ok = bool(torch.isfinite(t).all())
    print("%-60s %s" % (name, "finite" if ok else "NOT FINITE (%d bad of %d)" % (int((~torch.isfinite(t)).sum()), t.numel())))
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what")
    ap.add_argument("--poison", action="store_true")
    a = ap.parse_args()
    import emlight_b200 as E
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1)
    bad = 0
    if a.poison:
        poison(dev)
    if a.what in ("densenet_eval", "all"):
        net = E.DenseNet(n_anchors=128).to(dev).eval()
        x = torch.rand(2, 3, 192, 256, generator=g).to(dev)
        with torch.no_grad():
            o = net(x)
            pano = E.render_from_params(o["distribution"], o["intensity"], o["rgb_ratio"])
        bad += not all(finite("densenet_eval." + k, v) for k, v in o.items())
        bad += not finite("render", pano)
    if a.what in ("densenet_train", "all"):
        net = E.DenseNet(n_anchors=128).to(dev).train()
        x = torch.rand(2, 3, 192, 256, generator=g).to(dev)
        o = net(x)
        sum(v.sum() for v in o.values()).backward()
        for n, p in net.named_parameters():
            if not torch.isfinite(p.grad).all():
                bad += not finite("densenet_train.grad." + n, p.grad)
        print("densenet_train: %d parameter gradients checked" % len(list(net.parameters())))
    if a.what in ("sinkhorn", "all"):
        xs = (0.3 * torch.randn(4, 128, 1, generator=g)).to(dev).requires_grad_()
        ys = torch.softmax(3 * torch.randn(4, 128, generator=g), 1).view(4, 128, 1).to(dev)
        loss = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=4)(xs, ys)
        loss.sum().backward()
        bad += not (finite("sinkhorn.loss", loss) and finite("sinkhorn.grad", xs.grad))
    if a.what in ("gemm", "all"):
        from emlight_b200 import gp_ops
        for M, N, K in ((36, 3, 4160), (300, 70, 1000), (128, 260, 64), (2, 520, 130)):
            r = gp_ops.mm_nt(torch.randn(M, K, generator=g).to(dev), torch.randn(N, K, generator=g).to(dev))
            bad += not finite("mm_nt %dx%dx%d" % (M, N, K), r)
    if a.what in ("generator_train", "all"):
        opt = argparse.Namespace(ngf=4, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                                 num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
        G = E.SPADEGenerator(opt).to(dev).train()
        G.autograd = True
        guide = (torch.rand(2, 3, 128, 256, generator=g) * 2).to(dev)
        crop = torch.rand(2, 3, 96, 112, generator=g).to(dev)
        out = G(guide, crop)
        bad += not finite("generator_train.out", out)
        (out * torch.randn(out.shape, generator=g).to(dev)).sum().backward()
        for n, p in G.named_parameters():
            if p.grad is None or not torch.isfinite(p.grad).all():
                bad += 1
                finite("generator_train.grad." + n, p.grad if p.grad is not None else torch.tensor(float("nan")))
        print("generator_train: %d parameter gradients checked" % len(list(G.parameters())))
    if a.what in ("gan_step", "all"):
        from train_genprojector_synthetic import synthetic_batch as gan_batch
        gopt = argparse.Namespace(ngf=4, ndf=4, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", norm_D="spectralinstance",
                                  semantic_nc=3, label_nc=3, output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0,
                                  num_D=2, n_layers_D=4, netD_subarch="n_layer", no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0],
                                  isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gm = E.Pix2PixModel(gopt)
        gm.train()
        gm.autograd = True
        og, od = gm.create_optimizers(gopt)
        gd = gan_batch(1, g, dev)
        og.zero_grad(); gl, _ = gm(gd, "generator"); sum(gl.values()).mean().backward(); og.step()
        od.zero_grad(); dl = gm(gd, "discriminator"); sum(dl.values()).mean().backward(); od.step()
        for k, v in list(gl.items()) + list(dl.items()):
            bad += not finite("gan_step.loss." + k, v.detach())
    torch.cuda.synchronize()
    print("driver done: %d non-finite results" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
