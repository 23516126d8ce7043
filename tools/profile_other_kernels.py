"""Runs the non-DenseNet hot-path kernels once each at BASELINE sizes so that ncu can capture them:
Sinkhorn (B=256, N=128), SG render forward (B=256, N=128), needlet projection GEMM (B=64), generator forward (B=2, ngf=64)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import emlight_b200 as E

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="all")
args = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
if args.what in ("all", "sinkhorn"):
    B, N = 256, 128
    x = (0.3 * torch.randn(B, N, 1, generator=g)).to(dev).requires_grad_()
    y = torch.softmax(3 * torch.randn(B, N, generator=g), 1).view(B, N, 1).to(dev)
    loss = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
    for _ in range(3):
        loss(x, y).sum().backward()
if args.what in ("all", "render"):
    B, N = 256, 128
    dist = torch.softmax(3 * torch.randn(B, N, generator=g), 1).to(dev)
    inten = torch.rand(B, 1, generator=g).to(dev)
    rgb = torch.rand(B, 3, generator=g).to(dev)
    for _ in range(3):
        E.render_from_params(dist, inten, rgb)
if args.what in ("all", "needlets"):
    from emlight_b200.needlets import NeedletTransform
    nt = NeedletTransform(jmax=3, device=dev)
    p = torch.rand(64, 3, 128, 256, generator=g).to(dev)
    for _ in range(2):
        nt.reconstruct(nt.project(p))
if args.what in ("all", "generator"):
    from oracle import genprojector_oracle as GO
    opt = argparse.Namespace(ngf=64, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                             num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
    G = E.SPADEGenerator(opt, precision="bf16x3").to(dev).eval()
    G.load_state_dict(GO.init_generator_state_dict(0, 64))
    guide = (torch.rand(2, 3, 128, 256, generator=g) * 2).to(dev)
    crop = torch.rand(2, 3, 128, 128, generator=g).to(dev)
    with torch.no_grad():
        for _ in range(2):
            G(guide, crop)
torch.cuda.synchronize()
print("done")
