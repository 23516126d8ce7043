"""Debug helper (not collected by pytest): per-parameter gradient error of the SPADE generator tape vs autograd of the oracle,
for each precision mode.  Usage on a GPU box:  python tests/debug_gp_bwd.py [ngf] [batch]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import genprojector_oracle as GO  # noqa: E402


def main():
    import emlight_b200 as E
    ngf = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    Bn = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cuda = torch.device("cuda:0")
    opt = argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                             num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
    sd0 = GO.init_generator_state_dict(seed=4, ngf=ngf)
    gen = torch.Generator().manual_seed(9)
    guide = torch.rand(Bn, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(Bn, 3, 96, 112, generator=gen)
    gout = torch.randn(Bn, 3, 128, 256, generator=gen)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))
              else v.clone()) for k, v in sd0.items()}
    ref = GO.generator_forward(sd, guide, crop, ngf=ngf, upd={})
    (ref * gout).sum().backward()
    sd64 = {k: (v.double().clone().requires_grad_(True) if v.is_floating_point() and not k.endswith(("weight_u", "weight_v", "running_mean", "running_var"))
                else (v.double().clone() if v.is_floating_point() else v.clone())) for k, v in sd0.items()}
    try:
        ref64 = GO.generator_forward(sd64, guide.double(), crop.double(), ngf=ngf, upd={})
        (ref64 * gout.double()).sum().backward()
    except Exception as e:      # noqa: BLE001
        print("fp64 oracle failed:", e)
        sd64 = None
    if len(sys.argv) > 3 and sys.argv[3] == "warm":      # mimic the tests that run before this one in tests/test_gp_train_gpu.py
        sc = E.SphereConv2D(6, 5, stride=2).to(cuda)
        sc.autograd = True
        xw = torch.randn(2, 6, 16, 32).to(cuda).requires_grad_(True)
        sc(xw).sum().backward()
    from emlight_b200 import gp_train
    for prec, ovr in (("bf16x3", {}), ("bf16x3", {"fwd": "fp32"}), ("bf16x3", {"bwd": "fp32"}), ("fp32", {})):
        gp_train.PRECISION_OVERRIDE.clear()
        gp_train.PRECISION_OVERRIDE.update(ovr)
        G = E.SPADEGenerator(opt).to(cuda).train()
        G.load_state_dict(sd0)
        G.autograd = True
        G.precision = prec
        for m in G.modules():
            if hasattr(m, "precision"):
                m.precision = prec
        out = G(guide.to(cuda), crop.to(cuda))
        (out * gout.to(cuda)).sum().backward()
        print("==== precision", prec, "override", ovr, "fwd err", float((out.detach().cpu() - ref.detach()).abs().max()) / 50.0)
        for name, p in G.named_parameters():
            want = sd[name].grad
            e32 = float((p.grad.cpu() - want).norm()) / max(float(want.norm()), 1e-30)
            line = "%-48s |g| %10.4g  err_vs_fp32oracle %.3e" % (name, float(want.norm()), e32)
            if sd64 is not None:
                w64 = sd64[name].grad
                line += "  err_vs_fp64 %.3e  fp32oracle_vs_fp64 %.3e" % (float((p.grad.cpu().double() - w64).norm()) / max(float(w64.norm()), 1e-30),
                                                                       float((want.double() - w64).norm()) / max(float(w64.norm()), 1e-30))
            print(line)


if __name__ == "__main__":
    main()
