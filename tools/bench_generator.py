"""Throughput of the GenProjector generator forward (eval) at the reference size (ngf=64, 118.4 M params), with per-kernel-family timing."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import emlight_b200 as E
from oracle import genprojector_oracle as GO

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--ngf", type=int, default=64)
ap.add_argument("--precision", default="bf16x3")
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
opt = argparse.Namespace(ngf=args.ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                         num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
dev = torch.device("cuda:0")
G = E.SPADEGenerator(opt, precision=args.precision).to(dev).eval()
G.load_state_dict(GO.init_generator_state_dict(0, args.ngf))
g = torch.Generator().manual_seed(1)
guide = (torch.rand(args.batch, 3, 128, 256, generator=g) * 2).to(dev)
crop = torch.rand(args.batch, 3, 128, 128, generator=g).to(dev)
for _ in range(2):
    out = G(guide, crop)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    out = G(guide, crop)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
flops = 154.1e9 * args.batch * (args.ngf / 64) ** 2
print(json.dumps({"workload": "SPADEGenerator forward (eval)", "ngf": args.ngf, "batch": args.batch, "precision": args.precision,
                  "ms_per_step": ms, "maps_per_s": args.batch / ms * 1e3, "useful_TFLOPs": flops / ms / 1e9,
                  "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
