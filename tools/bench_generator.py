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
ap.add_argument("--profile", action="store_true")
ap.add_argument("--graph", action="store_true")
ap.add_argument("--eager", action="store_true", help="also time the reference ops (CPU restatement run on cuda: grid_sample + conv2d, fp32, TF32 off)")
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
if args.profile:
    # per-entry-point device time of one forward, measured with CUDA events around every C-ABI call (no syncs inside)
    from emlight_b200 import _lib
    lib = _lib.load()
    rec = []
    import ctypes
    class Wrap:
        def __init__(self, name, fn): self.name, self.fn = name, fn
        def __call__(self, *a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            key = self.name
            if self.name == "eml_im2col_lut_bf16": key += "[C=%d,px=%d]" % (a[2], a[12])          # channels, output pixels per image
            if self.name == "eml_gemm_bf16": key += "[M=%d,K=%d,N=%d]" % (a[2], a[3], a[5])
            e0.record(); r = self.fn(*a); e1.record(); rec.append((key, e0, e1)); return r
    names = [n for n in _lib.SIGNATURES if n not in ("eml_version", "eml_error_string", "eml_device_ok", "eml_conv_wpack_bytes", "eml_sinkhorn_workspace_bytes")]
    orig = {n: getattr(lib, n) for n in names}
    for n in names: setattr(lib, n, Wrap(n, orig[n]))
    t0 = time.perf_counter(); G(guide, crop); t_cpu = time.perf_counter() - t0
    torch.cuda.synchronize()
    for n in names: setattr(lib, n, orig[n])
    import collections
    agg = collections.Counter(); cnt = collections.Counter()
    for n, a, b in rec: agg[n] += a.elapsed_time(b); cnt[n] += 1
    print("profile: host time to enqueue one forward %.2f ms, %d C-ABI calls" % (t_cpu * 1e3, len(rec)))
    fam = collections.Counter()
    for n, v in agg.items(): fam[n.split("[")[0]] += v
    for n, v in fam.most_common(): print("  %-44s       %8.3f ms" % (n, v))
    print("  by shape:")
    for n, v in agg.most_common(40): print("  %-44s x%4d %8.3f ms" % (n, cnt[n], v))
if args.graph:
    G.use_cuda_graph = True
    for _ in range(3): out = G(guide, crop)
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
if args.eager:
    # PyTorch eager on the same GPU: the reference's formulation (9x grid_sample blow-up + stride-3 conv2d, cuDNN / cuBLAS fp32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(dev) for k, v in GO.init_generator_state_dict(0, args.ngf).items()}
    with torch.no_grad():
        for _ in range(2):
            ref = GO.generator_forward(sd, guide, crop, args.ngf)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ref = GO.generator_forward(sd, guide, crop, args.ngf)
        e1.record(); torch.cuda.synchronize()
    ems = e0.elapsed_time(e1) / args.steps
    err = float((out - ref).abs().max() / ref.abs().max())
    print(json.dumps({"workload": "the same forward, PyTorch eager (reference ops) on cuda:0", "ms_per_step": ems, "maps_per_s": args.batch / ems * 1e3,
                      "ours_over_eager": ems / ms, "max_rel_diff_ours_vs_eager": err, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))

