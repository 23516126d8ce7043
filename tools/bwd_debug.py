"""Per-parameter gradient error table: emlight_b200.DenseNet backward vs torch autograd through the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import emlight_b200 as E
from oracle import densenet_oracle as DO
KEYS = ("distribution", "intensity", "rgb_ratio", "ambient")
dev = torch.device("cuda:0")
sd = DO.init_state_dict(seed=0, n_anchors=96)
x = torch.rand(2, 3, 192, 256, generator=torch.Generator().manual_seed(21))
gen = torch.Generator().manual_seed(22)
R = {k: torch.randn(2, n, generator=gen) for k, n in (("distribution", 96), ("intensity", 1), ("rgb_ratio", 3), ("ambient", 3))}
sdo = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
taps = {}
out = DO.densenet_forward(sdo, x, training=True, taps=taps)
for k in ("trans3", "features", "pooled", "block3"): taps[k].retain_grad()
sum((out[k] * R[k]).sum() for k in KEYS).backward()
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
net = E.DenseNet(precision=prec).to(dev); net.load_state_dict(sd); net.train(); net._debug = {}
o = net(x.to(dev))
for k in KEYS: print("fwd", k, float((o[k].cpu() - out[k].detach()).abs().max() / out[k].detach().abs().max()))
sum((o[k] * R[k].to(dev)).sum() for k in KEYS).backward()
D = net._debug
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-20))
t_ref = taps["trans3"].detach().permute(0, 2, 3, 1)
print("t_last fwd   ", rel(D["t_last"].cpu()[..., :171], t_ref))
# d(relu output): oracle grad of relu(features) is grad of avg-pool input; reconstruct from pooled grad
gp = taps["pooled"].grad.view(2, 171, 6, 8).permute(0, 2, 3, 1)
dz_ref = gp.repeat_interleave(4, 1).repeat_interleave(4, 2) / 16
print("dz           ", rel(D["dz"].cpu(), dz_ref))
print("dt_last      ", rel(D["dt_last"].cpu()[..., :171], taps["trans3"].grad.permute(0, 2, 3, 1)))
mask_ref = (taps["features"].detach() > 0).permute(0, 2, 3, 1)
u = (D["t_last"].cpu()[..., :171] - D["ln_mean"].cpu()) * D["ln_inv"].cpu() * sd["features.last_norm3.weight"] + sd["features.last_norm3.bias"]
print("mask mismatches", int(((u > 0) != mask_ref).sum()), "of", mask_ref.numel())
tm = t_ref.reshape(-1, 171)
print("ln_mean      ", rel(D["ln_mean"].cpu(), tm.mean(0)), " ln_inv", rel(D["ln_inv"].cpu(), torch.rsqrt(tm.var(0, unbiased=False) + 1e-5)))
gfeat = taps["features"].grad.permute(0, 2, 3, 1)          # grad wrt last_norm3 output = dz * mask
print("dz*mask      ", rel(D["dz"].cpu() * (u > 0), gfeat))
import numpy as np
rows = []
for name, p in net.named_parameters():
    ref = sdo[name].grad; got = p.grad.cpu()
    emax = float((got - ref).abs().max() / (ref.abs().max() + 1e-12))
    el2 = float((got - ref).norm() / (ref.norm() + 1e-12))
    rows.append((emax, el2, name, float(ref.abs().max())))
gmax = max(r[3] for r in rows)
sig = [r for r in rows if r[3] > 1e-4 * gmax]            # last_norm{1,2}.* are analytically ~0 (BN of a BN input): noise only
print("precision", prec, "params", len(rows), "significant", len(sig))
print("max-rel : median %.2e p90 %.2e max %.2e" % (np.median([r[0] for r in sig]), np.percentile([r[0] for r in sig], 90), max(r[0] for r in sig)))
print("L2-rel  : median %.2e p90 %.2e max %.2e" % (np.median([r[1] for r in sig]), np.percentile([r[1] for r in sig], 90), max(r[1] for r in sig)))
for r in sorted(sig, key=lambda r: -r[1])[:6]: print("  worst L2 %-58s L2 %.2e max %.2e |ref| %.2e" % (r[2], r[1], r[0], r[3]))
for r in rows:
    if r not in sig: print("  negligible %-50s |ref| %.2e abs err/gmax %.2e" % (r[2], r[3], r[0] * r[3] / gmax))
