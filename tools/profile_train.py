"""Per-C-ABI-entry-point device time of one training step (forward + loss + backward), CUDA events around every call."""
import collections, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch
import emlight_b200 as E
from emlight_b200 import _lib
from train_regression_synthetic import synthetic_batch, train_step
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = E.DenseNet(n_anchors=128).to(dev).train()
opt = torch.optim.Adam(net.parameters(), lr=1e-4)
l2 = torch.nn.MSELoss(); sam = E.SamplesLoss("sinkhorn", p=2, blur=.025)
batch = synthetic_batch(B, 128, torch.Generator().manual_seed(1), dev)
for _ in range(2): train_step(net, sam, l2, opt, batch, 128, 1)
torch.cuda.synchronize()
lib = _lib.load(); rec = []
class Wrap:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        tag = self.name
        if self.name == "eml_conv_forward":
            p = a[0]; p = p._obj if hasattr(p, "_obj") else p
            tag += "[mode%d,Cin%s,Cout%s%s]" % (p.mode, "<=64" if p.C_in <= 64 else ">64", "<=64" if p.C_out <= 64 else ">64", ",stats" if p.stats else "")
        if self.name == "eml_wgrad_1x1":               # (dY, dy_pitch, N, x, x_pitch, C, scale, shift, relu, pool, H, W, dW, M, precision, stream)
            C, pool, W = int(a[5]), int(a[9]), int(a[11])
            tag += "[W%d,%s%s]" % (W, "C<=256" if C <= 256 else "C>256", ",pool" if pool else "")
        if self.name in ("eml_bn_bwd_reduce", "eml_bn_bwd_apply"):   # (grad, g_pitch, x, x_pitch, pa, pb, mean, inv, gamma, beta, relu, pool, H, W, M, C, ...)
            C, pool, W = int(a[15]), int(a[11]), int(a[13])
            tag += "[W%d,%s%s]" % (W, "C<=48" if C <= 48 else "C>48", ",pool" if pool else "")
        if self.name in ("eml_wgrad_3x3",):
            tag += "[W%d]" % int(a[11])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = self.fn(*a); e1.record(); rec.append((tag, e0, e1)); return r
names = [n for n in _lib.SIGNATURES if n not in ("eml_version", "eml_error_string", "eml_device_ok", "eml_conv_wpack_bytes", "eml_sinkhorn_workspace_bytes")]
orig = {n: getattr(lib, n) for n in names}
for n in names: setattr(lib, n, Wrap(n, orig[n]))
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record(); c0 = time.perf_counter(); train_step(net, sam, l2, opt, batch, 128, 1); c1 = time.perf_counter(); t1.record()
torch.cuda.synchronize()
agg = collections.Counter(); cnt = collections.Counter()
for n, a, b in rec: agg[n] += a.elapsed_time(b); cnt[n] += 1
print("B=%d step %.1f ms device, %.1f ms host enqueue, %d C-ABI calls, sum of kernels %.1f ms" % (B, t0.elapsed_time(t1), (c1 - c0) * 1e3, len(rec), sum(agg.values())))
for n, v in agg.most_common(): print("  %-52s x%4d %9.2f ms" % (n, cnt[n], v))
