"""Median device time of the eval forward (configs[2] encoder part) at batch B over 10 runs.  Usage: python tools/fwd_time.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import emlight_b200 as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = E.DenseNet(n_anchors=128).to(dev).eval()
x = torch.rand(B, 3, 192, 256, device=dev)
ts = []
with torch.no_grad():
    for _ in range(3):
        net(x)
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); net(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ts.sort()
print("B=%d forward: median %.3f ms, min %.3f ms" % (B, ts[5], ts[0]))
