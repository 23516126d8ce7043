"""Per-layer device time of the eval forward at B = 256 (CUDA events around every conv-family launch; `DenseNet.launch_log`):
name, C_in-derived algorithmic bytes, ms, GB/s.  Usage: python tools/layer_times.py [batch]"""
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
with torch.no_grad():
    for _ in range(3):
        net(x)
    acc = {}
    for rep in range(5):
        net.launch_log = []
        net(x)
        torch.cuda.synchronize()
        for fam, name, abytes, flops, a, b in net.launch_log:
            acc.setdefault((fam, name), [abytes, []])[1].append(a.elapsed_time(b))
    net.launch_log = None
tot = 0.0
for (fam, name), (abytes, ts) in acc.items():
    ms = sorted(ts)[len(ts) // 2]
    tot += ms
    print("%-12s %-14s %8.1f MB  %7.3f ms  %7.1f GB/s" % (fam, name, abytes / 1e6, ms, abytes / ms / 1e6))
print("sum of medians %.2f ms" % tot)
