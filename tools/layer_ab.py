"""A/B of dense-layer kernel variants inside ONE process (interleaved repetitions, so that the board's power-capped clock affects every
variant alike): per-layer median ms at B = 256.   python tools/layer_ab.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import emlight_b200 as E

VARIANTS = [("smemA_r1", {"EML_DENSE_SMEM_A": "1"}), ("zstencil_all", {"EML_DENSE_RS_MAX_C": "0"}), ("rowsum<=96", {"EML_DENSE_RS_MAX_C": "96"}),
            ("rowsum<=144", {"EML_DENSE_RS_MAX_C": "144"}), ("rowsum_all", {"EML_DENSE_RS_MAX_C": "400"})]
if os.environ.get("EML_AB_SET") == "planes":          # block-1 slab layout (NHWC records vs channel planes) x epilogue choice
    VARIANTS = [("records", {"EML_DENSE_PLANES": "0"}), ("planes", {}), ("planes_zst", {"EML_DENSE_RS_MAX_C": "0"}),
                ("planes_rs<=96", {"EML_DENSE_RS_MAX_C": "96"}), ("planes_rs<=144", {"EML_DENSE_RS_MAX_C": "144"})]
KEYS = sorted({k for _, env in VARIANTS for k in env})
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = E.DenseNet(n_anchors=128).to(dev).eval()
x = torch.rand(256, 3, 192, 256, device=dev)
acc = {name: {} for name, _ in VARIANTS}
with torch.no_grad():
    for _ in range(2):
        net(x)
    for rep in range(7):
        for name, env in VARIANTS:
            for k in KEYS:
                os.environ.pop(k, None)
            os.environ.update(env)
            net.launch_log = []
            net(x)
            torch.cuda.synchronize()
            for fam, lname, abytes, flops, a, b in net.launch_log:
                acc[name].setdefault(lname, []).append(a.elapsed_time(b))
            net.launch_log = None
names = [n for n in acc[VARIANTS[0][0]] if all(n in acc[v] for v, _ in VARIANTS)]      # layers every variant runs as the same family
print("%-14s" % "layer" + "".join("%16s" % n for n, _ in VARIANTS))
tot = {n: 0.0 for n, _ in VARIANTS}
for ln in names:
    row = "%-14s" % ln
    for n, _ in VARIANTS:
        ts = sorted(acc[n][ln]); m = ts[len(ts) // 2]; tot[n] += m
        row += "%16.3f" % m
    print(row)
print("%-14s" % "sum" + "".join("%16.2f" % tot[n] for n, _ in VARIANTS))
