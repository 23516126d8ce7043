#!/usr/bin/env python
"""EMLight's two-stage inference in ONE process on synthetic data:

    HDR crop --TonemapHDR--> LDR crop, alpha --DenseNet--> distribution / intensity / rgb_ratio / ambient
             --genprojector_guide--> Gaussian-map panorama --SPADEGenerator(guide, crop)--> 128x256 HDR illumination map

The reference does this with two programs and a directory of pickles in between (RegressionNetwork/test.py:79-85 writes
{distribution, rgb_ratio, intensity*500}; GenProjector/data.py:64-102 reads them, scales intensity by 0.01 and renders the guide).
Scale bookkeeping: the regression targets are intensity*alpha/500 and ambient*alpha/(128*256) (RegressionNetwork/data.py:68-71), so
with predictions the guide is  render(dist * (pred_intensity*500*0.01) * rgb) + pred_ambient  -- alpha is already inside.
Random-initialised networks: this demonstrates the data flow and its throughput, not image quality.

    python examples/pipeline_synthetic.py --batch 16 --ngf 64
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(batch=4, ngf=16, n_anchors=128, steps=1, device="cuda:0", seed=0):
    import torch
    import emlight_b200 as E
    from emlight_b200.tonemap import TonemapHDR
    dev = torch.device(device)
    torch.manual_seed(seed)
    net = E.DenseNet(n_anchors=n_anchors).to(dev).eval()
    opt = argparse.Namespace(ngf=ngf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                             num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0)
    G = E.SPADEGenerator(opt).to(dev).eval()
    tone = TonemapHDR(gamma=2.4, percentile=50, max_mapping=0.5)
    gen = torch.Generator().manual_seed(seed + 1)
    hdr_crop = torch.exp(torch.randn(batch, 192, 256, 3, generator=gen) - 1.0).to(dev)          # (B,H,W,3) radiance like load_exr
    out = None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        with torch.no_grad():
            ldr, alpha = tone(hdr_crop)                                                        # util.py:36-66
            crop = ldr.permute(0, 3, 1, 2).contiguous()                                        # to_tensor: (B,3,192,256) in [0,1]
            pred = net(crop)                                                                   # DenseNet.py:135-157
            dist = torch.softmax(pred["distribution"], 1)                                      # a valid distribution for the demo
            rgb = torch.nn.functional.normalize(pred["rgb_ratio"].abs() + 1e-3, dim=1)
            guide = E.genprojector_guide(dist, pred["intensity"].abs() * 500.0, rgb, pred["ambient"].abs() * (128 * 256), alpha=1.0)
            out = G(guide, crop)                                                               # generator.py:65-88 (crop resized inside)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"guide": guide, "output": out, "alpha": alpha, "pred": pred, "maps_per_s": batch / dt}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--ngf", type=int, default=16)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    r = run(a.batch, a.ngf, steps=a.steps)
    print("guide %s, output %s in [%.2f, %.2f], %.1f maps/s end to end (crop -> illumination map)" % (
        tuple(r["guide"].shape), tuple(r["output"].shape), float(r["output"].min()), float(r["output"].max()), r["maps_per_s"]))
