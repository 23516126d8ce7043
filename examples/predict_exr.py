#!/usr/bin/env python
"""File-based inference with the reference's wire formats, mirroring RegressionNetwork/test.py:

    crop.exr --load_exr--> HDR crop --TonemapHDR--> LDR crop --DenseNet--> heads
             --> <name>.pickle {distribution, rgb_ratio, intensity*500}   (test.py:79-85; what GenProjector/data.py:64-94 reads)
             --> <name>_pano.exr   the spherical-Gaussian panorama of the prediction (util.convert_to_panorama + write_exr)

Without --crop a synthetic HDR crop is written first, so the script runs end to end on a box without data.  Random-initialised
network unless --weights points at a reference checkpoint (`latest_net.pth`, same state_dict keys).

    python examples/predict_exr.py --out /tmp/emlight_results
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(crop_path, out_dir, weights=None, n_anchors=128, device="cuda:0"):
    import numpy as np
    import torch
    import emlight_b200 as E
    from emlight_b200 import wire
    from emlight_b200.tonemap import TonemapHDR
    dev = torch.device(device)
    os.makedirs(out_dir, exist_ok=True)
    nm = os.path.splitext(os.path.basename(crop_path))[0]
    net = E.DenseNet(n_anchors=n_anchors).to(dev).eval()
    if weights:
        net.load_state_dict(torch.load(weights, map_location=dev))
    tone = TonemapHDR(gamma=2.4, percentile=99, max_mapping=0.9)                       # test.py:34
    hdr = torch.from_numpy(wire.load_exr(crop_path)).to(dev)[None]                     # (1,H,W,3)
    with torch.no_grad():
        ldr, alpha = tone(hdr)
        crop = torch.nn.functional.interpolate(ldr.permute(0, 3, 1, 2), size=(192, 256), mode="bilinear", align_corners=False)
        pred = net(crop.contiguous())
        intensity = pred["intensity"] * 500                                            # test.py:53
        dirs = torch.from_numpy(E.sphere_points(n_anchors)).float().view(1, n_anchors * 3).to(dev)
        size = torch.full((1, n_anchors), 0.0025, device=dev)
        color = (pred["distribution"].view(1, n_anchors, 1) * intensity.view(1, 1, 1) * pred["rgb_ratio"].view(1, 1, 3)).reshape(1, -1)
        pano = E.convert_to_panorama(dirs, size, color.contiguous())                   # (1,3,128,256)
    pkl = os.path.join(out_dir, nm + ".pickle")
    wire.save_parametric_lights(pkl, pred["distribution"][0].view(n_anchors), pred["rgb_ratio"][0].view(3), intensity[0])
    exr = os.path.join(out_dir, nm + "_pano.exr")
    wire.write_exr(exr, np.transpose(pano[0].cpu().numpy(), (1, 2, 0)))
    return pkl, exr


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--crop", default=None, help="HDR crop (.exr); a synthetic one is written when omitted")
    ap.add_argument("--out", default="./results")
    ap.add_argument("--weights", default=None)
    a = ap.parse_args()
    if a.crop is None:
        import numpy as np
        from emlight_b200 import wire
        os.makedirs(a.out, exist_ok=True)
        a.crop = os.path.join(a.out, "synthetic_crop.exr")
        wire.write_exr(a.crop, np.exp(np.random.default_rng(0).normal(-1.0, 1.0, (192, 256, 3))).astype(np.float32))
    print("wrote %s and %s" % run(a.crop, a.out, a.weights))
