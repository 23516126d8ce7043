#!/usr/bin/env python
"""Ground-truth parameter pickles from a directory of HDR panoramas -- the batch loop at the bottom of
RegressionNetwork/representation/distribution_representation.py:160-182 (`extract_mesh(ln=128).compute(hdr)` per .exr -> one pickle
each), with the extraction on the GPU kernel in batches and the files going through `emlight_b200.wire` (no OpenEXR bindings).

    python examples/make_gt_pickles.py --hdr-dir <dir of 128x256 .exr> --out-dir <pkl dir> [--ln 128] [--batch 64]

Without --hdr-dir a few synthetic panoramas are written first so that the script runs end to end on a box without data.
Each pickle holds {distribution (ln,), intensity, rgb_ratio (3,), ambient (3,)} as numpy float32, what data.py:64-71 reads.
"""
import argparse
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from emlight_b200 import wire
from emlight_b200.representation import extract_mesh


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hdr-dir", default=None)
    ap.add_argument("--out-dir", default="./pkl")
    ap.add_argument("--ln", type=int, default=128)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    os.makedirs(a.out_dir, exist_ok=True)
    if a.hdr_dir is None:
        a.hdr_dir = os.path.join(a.out_dir, "_synthetic_hdr")
        os.makedirs(a.hdr_dir, exist_ok=True)
        rng = np.random.default_rng(0)
        for i in range(5):
            pano = np.exp(rng.normal(-2.0, 1.0, (128, 256, 3))).astype(np.float32)
            y, x = rng.integers(8, 120), rng.integers(8, 248)
            pano[y - 3:y + 3, x - 5:x + 5] += rng.uniform(100, 500)
            wire.write_exr(os.path.join(a.hdr_dir, "pano%02d.exr" % i), pano)
    names = sorted(n for n in os.listdir(a.hdr_dir) if n.endswith(".exr"))
    extractor = extract_mesh(ln=a.ln)
    done = 0
    for i in range(0, len(names), a.batch):
        chunk = names[i:i + a.batch]
        hdr = torch.from_numpy(np.stack([wire.load_exr(os.path.join(a.hdr_dir, n)) for n in chunk])).cuda()      # (B,128,256,3)
        para, _ = extractor.compute(hdr)
        para = {k: v.cpu().numpy() for k, v in para.items()}
        for j, n in enumerate(chunk):
            rec = {k: (v[j] if v[j].ndim else np.float32(v[j])) for k, v in para.items()}
            with open(os.path.join(a.out_dir, n.replace("exr", "pickle")), "wb") as handle:
                pickle.dump(rec, handle, protocol=pickle.HIGHEST_PROTOCOL)
            done += 1
    print("wrote %d pickles to %s" % (done, a.out_dir))


if __name__ == "__main__":
    main()
