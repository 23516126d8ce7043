"""`import data; data.ParameterDataset(train_dir)` (RegressionNetwork/train.py:7,33; test.py:6,25) -- the regression dataset of
RegressionNetwork/data.py:20-86 (whose own file carries unresolved merge markers), on the wire formats of `emlight_b200.wire`:

    <train_dir>/pkl/<name>.pickle   {distribution (N,), intensity, rgb_ratio (3,), ambient (3,)}   (distribution_representation.py)
    <train_dir>/crop/<name>.exr     HDR crop

`__getitem__` returns the reference's dictionary: 'crop' = ToTensor(TonemapHDR(2.4, 50, 0.5)(exr)) -> (3,H,W) in [0,1],
'distribution', 'intensity' * alpha / 500, 'rgb_ratio', 'ambient' * alpha / (128*256), 'name' (data.py:62-73).  The tone mapping runs
on the GPU kernel, so items are CUDA tensors (the reference's DataLoader has no worker processes: train.py:35); `.to(device)` in the
training loop is then a no-op."""
import os
import pickle

import numpy as np
import torch
from torch.utils.data import Dataset

from emlight_b200 import wire
from emlight_b200.tonemap import TonemapHDR


class ParameterDataset(Dataset):
    def __init__(self, train_dir, device="cuda"):
        assert os.path.exists(train_dir)
        gt_dir, crop_dir = os.path.join(train_dir, "pkl"), os.path.join(train_dir, "crop")
        self.pairs = []
        for nm in sorted(os.listdir(gt_dir)):
            if nm.endswith("pickle"):
                crop_path = os.path.join(crop_dir, nm.replace("pickle", "exr"))
                if os.path.exists(crop_path):
                    self.pairs.append([crop_path, os.path.join(gt_dir, nm)])
        self.data_len = len(self.pairs)
        self.device = torch.device(device)
        self.tone = TonemapHDR(gamma=2.4, percentile=50, max_mapping=0.5)

    def __getitem__(self, index):
        crop_path, gt_path = self.pairs[index]
        exr = torch.from_numpy(wire.load_exr(crop_path)).to(self.device)                     # (H,W,3) radiance
        ldr, alpha = self.tone(exr)
        alpha = float(alpha)
        with open(gt_path, "rb") as handle:
            gt = pickle.load(handle)
        dev = self.device
        return {
            "crop": ldr.permute(2, 0, 1).contiguous(),                                        # transforms.ToTensor on a float HWC array
            "distribution": torch.from_numpy(np.asarray(gt["distribution"])).float().to(dev),
            "intensity": torch.from_numpy(np.array(gt["intensity"])).float().to(dev) * alpha / 500,
            "rgb_ratio": torch.from_numpy(np.asarray(gt["rgb_ratio"])).float().to(dev),
            "ambient": torch.from_numpy(np.asarray(gt["ambient"])).float().to(dev) * alpha / (128 * 256),
            "name": os.path.basename(gt_path).split(".pickle")[0],
        }

    def __len__(self):
        return self.data_len
