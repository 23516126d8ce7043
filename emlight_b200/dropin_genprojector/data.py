"""`import data; data.create_dataloader(opt)` (GenProjector/train.py, test.py) -- the Laval dataset of GenProjector/data.py:15-110 on
the wire formats of `emlight_b200.wire`, with the device work on the sm_100a kernels:

    <dataroot>/pkl/<name>.pickle     {distribution (128,), intensity, rgb_ratio (3,), ambient (3,)}   (RegressionNetwork/test.py:79-85 +
                                     distribution_representation.py)
    <dataroot>/warped/<name>.exr     ground-truth 128x256 HDR panorama;   <dataroot>/crop/<name>.exr   HDR crop

Item dictionary as in the reference (:104-107): 'input' = Gaussian-map guide rendered from the parameters (`genprojector_guide`,
data.py:86-102 in one launch), 'crop' = tone-mapped crop resized to 128x128 (3,128,128), 'warped' = panorama * alpha (3,128,256),
'map' = [intensity > 5 % of the maximum] (1,128,256), 'distribution', 'intensity' (the reference's repeated (1,128,3) views), 'name'."""
import os
import pickle

import numpy as np
import torch
import torch.utils.data

from emlight_b200 import wire
from emlight_b200.panorama import genprojector_guide
from emlight_b200.tonemap import TonemapHDR


def light_mask(hdr):
    """data.py:75-80: (1,H,W) float mask of the pixels brighter than 5 % of the brightest (0.3 R + 0.59 G + 0.11 B)."""
    inten = 0.3 * hdr[..., 0] + 0.59 * hdr[..., 1] + 0.11 * hdr[..., 2]
    return torch.from_numpy((inten > inten.max() * 0.05)[None].astype("uint8")).float()


class LavalIndoorDataset():
    def __init__(self, opt):
        self.opt = opt
        self.pairs = self.get_paths(opt)
        self.dataset_size = len(self.pairs)
        self.tone = TonemapHDR(gamma=2.4, percentile=50, max_mapping=0.5)

    def get_paths(self, opt):
        pkl_dir = os.path.join(opt.dataroot, "pkl")
        pairs = []
        for nm in sorted(os.listdir(pkl_dir)):
            if nm.endswith(".pickle"):
                warped_path = os.path.join(opt.dataroot, "warped", nm.replace("pickle", "exr"))
                if os.path.exists(warped_path):
                    pairs.append([os.path.join(pkl_dir, nm), warped_path])
        return pairs

    def __getitem__(self, index):
        ln = 128
        pkl_path, warped_path = self.pairs[index]
        with open(pkl_path, "rb") as handle:
            pkl = pickle.load(handle)
        dev = torch.device("cuda")
        crop = torch.from_numpy(wire.load_exr(warped_path.replace("warped", "crop"))).to(dev)
        crop, alpha = self.tone(crop)                                                           # (H,W,3) in [0,1]
        alpha = float(alpha)
        crop = torch.nn.functional.interpolate(crop.permute(2, 0, 1)[None], size=(128, 128), mode="bilinear", align_corners=False)[0]
        hdr = wire.load_exr(warped_path)
        warped = torch.from_numpy(np.ascontiguousarray(np.transpose(hdr, (2, 0, 1)))) * alpha
        dist = torch.from_numpy(np.asarray(pkl["distribution"])).float().to(dev)
        inten = torch.from_numpy(np.array(pkl["intensity"])).float().to(dev)
        rgb = torch.from_numpy(np.array(pkl["rgb_ratio"])).float().to(dev)
        amb = torch.from_numpy(np.asarray(pkl["ambient"])).float().to(dev)
        env = genprojector_guide(dist.view(1, ln), inten.view(1), rgb.view(1, 3), amb.view(1, 3), alpha=alpha)[0]
        return {"input": env, "crop": crop.contiguous(), "warped": warped, "map": light_mask(hdr),
                "distribution": dist.view(1, ln, 1).repeat(1, 1, 3), "intensity": (inten * 0.01).view(1, 1, 1).repeat(1, ln, 3),
                "name": os.path.basename(pkl_path).split(".")[0]}

    def __len__(self):
        return self.dataset_size


def create_dataloader(opt):
    return torch.utils.data.DataLoader(LavalIndoorDataset(opt), batch_size=opt.batchSize, shuffle=not opt.serial_batches,
                                       num_workers=0, drop_last=opt.isTrain)      # items are produced on the GPU: no worker processes
