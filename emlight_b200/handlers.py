"""Host-side helpers of `RegressionNetwork/util.py` that the reference's scripts import next to the hot-path functions
(`from util import PanoramaHandler, TonemapHDR, tonemapping`, train.py:10 / test.py:10; `util.print_model_parm_nums`, train.py:53).

These are the callers' data-preparation utilities (numpy on the host, as in the reference), restated so that the module-name shim
`dropin/util.py` satisfies every name those scripts import; file reading goes through `emlight_b200.wire` (no OpenEXR / Imath), and
`tonemapping` -- the same arithmetic as `TonemapHDR(2.4, 99, 0.8)` -- runs on the tonemap kernel.  Pinned against the reference's own
code (exec'd from util.py:69-220) through `tests/golden/handlers.npz` (`tests/test_handlers_cpu.py`).
"""
import numpy as np

from . import wire


class PanoramaHandler(object):
    """util.py:69-186 -- static helpers on (H, W, 3) float panoramas."""

    @staticmethod
    def rgb_to_intenisty(rgbs):
        # util.py:74-76 (the third term reads channel 0 again, as in the reference)
        return 0.2126 * rgbs[..., 0] + 0.7152 * rgbs[..., 1] + 0.0722 * rgbs[..., 0]

    @staticmethod
    def read_exr(exr_path):
        """util.py:78-93: (hdr (H,W,3) float32, alpha (H,W) float32) of an RGBA OpenEXR file."""
        ch = wire.read_exr_channels(exr_path)
        missing = [c for c in "RGBA" if c not in ch]
        if missing:
            raise ValueError("OpenEXR file has no channel(s) %s" % missing)
        hdr = np.stack([ch[c].astype(np.float32) for c in "RGB"], axis=-1)
        return hdr, ch["A"].astype(np.float32)

    @staticmethod
    def read_hdr(hdr_path):
        """util.py:95-99 (`cv2.imread(..., IMREAD_UNCHANGED)[..., ::-1]`): RGB float32; .exr through `wire.load_exr`."""
        if str(hdr_path).lower().endswith(".exr"):
            return wire.load_exr(hdr_path)
        import cv2
        img = cv2.imread(hdr_path, flags=cv2.IMREAD_UNCHANGED | cv2.IMREAD_ANYCOLOR | cv2.IMREAD_ANYDEPTH)
        if img is None:
            raise FileNotFoundError(hdr_path)
        return img[..., ::-1]

    @staticmethod
    def horizontal_rotate_panorama(hdr_img, deg):
        return np.roll(hdr_img, shift=int(deg / 360.0 * hdr_img.shape[1]), axis=1)        # util.py:101-105

    @staticmethod
    def generate_steradian(height, width, multiply=True):
        """util.py:107-116: sin(latitude of the row centre), times the equirect pixel area when `multiply`."""
        row = np.sin((np.arange(height, dtype=np.float64) + 0.5) / height * np.pi)
        ster = np.repeat(row[:, None], width, axis=1)
        if multiply:
            ster = ster * (((2 * np.pi) / width) * ((1 * np.pi) / height))
        return ster.astype(np.float32)

    @staticmethod
    def prepare_gt_panorama(hdr_img, threshold=None):
        """util.py:118-136: pixels darker than max/20 become the steradian-weighted ambient term and are zeroed IN PLACE."""
        weight = PanoramaHandler.generate_steradian(hdr_img.shape[0], hdr_img.shape[1])
        inten = PanoramaHandler.rgb_to_intenisty(hdr_img)
        if threshold is None or threshold < 0.0:
            threshold = inten.max() / 20.
        dark = inten < threshold
        if dark.any():
            ambient = np.sum(hdr_img[dark] * weight[dark][:, None], axis=0, dtype=np.float32) / np.sum(weight[dark], dtype=np.float32)
        else:
            ambient = np.zeros([3], dtype=np.float32)
        hdr_img[dark] = 0.0
        return hdr_img, ambient

    @staticmethod
    def resize_panorama(hdr_img, new_shape):
        import cv2                                                                            # util.py:138-144 (INTER_AREA)
        if isinstance(new_shape, tuple) and len(new_shape) == 2:
            return cv2.resize(hdr_img, new_shape, interpolation=cv2.INTER_AREA)
        if isinstance(new_shape, int):
            return cv2.resize(hdr_img, (2 * new_shape, new_shape), interpolation=cv2.INTER_AREA)
        return hdr_img

    @staticmethod
    def crop_panorama(hdr_img, fov_deg, crop_image_h=720, crop_image_aspect_ratio="4:3"):
        """util.py:146-185: perspective crop looking at the panorama centre, bilinear lookup on the pixel grid."""
        from scipy import interpolate
        if hdr_img.dtype == np.uint8:
            hdr_img = hdr_img / 255.0
        num, den = [int(v) for v in crop_image_aspect_ratio.split(":")]
        ratio = num / den
        crop_w = int(crop_image_h * ratio)
        scl = np.tan(np.deg2rad(fov_deg) / 2)
        sx, sy = np.meshgrid(np.linspace(-scl, scl, crop_w), np.linspace(-scl / ratio, scl / ratio, crop_image_h))
        r = np.sqrt(sy * sy + sx * sx + 1)
        sx, sy = sx / r, sy / r
        sz = np.sqrt(1 - sy * sy - sx * sx)
        x = (1 + np.arctan2(sx, sz) / np.pi) / 2 * hdr_img.shape[1]
        y = (1 + np.arcsin(sy) / (np.pi / 2)) / 2 * hdr_img.shape[0]
        f = interpolate.RegularGridInterpolator((np.arange(0, hdr_img.shape[0]), np.arange(0, hdr_img.shape[1])), hdr_img)
        return f(np.c_[y.ravel(), x.ravel()]).reshape((x.shape[0], x.shape[1], -1))


def tonemapping(im):
    """util.py:187-200: pow 1/2.4, alpha = 0.8 / 99th percentile of the positive values, clip to [0,1] -- TonemapHDR(2.4, 99, 0.8)
    without returning alpha, so it runs on the same kernel (`csrc/tonemap.cu`).  numpy (H,W,3) in -> numpy out like the reference;
    a CUDA tensor stays a CUDA tensor."""
    import torch
    from .tonemap import TonemapHDR
    is_np = not torch.is_tensor(im)
    x = torch.from_numpy(np.ascontiguousarray(im, dtype=np.float32)).cuda() if is_np else im
    y, _ = TonemapHDR(gamma=2.4, percentile=99, max_mapping=0.8)(x)
    return y.cpu().numpy() if is_np else y


def cartesian_to_polar(xyz):
    return np.arctan2(xyz[1], xyz[0]), np.arccos(np.clip(xyz[2], -1.0, 1.0))              # util.py:206-209 (phi, theta)


def polar_to_cartesian(phi_theta):
    phi, theta = phi_theta                                                                  # util.py:212-220
    return np.stack((np.sin(theta) * np.cos(phi), np.sin(theta) * np.sin(phi), np.cos(theta)), axis=1)


def print_model_parm_nums(model):
    total = sum(p.nelement() for p in model.parameters())                                   # util.py:353-356
    print('  + Number of params: %.2fM' % (total / 1e6))


# ------------------------------------------------------------------------------------------------- GenProjector/util.py output side
def tonemapping_to_file(im, sv_path, gamma=2.4, percentile=50, max_mapping=0.5):
    """GenProjector/util.py:226-245 `tonemapping(im, sv_path, ...)`: tone-map on the tonemap kernel and save as an 8-bit image."""
    import torch
    from PIL import Image
    from .tonemap import TonemapHDR
    x = im if torch.is_tensor(im) else torch.from_numpy(np.ascontiguousarray(im, dtype=np.float32))
    y, _ = TonemapHDR(gamma=gamma, percentile=percentile, max_mapping=max_mapping)(x.cuda())
    Image.fromarray((y.cpu().numpy() * 255.0).astype("uint8")).save(sv_path)


def convert_visuals_to_numpy(visuals):
    for key, t in visuals.items():                                                          # util.py:434-439: (C,H,W) tensor -> (H,W,C) array
        visuals[key] = np.squeeze(t.permute(1, 2, 0).detach().cpu().numpy())
    return visuals


def print_current_errors(epoch, i, errors, t):
    message = '(epoch: %d, iters: %d, time: %.3f)' % (epoch, i, t)                          # util.py:442-447
    for k, v in errors.items():
        message += '%s: %.3f ' % (k, v.mean().float())
    print(message)


def save_test_images(visuals, nm, out_dir="./results"):
    """util.py:468-500 (what GenProjector/test.py:39 calls per image): `<nm>_fake_image.exr` (the HDR illumination map) plus
    tone-mapped previews `<nm>_fake_image.jpg`, `<nm>_warped.jpg`, `<nm>_input.jpg`."""
    import os
    os.makedirs(out_dir, exist_ok=True)
    visuals = convert_visuals_to_numpy(visuals)
    for label, image in visuals.items():
        base = os.path.join(out_dir, nm + '_' + label)
        if label == 'fake_image':
            tonemapping_to_file(image, base + '.jpg')
            wire.write_exr(base + '.exr', image)
        if label == 'warped':
            tonemapping_to_file(image, base + '.jpg')
        if label == 'input':
            tonemapping_to_file(image * 255.0, base + '.jpg', percentile=99, max_mapping=0.99)


def save_current_images(visuals, epoch, step, out_dir="./summary"):
    """util.py:449-466: training-time previews."""
    import os
    from PIL import Image
    os.makedirs(out_dir, exist_ok=True)
    visuals = convert_visuals_to_numpy(visuals)
    for label, image in visuals.items():
        path = os.path.join(out_dir, 'epoch%.3d_iter%.3d_%s.png' % (epoch, step, label))
        if label == 'input':
            tonemapping_to_file(image, path, gamma=2.4, percentile=99, max_mapping=0.8)
        elif label == 'im':
            Image.fromarray((image * 255.0).astype('uint8')).save(path)
        else:
            tonemapping_to_file(image, path)


def save_network(net, label, epoch, opt):
    import os
    import torch
    path = os.path.join(opt.checkpoints_dir, opt.name, '%s_net_%s.pth' % (epoch, label))   # util.py:173-178 (state_dict on the CPU)
    torch.save({k: v.cpu() for k, v in net.state_dict().items()}, path)


def load_network(net, label, epoch, opt):
    import os
    import torch
    path = os.path.join(opt.checkpoints_dir, opt.name, '%s_net_%s.pth' % (epoch, label))   # util.py:181-191
    net.load_state_dict(torch.load(path))
    return net


# ------------------------------------------------------------------------------------------------- GenProjector/util.py small helpers
def mkdir(path):
    import os
    if not os.path.exists(path):                                                            # util.py:127-129
        os.makedirs(path)


def mkdirs(paths):
    for path in (paths if isinstance(paths, list) and not isinstance(paths, str) else [paths]):   # util.py:119-124
        mkdir(path)


def str2bool(v):
    import argparse
    if v.lower() in ('yes', 'true', 't', 'y', '1'):                                         # util.py:149-155
        return True
    if v.lower() in ('no', 'false', 'f', 'n', '0'):
        return False
    raise argparse.ArgumentTypeError('Boolean value expected.')


def copyconf(default_opt, **kwargs):
    import argparse
    conf = argparse.Namespace(**vars(default_opt))                                          # util.py:40-45
    for key, value in kwargs.items():
        setattr(conf, key, value)
    return conf


def natural_sort(items):
    import re
    items.sort(key=lambda text: [int(c) if c.isdigit() else c for c in re.split(r'(\d+)', text)])   # util.py:132-146
