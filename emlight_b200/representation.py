"""Ground-truth light parameters from HDR panoramas on the GPU -- drop-in for the reference's
``RegressionNetwork/representation/distribution_representation.py:65-120`` ``extract_mesh`` (the inverse of
``convert_to_panorama``; it defines the training targets ``distribution / intensity / rgb_ratio / ambient`` that ``data.py`` loads
from pickles, so panoramas can drive training without the offline pickle step).

``extract_mesh(h, w, ln)`` builds the same tables as the reference constructor (host numpy, once): row weights
``sin((r + .5) / h * pi)`` (:69-74), the endpoint-inclusive direction grid (:76-83), the Fibonacci anchors and the nearest-anchor
LUT ``argsort(|xyz - anchors|)[..., 0]`` (:84-87).  ``compute(hdr)`` runs ``eml_extract_params`` (one CTA per panorama, float64
accumulation) on a CUDA tensor ``(h, w, 3)`` or ``(B, h, w, 3)`` and returns ``(parametric_lights, map)`` like the reference."""
import numpy as np
import torch

from . import _lib
from .panorama import sphere_points


class extract_mesh:                                        # noqa: N801  (the reference's class name)
    def __init__(self, h=128, w=256, ln=64, device=None):
        self.h, self.w, self.ln = h, w, ln
        self.device = torch.device(device if device is not None else "cuda")
        ster = np.sin((np.linspace(0, h, num=h, endpoint=False) + 0.5) / h * np.pi)            # (:69-70), one value per row
        y_ = np.linspace(0, np.pi, num=h)
        x_ = np.linspace(0, 2 * np.pi, num=w)
        X, Y = np.meshgrid(x_, y_)
        xyz = np.stack((np.sin(Y) * np.cos(X), np.sin(Y) * np.sin(X), np.cos(Y)), -1)          # representation/util.py:184-188
        self.anchors = sphere_points(ln)
        dis = np.linalg.norm(xyz[:, :, None, :] - self.anchors[None, None], axis=-1)
        self.idx = np.argsort(dis, axis=-1)[:, :, 0]                                            # (:86)
        self._ster = torch.from_numpy(ster).to(self.device)
        self._idx = torch.from_numpy(self.idx.astype(np.int32).reshape(-1)).to(self.device)

    @_lib.on_tensor_device
    @torch.no_grad()
    def compute(self, hdr):
        lib = _lib.load()
        _lib.require_cuda(hdr)
        single = hdr.dim() == 3
        x = (hdr[None] if single else hdr).float().contiguous()
        if x.dim() != 4 or tuple(x.shape[1:]) != (self.h, self.w, 3):
            raise ValueError("expected hdr (%d, %d, 3) or (B, %d, %d, 3), got %s" % (self.h, self.w, self.h, self.w, tuple(hdr.shape)))
        B = x.shape[0]
        dev = x.device
        dist = torch.empty(B, self.ln, device=dev)
        inten = torch.empty(B, device=dev)
        rgb = torch.empty(B, 3, device=dev)
        amb = torch.empty(B, 3, device=dev)
        mp = torch.empty(B, self.h, self.w, dtype=torch.uint8, device=dev)
        _lib.check(lib.eml_extract_params(_lib.ptr(x), _lib.ptr(self._idx), _lib.ptr(self._ster), B, self.h, self.w, self.ln, _lib.ptr(dist),
                                          _lib.ptr(inten), _lib.ptr(rgb), _lib.ptr(amb), _lib.ptr(mp), _lib.stream_ptr()), "eml_extract_params")
        mp = mp.bool().unsqueeze(-1)
        if single:
            return {"distribution": dist[0], "intensity": inten[0], "rgb_ratio": rgb[0], "ambient": amb[0]}, mp[0]
        return {"distribution": dist, "intensity": inten, "rgb_ratio": rgb, "ambient": amb}, mp
