"""CPU oracle for the ground-truth parameter extraction (TEST INFRASTRUCTURE ONLY, see oracle/__init__.py).

Restates RegressionNetwork/representation/distribution_representation.py:65-120 (`extract_mesh`) in numpy float64:
  __init__  :66-87   row weights sin((r+.5)/h*pi), endpoint-inclusive (phi, theta) grid -> unit vectors (representation/util.py:184-188),
                     Fibonacci anchors (util.py:190-203), nearest-anchor LUT argsort(|xyz - anchors|)[..., 0]
  compute   :89-119  weighted panorama, 5 % threshold of the brightest weighted intensity, per-anchor sums, ambient, distribution,
                     intensity, rgb_ratio
Pinned against the reference class itself (exec'd from the cited lines) by oracle/make_golden_extract.py -> tests/golden/extract.npz."""
import numpy as np

from .render_oracle import sphere_points


class ExtractMesh:
    def __init__(self, h=128, w=256, ln=64):
        self.h, self.w, self.ln = h, w, ln
        ster = np.sin((np.linspace(0, h, num=h, endpoint=False) + 0.5) / h * np.pi)
        self.steradian = np.tile(ster[:, None], (1, w))[..., None]
        X, Y = np.meshgrid(np.linspace(0, 2 * np.pi, num=w), np.linspace(0, np.pi, num=h))
        xyz = np.stack((np.sin(Y) * np.cos(X), np.sin(Y) * np.sin(X), np.cos(Y)), -1)
        self.anchors = sphere_points(ln)
        dis = np.linalg.norm(xyz[:, :, None, :] - self.anchors[None, None], axis=-1)
        self.idx = np.argsort(dis, axis=-1)[:, :, 0]

    def compute(self, hdr):
        hdr = self.steradian * hdr
        it = 0.3 * hdr[..., 0] + 0.59 * hdr[..., 1] + 0.11 * hdr[..., 2]
        mp = (it > it.max() * 0.05)[..., None]
        light, remain = hdr * mp, hdr * (1 - mp)
        ambient = remain.sum(axis=(0, 1))
        anchors = np.zeros((self.ln, 3))
        np.add.at(anchors, self.idx.reshape(-1), light.reshape(-1, 3))
        energy = 0.3 * anchors[..., 0] + 0.59 * anchors[..., 1] + 0.11 * anchors[..., 2]
        rgb = anchors.sum(0)
        inten = np.linalg.norm(rgb)
        return {"distribution": energy / energy.sum(), "intensity": inten, "rgb_ratio": rgb / inten, "ambient": ambient}, mp
