"""CPU: every oracle restatement against the vectors the REFERENCE modules produced (oracle/make_golden.py)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import densenet_oracle as DO, render_oracle as RO, sinkhorn_oracle as SO
from oracle.make_golden import render_go


def test_densenet_oracle_matches_reference_outputs():
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    sd = DO.init_state_dict(seed=int(g["sd_seed"]), n_anchors=96)
    x = torch.rand(2, 3, 192, 256, generator=torch.Generator().manual_seed(int(g["x_seed"])))
    torch.set_num_threads(os.cpu_count())
    for mode in ("eval", "train"):
        with torch.no_grad():
            out = DO.densenet_forward(sd, x, training=(mode == "train"))
        for k, v in out.items():
            ref = g["%s_%s" % (mode, k)]
            # identical op sequence -> agreement to fp32 round-off (thread-count dependent summation order only)
            assert np.abs(v.numpy() - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max()), (mode, k)


def test_sinkhorn_oracle_matches_geomloss_and_gmloss():
    g = np.load(os.path.join(GOLDEN, "sinkhorn.npz"))
    for name in ("geomloss", "gmloss"):
        x, y = g[name + "_x"], g[name + "_y"]
        N = x.shape[1]
        M = SO.anchor_distances(N) if name == "geomloss" else SO.geometric_anchor_distances(g["gmloss_geometry"])
        assert np.abs(M - g[name + "_M"]).max() < 2e-6
        for dt in (np.float64, np.float32):
            loss, grad = SO.sinkhorn_loss(x, y, M=M, dtype=dt)
            assert np.abs(loss - g[name + "_loss"]).max() / np.abs(g[name + "_loss"]).max() < 5e-5
            assert np.abs(grad - g[name + "_grad"][:, :, 0]).max() / np.abs(g[name + "_grad"]).max() < 5e-5


def test_epsilon_schedule_shape():
    e = SO.epsilon_schedule(1.0, 0.025, 0.5)
    assert e[0] == 1.0 and abs(e[-1] - 0.025 ** 2) < 1e-12 and len(e) == 8     # d^2, 6 halvings of the blur, blur^2
    assert len(SO.epsilon_schedule(0.01, 0.025, 0.5)) == 2                       # diameter below blur: no descent steps


def test_render_oracle_matches_reference():
    g = np.load(os.path.join(GOLDEN, "render.npz"))
    for N in (96, 128):
        assert np.abs(RO.sphere_points(N) - g["points_%d" % N]).max() == 0.0
        d, s, c = g["dirs_%d" % N], g["sizes_%d" % N], g["colors_%d" % N]
        pano = RO.convert_to_panorama(d, s, c)
        ref = g["pano_%d" % N]
        assert np.abs(pano[:, :, ::2, ::2] - ref).max() <= 2e-4 * np.abs(ref).max()
        gd, gs, gc = RO.convert_to_panorama_grad(d, s, c, render_go(d.shape[0]))
        for mine, key in ((gd, "gdirs"), (gs, "gsizes"), (gc, "gcolors")):
            r = g["%s_%d" % (key, N)]
            assert np.abs(mine - r).max() <= 2e-3 * np.abs(r).max(), key


def test_compose_colors_order():
    d = np.arange(6, dtype=np.float32).reshape(1, 6)
    c = RO.compose_colors(d, np.array([[2.0]], np.float32), np.array([[1, 10, 100]], np.float32), gain=1.0)
    assert c.shape == (1, 18) and list(c[0, 3:6]) == [2.0, 20.0, 200.0]        # k-major, channel-minor
