"""GPU: whole regression network through the drop-in module vs the reference-generated golden outputs."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import densenet_oracle as DO

pytestmark = pytest.mark.gpu
KEYS = ("distribution", "intensity", "rgb_ratio", "ambient")
# north-star: 1e-3 relative fp32 for the parity modes; single-pass bf16 is reported, not gated at 1e-3
TOL = {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 8e-2}


def _net(cuda, precision, g):
    import emlight_b200 as E
    net = E.DenseNet(precision=precision).to(cuda)
    net.load_state_dict(DO.init_state_dict(seed=int(g["sd_seed"]), n_anchors=96))
    return net


@pytest.mark.parametrize("precision", ["bf16x3", "fp32", "bf16"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_densenet_matches_reference_golden(cuda, precision, mode):
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    net = _net(cuda, precision, g).train(mode == "train")
    x = torch.rand(2, 3, 192, 256, generator=torch.Generator().manual_seed(int(g["x_seed"]))).to(cuda)
    with torch.no_grad():
        out = net(x)
    assert set(out) == set(KEYS)
    for k in KEYS:
        ref = g["%s_%s" % (mode, k)]
        got = out[k].cpu().numpy()
        assert got.shape == ref.shape
        err = np.abs(got - ref).max() / np.abs(ref).max()
        assert err <= TOL[precision], (k, err)
    if precision != "bf16":
        # anchor-index argmax is bit-exact
        assert np.array_equal(out["distribution"].argmax(1).cpu().numpy(), g[mode + "_distribution"].argmax(1))
    if mode == "train" and precision != "bf16":
        # one training-mode forward updates the running statistics exactly like nn.BatchNorm2d
        for name, mod in (("norm0", net.features.norm0), ("last_norm3", net.features.last_norm3)):
            rv = g["train_%s_running_var" % name]
            assert np.abs(mod.running_var.cpu().numpy() - rv).max() <= 1e-3 * np.abs(rv).max()
        rm = g["train_norm0_running_mean"]
        assert np.abs(net.features.norm0.running_mean.cpu().numpy() - rm).max() <= 1e-3 * np.abs(rm).max() + 1e-6
        assert int(net.features.norm0.num_batches_tracked) == 1


def test_densenet_batch_independence_and_determinism(cuda):
    """Eval mode: each sample's output is independent of its batch; repeated calls are bit-identical."""
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    net = _net(cuda, "bf16x3", g).eval()
    x = torch.rand(5, 3, 192, 256, generator=torch.Generator().manual_seed(99)).to(cuda)
    with torch.no_grad():
        a = net(x)["distribution"].clone()
        b = net(x)["distribution"].clone()
        c = net(x[2:3])["distribution"].clone()
    assert torch.equal(a, b)
    assert torch.equal(a[2:3], c)


@pytest.mark.parametrize("precision,bwd_precision,l2_tol,med_tol", [("fp32", "fp32", 1.5e-2, 5e-3), ("fp32", "bf16x3", 1.5e-2, 5e-3),
                                                                     ("bf16x3", "bf16x3", 5e-2, 2.5e-2)],
                         ids=["fp32", "fp32fwd_bf16x3bwd", "bf16x3"])
def test_densenet_backward_matches_oracle_autograd(cuda, precision, bwd_precision, l2_tol, med_tol):
    """Training-mode BN (what train.py runs): every parameter gradient vs torch autograd through the CPU oracle.

    Conditioning: these B=2 gradients are sums over ReLU masks; the reference's OWN fp32 and fp64 gradients differ by
    median 1.3e-3 / p90 2.5e-3 (max-relative, forward agreeing to 9e-7) and a 1e-6 relative input perturbation moves them by
    1.9e-3 / 3.6e-3 (measured with oracle/).  The fp32 mode (forward error ~2e-6) lands on that floor
    (measured: median 1.9e-3, worst per-tensor L2 5.8e-3); bf16x3's forward differs by ~2e-5, flips ~20x more masks and lands at
    ~1.2e-2.  The middle case ISOLATES that claim: forward in fp32 mode (the oracle's masks), backward with the bf16x3 tensor-core
    kernels (dgrad GEMMs, MN-major wgrads) -- it must meet the fp32 bounds, i.e. the looser bf16x3 bound is a property of the forward's
    rounding, not of the backward kernels.  Tensors whose true gradient is analytically ~0 (last_norm{1,2}: a BatchNorm feeding only
    BatchNorms) are compared absolutely."""
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    sd = DO.init_state_dict(seed=int(g["sd_seed"]), n_anchors=96)
    x = torch.rand(2, 3, 192, 256, generator=torch.Generator().manual_seed(21))
    gen = torch.Generator().manual_seed(22)
    R = {k: torch.randn(2, n, generator=gen) for k, n in (("distribution", 96), ("intensity", 1), ("rgb_ratio", 3), ("ambient", 3))}
    sdo = {k: (v.clone().requires_grad_() if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
    out = DO.densenet_forward(sdo, x, training=True)
    sum((out[k] * R[k]).sum() for k in KEYS).backward()
    net = _net(cuda, precision, g).train()
    o = net(x.to(cuda))
    net.precision = bwd_precision                      # the backward reads the precision at call time; the packed weights exist in every mode
    sum((o[k] * R[k].to(cuda)).sum() for k in KEYS).backward()
    gmax = max(float(v.grad.abs().max()) for v in sdo.values() if getattr(v, "grad", None) is not None)
    l2s = []
    for name, p in net.named_parameters():
        ref = sdo[name].grad
        assert p.grad is not None and p.grad.shape == ref.shape, name
        got = p.grad.cpu()
        if float(ref.abs().max()) <= 1e-4 * gmax:
            assert float((got - ref).abs().max()) <= 1e-5 * gmax, name
            continue
        l2 = float((got - ref).norm() / ref.norm())
        assert l2 <= l2_tol, (name, l2)
        l2s.append(l2)
    assert len(l2s) > 300
    assert float(np.median(l2s)) <= med_tol, float(np.median(l2s))
    # head / fc gradients do not pass through any ReLU mask of the encoder: tight
    for name in ("fc.weight", "fc_dist.weight", "fc_ambient.bias"):
        ref = sdo[name].grad
        assert float((dict(net.named_parameters())[name].grad.cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), name


def test_densenet_b256_rows_match_oracle(cuda):
    """BASELINE configs[2] size: one B = 256 eval forward (fused dense layers, pair-mode block 3, transition on the TMA pipeline, tcgen05 fc)
    against the CPU oracle on three of its rows -- samples are independent in eval mode -- plus bit-exact argmax."""
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    net = _net(cuda, "bf16x3", g).eval()
    x = torch.rand(256, 3, 192, 256, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        out = net(x.to(cuda))
    rows = [0, 101, 255]
    sd = DO.init_state_dict(seed=int(g["sd_seed"]), n_anchors=96)
    with torch.no_grad():
        ref = DO.densenet_forward(sd, x[rows], training=False)
    for k in KEYS:
        got = out[k][rows].cpu()
        assert float((got - ref[k]).abs().max()) <= 1e-3 * float(ref[k].abs().max()), k
    assert torch.equal(out["distribution"][rows].argmax(1).cpu(), ref["distribution"].argmax(1))
    assert bool(torch.isfinite(out["distribution"]).all())


def test_densenet_backward_contract(cuda):
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    net = _net(cuda, "bf16x3", g).eval()
    out = net(torch.rand(1, 3, 192, 256, device=cuda))
    with pytest.raises(NotImplementedError):                 # eval-mode BN backward is not implemented (train.py trains in train mode)
        out["distribution"].sum().backward()
    with pytest.raises(ValueError):
        net(torch.rand(1, 3, 256, 256, device=cuda))       # the network is built for 192x256
    net.train()
    a = net(torch.rand(2, 3, 192, 256, device=cuda))
    net(torch.rand(2, 3, 192, 256, device=cuda))           # a second forward overwrites the workspace the first backward needs
    with pytest.raises(RuntimeError, match="must follow"):
        a["distribution"].sum().backward()


@pytest.mark.parametrize("B", [1, 3, 8])
def test_channel_plane_slab_equals_pixel_records(cuda, B, monkeypatch):
    """Block 1's slab as channel planes (eml_dense_layer_params.plane_pixels: contiguous TMA boxes, row-contiguous writes) is only a
    different ADDRESSING of the same computation: outputs are bit-identical to the NHWC pixel-record layout, also after the workspace
    has been used by a previous batch (stale channels behind C_in are multiplied by zero weights) and next to a training-mode call."""
    g = np.load(os.path.join(GOLDEN, "densenet.npz"))
    net = _net(cuda, "bf16x3", g).eval()
    gen = torch.Generator().manual_seed(B)
    x1 = torch.rand(B, 3, 192, 256, generator=gen).to(cuda)
    x2 = torch.rand(B, 3, 192, 256, generator=gen).to(cuda) * 3.0
    with torch.no_grad():
        monkeypatch.setenv("EML_DENSE_PLANES", "0")
        want1 = [t.clone() for t in net(x1).values()]
        want2 = [t.clone() for t in net(x2).values()]
        monkeypatch.delenv("EML_DENSE_PLANES")
        got1 = [t.clone() for t in net(x1).values()]
        ws = next(iter(net._ws.values()))
        assert "slab_planes0" in ws                                   # the plane path really ran
        got2 = [t.clone() for t in net(x2).values()]                  # second batch on the same (now dirty) planes
        got1b = [t.clone() for t in net(x1).values()]
    for a, b in zip(want1 + want2 + want1, got1 + got2 + got1b):
        assert torch.equal(a, b)
