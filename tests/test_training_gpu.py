"""GPU: the reference training step (RegressionNetwork/train.py:79-102) on the drop-in modules vs the same step run through the
CPU oracle with torch autograd: identical initial weights, batch, loss weights and torch.optim.Adam."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import ROOT
from oracle import densenet_oracle as DO, sinkhorn_oracle as SO

pytestmark = pytest.mark.gpu


class _OracleSinkhorn(torch.autograd.Function):
    """SamplesLoss value + analytic gradient from the numpy oracle (float64), as an autograd node."""

    @staticmethod
    def forward(ctx, x, y):
        loss, grad = SO.sinkhorn_loss(x.detach().numpy(), y.numpy(), dtype=np.float64)
        ctx.save_for_backward(torch.from_numpy(grad).float())
        return torch.from_numpy(loss).float()

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return (g * go.view(-1, 1)).view(g.shape[0], -1, 1), None


def _losses(pred, batch, sam, ln):
    crop, dist, inten, rgb, amb = batch
    l2 = nn.functional.mse_loss
    dp = pred["distribution"].view(-1, ln, 1)
    return (sam(dp, dist.view(-1, ln, 1)).sum() * 1000.0 + l2(dp, dist.view(-1, ln, 1)) * 1000.0 + l2(pred["intensity"], inten) * 0.1 +
            l2(pred["rgb_ratio"], rgb) * 100.0 + l2(pred["ambient"], amb) * 1.0)


def test_training_trajectory_matches_oracle(cuda):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import emlight_b200 as E
    from train_regression_synthetic import synthetic_batch
    B, ln, steps = 2, 96, 3
    sd = DO.init_state_dict(seed=0, n_anchors=ln)
    batch = synthetic_batch(B, ln, torch.Generator().manual_seed(4), torch.device("cpu"))
    # ---- oracle: functional network over leaf parameters, torch autograd, Adam
    params = {k: v.clone().requires_grad_() for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    state = dict(sd); state.update(params)
    opt = torch.optim.Adam(list(params.values()), lr=1e-4, betas=(0.9, 0.999))
    ref = []
    for _ in range(steps):
        loss = _losses(DO.densenet_forward(state, batch[0], training=True), batch, _OracleSinkhorn.apply, ln)
        opt.zero_grad(); loss.backward(); opt.step()
        ref.append(float(loss))
    # ---- drop-in modules on the GPU (fp32-grade precision mode)
    net = E.DenseNet(n_anchors=ln, precision="bf16x3").to(cuda).train()
    net.load_state_dict(sd)
    opt2 = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
    sam = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
    gb = [t.to(cuda) for t in batch]
    got = []
    for _ in range(steps):
        loss = _losses(net(gb[0]), gb, sam, ln)
        opt2.zero_grad(); loss.backward(); opt2.step()
        got.append(float(loss))
    assert abs(got[0] - ref[0]) <= 1e-3 * abs(ref[0])                       # same forward
    for a, b in zip(got[1:], ref[1:]):                                       # after 1 and 2 Adam updates driven by our gradients
        assert abs(a - b) <= 0.05 * abs(b), (got, ref)
    # the updated weights themselves: Adam's first steps move every weight by ~lr; compare a few tensors
    for name in ("features.conv0.weight", "features.denseblock2.denselayer5.conv1.weight", "fc.weight"):
        d = (dict(net.named_parameters())[name].detach().cpu() - params[name].detach()).abs().max().item()
        assert d <= 3 * 1e-4 * steps, (name, d)


def test_flat_adam_with_sink_equals_torch_adam(cuda):
    """parallel.FlatAdam on the GPU -- parameters / gradients as views of flat buffers, gradients delivered by DenseNet._backward through
    the sink (the fc bucket before the convolutional backward starts), eml_adam_step -- against torch.optim.Adam on a twin network:
    same kernels produce the gradients, so the two trajectories must agree to rounding of the update formula."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    import emlight_b200 as E
    from emlight_b200 import parallel
    from train_regression_synthetic import synthetic_batch
    B, ln, steps = 2, 96, 3
    sd = DO.init_state_dict(seed=0, n_anchors=ln)
    gb = synthetic_batch(B, ln, torch.Generator().manual_seed(4), cuda)
    sam = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
    nets = [E.DenseNet(n_anchors=ln, precision="bf16x3").to(cuda).train() for _ in range(2)]
    for n in nets:
        n.load_state_dict(sd)
    ref = torch.optim.Adam(nets[0].parameters(), lr=1e-4, betas=(0.9, 0.999))
    flat = parallel.FlatAdam(nets[1].named_parameters(), lr=1e-4, betas=(0.9, 0.999))
    recorded = {}

    def recording_sink(named):                       # what DenseNet._backward hands over, block by block
        recorded.update({k: v.detach().clone() for k, v in named.items()})
        flat.sink(named)
    nets[1]._grad_sink = recording_sink
    assert len(flat.buckets) == 2 and "fc.weight" in flat.buckets[0][2]          # heads + fc (34 MB) | everything else
    pa, pb = dict(nets[0].named_parameters()), dict(nets[1].named_parameters())
    losses = [[], []]
    for step in range(steps):
        recorded.clear()
        for i, (net, opt) in enumerate(zip(nets, (ref, flat))):
            loss = _losses(net(gb[0]), gb, sam, ln)
            opt.zero_grad(); loss.backward()
            losses[i].append(float(loss))
        assert flat.early_buckets == 2                                         # both buckets were complete before backward() returned
        # (1) delivery: every gradient the backward produced sits, bit for bit, in its view of the flat buffer (0 + g), nothing came
        # through autograd a second time, and parameters / gradients are views of the flat buffers
        assert set(recorded) == set(pb)
        for name in pb:
            assert pb[name].grad.data_ptr() >= flat.flat_g.data_ptr() and pb[name].data_ptr() >= flat.flat_p.data_ptr(), name
            assert torch.equal(pb[name].grad, recorded[name].reshape(pb[name].shape)), name
        # the twin network (autograd delivery) computed the same gradients up to the run-to-run noise of float atomics (statistics,
        # weight gradients) amplified by ReLU-mask flips: compare the head / fc gradients, which pass no encoder mask, tightly
        for name in ("fc.weight", "fc_dist.weight", "fc_ambient.bias"):
            d = float((pa[name].grad - pb[name].grad).abs().max())
            assert d <= 1e-4 * float(pa[name].grad.abs().max()), (name, d)
        # (2) the fused update equals torch.optim.Adam on IDENTICAL gradients (Adam divides by sqrt(v): a 1e-6 difference in a
        # near-zero gradient would otherwise move a weight by a good fraction of lr and hide formula errors behind a loose bound)
        with torch.no_grad():
            for name in pa:
                pb[name].grad.copy_(pa[name].grad)
                if step == 0:
                    pb[name].copy_(pa[name])
        ref.step(); flat.step()
        for name in pa:
            d = float((pa[name].detach() - pb[name].detach()).abs().max())
            assert d <= 2e-7 + 2e-6 * 1e-4 * (step + 1), (name, step, d)        # fp32 rounding of the update; |update| ~ lr = 1e-4
    assert all(abs(a - b) <= 1e-3 * abs(a) for a, b in zip(*losses)), losses      # the packed-weight caches saw every update
    assert torch.equal(nets[0].state_dict()["features.norm0.running_mean"], nets[1].state_dict()["features.norm0.running_mean"])
