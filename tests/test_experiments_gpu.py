"""GPU: A/B switches -- each must reproduce the default path's results before it can become the default.
  EML_STEM_V1=1    the round-1 stem; the default is variant 2 ([tap][o] shared-memory weights, LDS.128 broadcast, 24 accumulators): bit-identical
  EML_FC_SPLITK=1  fc GEMM split over K (M = B rows fill only 8 CTAs otherwise): equal to fp32 summation order"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, torch
sys.path.insert(0, %r)
import emlight_b200 as E
torch.manual_seed(0)
net = E.DenseNet(n_anchors=128).cuda().eval()
x = torch.rand(64, 3, 192, 256, generator=torch.Generator().manual_seed(1)).cuda()
with torch.no_grad():
    out = net(x)
torch.save({k: v.cpu() for k, v in out.items()}, sys.argv[1])
net.train()
out = net(x[:4])
torch.save({k: v.detach().cpu() for k, v in out.items()}, sys.argv[1] + ".train")
""" % ROOT


def _run(tmp_path, name, **env):
    path = str(tmp_path / name)
    e = dict(os.environ)
    e.update(env)
    subprocess.check_call([sys.executable, "-c", CODE, path], env=e)
    import torch
    return torch.load(path), torch.load(path + ".train")


def test_stem_v2_and_fc_splitk_reproduce_the_default_path(cuda, tmp_path):
    import torch
    base, base_t = _run(tmp_path, "base.pt")
    stem, stem_t = _run(tmp_path, "stem.pt", EML_STEM_V1="1")
    for k in base:
        assert torch.equal(base[k], stem[k]), k                         # same fmaf order -> bit-identical (eval and batch-stat BN)
        assert torch.allclose(base_t[k], stem_t[k], rtol=1e-5, atol=1e-6), k   # statistics go through atomics: summation order
    fc, _ = _run(tmp_path, "fc.pt", EML_FC_SPLITK="1")
    for k in base:
        assert float((base[k] - fc[k]).abs().max()) <= 1e-5 * float(base[k].abs().max()) + 1e-7, k
