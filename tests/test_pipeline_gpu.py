"""GPU: the one-process pipeline of examples/pipeline_synthetic.py (tonemap -> DenseNet -> guide render -> SPADE generator)."""
import os
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_pipeline_runs_and_is_consistent(cuda):
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    from pipeline_synthetic import run
    import emlight_b200 as E
    r = run(batch=2, ngf=8, steps=1)
    assert tuple(r["guide"].shape) == (2, 3, 128, 256) and tuple(r["output"].shape) == (2, 3, 128, 256)
    assert torch.isfinite(r["output"]).all() and float(r["output"].min()) >= 0.0 and float(r["output"].max()) <= 50.0
    assert tuple(r["alpha"].shape) == (2,) and (r["alpha"] > 0).all()
    # the guide is the plain render of the predicted parameters plus the predicted ambient term
    p = r["pred"]
    dist = torch.softmax(p["distribution"], 1)
    rgb = torch.nn.functional.normalize(p["rgb_ratio"].abs() + 1e-3, dim=1)
    want = E.render_from_params(dist, p["intensity"].abs() * 5.0, rgb, ambient=p["ambient"].abs(), gain=1.0)
    assert float((r["guide"] - want).abs().max()) <= 1e-4 * float(want.abs().max()) + 1e-6
