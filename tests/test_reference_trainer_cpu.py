"""The reference's OWN GenProjector training scaffolding -- `options/train_options.py`, `options/base_options.py`, `model_trainer.py`,
`iter_counter.py`, unchanged, imported from /root/reference -- running on the module-name shims of `emlight_b200/dropin_genprojector`
(`models`, `models.networks[.sync_batchnorm]`, `util`, `data`): option parsing with the by-name network lookup, `Trainer(opt)`,
`run_generator_one_step`, `run_discriminator_one_step`, `get_latest_losses`, `util.print_current_errors`, `update_learning_rate`,
`save('latest')` and a reload through `--continue_train`.  CPU only: the device primitives are the stand-ins of test_gp_train_cpu
(adjoint kernels through their host-emulation build), `.cuda()` is identity.  Skipped where the reference tree is absent (GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/GenProjector"

SCRIPT = r'''
import os, sys, tempfile
root, ref, work = sys.argv[1], sys.argv[2], sys.argv[3]
sys.argv = ["train.py", "--name", "run", "--checkpoints_dir", work, "--gpu_ids", "-1", "--ngf", "2", "--ndf", "2", "--batchSize", "1",
            "--dataroot", work, "--niter", "1", "--niter_decay", "1"] + sys.argv[4:]
sys.path[:0] = [os.path.join(root, "emlight_b200", "dropin_genprojector"), ref, root, os.path.join(root, "tests")]
import torch
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
import test_gp_train_cpu as C
from emlight_b200 import gp_ops, gp_train, genprojector, _lib
C.install_sims(gp_ops, C.build_emu(tempfile.mkdtemp()))
_lib.require_cuda = lambda *a: None
# the D step's no-grad generator forward is the forward-only product path (real kernels): same network over the stand-ins here
genprojector.Pix2PixModel.generate_fake = lambda self, inp, crop: gp_train.generator(gp_train.Tape(), self.netG, inp, crop, True)

from options.train_options import TrainOptions          # the reference's files from here on
from iter_counter import IterationCounter
from model_trainer import Trainer
import models, util
assert models.__file__.startswith(os.path.join(root, "emlight_b200")) and util.__file__.startswith(os.path.join(root, "emlight_b200"))
import options.base_options, model_trainer, iter_counter
assert all(m.__file__.startswith(ref) for m in (options.base_options, model_trainer, iter_counter))

opt = TrainOptions().parse()
assert opt.norm_G == "spectralspadesyncbatch3x3" and opt.netD == "multiscale" and opt.num_D == 2 and opt.n_layers_D == 4
trainer = Trainer(opt)
model = trainer.pix2pix_model_on_one_gpu
assert type(model).__module__ == "models.pix2pix_model" and model.training and model.autograd
if opt.continue_train:                                              # second invocation: the checkpoints of the first one were loaded
    for label, net in (("G", model.netG), ("D", model.netD)):
        saved = torch.load(os.path.join(work, "run", "latest_net_%s.pth" % label))
        assert all(torch.equal(v, saved[k]) for k, v in net.state_dict().items()), label
    print("TRAINER-OK continue_train")
    sys.exit(0)
counter = IterationCounter(opt, 1)
gen = torch.Generator().manual_seed(3)
data_i = {"input": torch.rand(1, 3, 128, 256, generator=gen) * 2, "crop": torch.rand(1, 3, 128, 128, generator=gen),
          "warped": torch.rand(1, 3, 128, 256, generator=gen) * 20, "map": (torch.rand(1, 1, 128, 256, generator=gen) > 0.9).float()}
w0 = [p.detach().clone() for p in model.netG.parameters()]
d0 = [p.detach().clone() for p in model.netD.parameters()]
for epoch in counter.training_epochs():
    counter.record_epoch_start(epoch)
    counter.record_one_iteration()
    trainer.run_generator_one_step(data_i)
    trainer.run_discriminator_one_step(data_i)
    losses = trainer.get_latest_losses()
    util.print_current_errors(epoch, counter.epoch_iter, losses, counter.time_per_iter)
    trainer.update_learning_rate(epoch)
    counter.record_epoch_end()
    break                                                            # one epoch of work is enough; the schedule is exercised below
trainer.update_learning_rate(opt.niter + 1)
assert set(losses) == {"GAN", "GAN_Feat", "VGG", "COS", "D_Fake", "D_real"}
assert all(torch.isfinite(v).all() for v in losses.values())
assert any(not torch.equal(a, b) for a, b in zip(w0, model.netG.parameters()))
assert any(not torch.equal(a, b) for a, b in zip(d0, model.netD.parameters()))
assert trainer.get_latest_generated().shape == (1, 3, 128, 256)
assert trainer.old_lr < opt.lr                                       # niter = 1, niter_decay = 1: past niter the rate decays
trainer.save("latest")
assert os.path.exists(os.path.join(work, "run", "latest_net_G.pth")) and os.path.exists(os.path.join(work, "run", "latest_net_D.pth"))
print("TRAINER-OK", {k: round(float(v.mean()), 4) for k, v in losses.items()})
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_trainer_runs_unchanged_on_the_shims(tmp_path):
    work = str(tmp_path)
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF, work], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "TRAINER-OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    assert "(epoch: 1, iters: 1, time:" in r.stdout and "update learning rate" in r.stdout
    # --continue_train reloads what save('latest') wrote (pix2pix_model.py:84-87 through util.load_network)
    r2 = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF, work, "--continue_train"], capture_output=True, text=True, timeout=900)
    assert r2.returncode == 0 and "TRAINER-OK" in r2.stdout, (r2.stdout[-1500:], r2.stderr[-3000:])


TEST_SCRIPT = r'''
import os, pickle, runpy, sys, tempfile
root, ref, work = sys.argv[1], sys.argv[2], sys.argv[3]
sys.path[:0] = [os.path.join(root, "emlight_b200", "dropin_genprojector"), ref, root, os.path.join(root, "tests")]
import numpy as np
import torch
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.Tensor.cuda = lambda self, *a, **k: self
_orig_to = torch.Tensor.to
torch.Tensor.to = lambda self, *a, **k: self if (a and isinstance(a[0], (str, torch.device)) and "cuda" in str(a[0])) else _orig_to(self, *a, **k)
_orig_device = torch.device
import test_gp_train_cpu as C
from emlight_b200 import gp_ops, gp_train, genprojector, _lib, wire, tonemap, panorama
from oracle import tonemap_oracle, render_oracle
C.install_sims(gp_ops, C.build_emu(tempfile.mkdtemp()))
_lib.require_cuda = lambda *a: None
genprojector.Pix2PixModel.generate_fake = lambda self, inp, crop: gp_train.generator(gp_train.Tape(), self.netG, inp, crop, self.netG.training)


def cpu_tonemap(self, img, clip=True, alpha=None, gamma=True):               # stand-in for eml_tonemap_hdr (the checker's restatement)
    y, a = tonemap_oracle.tonemap_hdr(img.numpy(), self.gamma, self.percentile, self.max_mapping, clip, alpha, gamma)
    return torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32)), float(a)


def cpu_guide(dist, inten, rgb, amb, alpha=1.0, dirs=None, size=0.0025):      # stand-in for the render kernel behind genprojector_guide
    n = dist.shape[1]
    d = torch.from_numpy(panorama.sphere_points(n)).float().view(1, -1)
    col = (dist.view(1, n, 1) * (inten.view(1, 1, 1) * 0.01) * rgb.view(1, 1, 3)).reshape(1, -1)
    env = render_oracle.convert_to_panorama_torch(d, torch.full((1, n), size), col)
    return (env + amb.view(1, 3, 1, 1) / (128 * 256)) * alpha


tonemap.TonemapHDR.__call__ = cpu_tonemap
panorama.genprojector_guide = cpu_guide
# ---- a two-sample dataset in the reference's layout + a checkpoint for --which_epoch latest
rng = np.random.default_rng(0)
for d in ("pkl", "warped", "crop", "ckpt/run"):
    os.makedirs(os.path.join(work, d))
for nm in ("a", "b"):
    dist = rng.random(128).astype(np.float32); dist /= dist.sum()
    with open(os.path.join(work, "pkl", nm + ".pickle"), "wb") as f:
        pickle.dump({"distribution": dist, "intensity": np.float32(300.0), "rgb_ratio": np.array([0.6, 0.6, 0.5], np.float32),
                     "ambient": np.array([500.0, 400.0, 300.0], np.float32)}, f)
    wire.write_exr(os.path.join(work, "warped", nm + ".exr"), np.exp(rng.normal(-2, 1, (128, 256, 3))).astype(np.float32))
    wire.write_exr(os.path.join(work, "crop", nm + ".exr"), np.exp(rng.normal(-1, 1, (96, 128, 3))).astype(np.float32))
import argparse
from models.networks.generator import SPADEGenerator
g0 = SPADEGenerator(argparse.Namespace(ngf=2, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", semantic_nc=3,
                                       num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0))
torch.save(g0.state_dict(), os.path.join(work, "ckpt", "run", "latest_net_G.pth"))
import data as data_shim
import importlib
data_shim = importlib.reload(data_shim)                                        # pick up the patched genprojector_guide
os.chdir(work)
sys.argv = ["test.py", "--name", "run", "--checkpoints_dir", os.path.join(work, "ckpt"), "--gpu_ids", "-1", "--ngf", "2", "--batchSize", "1",
            "--dataroot", work + "/"]
runpy.run_path(os.path.join(ref, "test.py"), run_name="__main__")              # the reference's test.py, unchanged
out = sorted(os.listdir(os.path.join(work, "results")))
print("TEST-OK", out)
fake = wire.load_exr(os.path.join(work, "results", "a_fake_image.exr"))
assert fake.shape == (128, 256, 3) and np.isfinite(fake).all() and fake.min() >= 0 and fake.max() <= 50
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_test_script_runs_unchanged_on_the_shims(tmp_path):
    """GenProjector/test.py executed as-is (runpy): TestOptions, data.create_dataloader over pickle + EXR files, Pix2PixModel loaded from a
    checkpoint, mode='inference', util.save_test_images -> results/<name>_fake_image.exr + previews.  Device work through CPU stand-ins."""
    r = subprocess.run([sys.executable, "-c", TEST_SCRIPT, ROOT, REF, str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "TEST-OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
    for nm in ("a", "b"):
        for suffix in ("_fake_image.exr", "_fake_image.jpg", "_warped.jpg", "_input.jpg"):
            assert nm + suffix in r.stdout, (nm + suffix, r.stdout[-600:])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is only present in the build container")
def test_reference_train_script_runs_unchanged_on_the_shims(tmp_path):
    """GenProjector/train.py executed as-is (runpy) for one epoch over a two-sample dataset: TrainOptions, data.create_dataloader,
    Trainer, IterationCounter, util.print_current_errors / save_current_images, trainer.save('latest') and the per-epoch save."""
    script = TEST_SCRIPT.replace('sys.argv = ["test.py",', 'sys.argv = ["train.py", "--ndf", "2", "--niter", "1", "--niter_decay", "0", "--print_freq", "1", '
                                 '"--display_freq", "2", "--save_latest_freq", "2", "--save_epoch_freq", "1",')
    script = script.replace('runpy.run_path(os.path.join(ref, "test.py"), run_name="__main__")', 'runpy.run_path(os.path.join(ref, "train.py"), run_name="__main__")')
    script = script[:script.index('out = sorted(os.listdir(os.path.join(work, "results")))')] + \
        'print("TRAIN-OK", sorted(os.listdir(os.path.join(work, "ckpt", "run"))), sorted(os.listdir(os.path.join(work, "summary"))))\n'
    r = subprocess.run([sys.executable, "-c", script, ROOT, REF, str(tmp_path)], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "TRAIN-OK" in r.stdout, (r.stdout[-2500:], r.stderr[-3000:])
    assert "(epoch: 1, iters: 1, time:" in r.stdout and "(epoch: 1, iters: 2, time:" in r.stdout and "End of epoch 1 / 1" in r.stdout
    for f in ("latest_net_G.pth", "latest_net_D.pth", "1_net_G.pth", "1_net_D.pth", "iter.txt", "opt.txt"):
        assert f in r.stdout, (f, r.stdout[-800:])
    assert "epoch001_iter002_fake_image.png" in r.stdout and "epoch001_iter002_input.png" in r.stdout
