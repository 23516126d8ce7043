"""PyTorch-eager baseline of one GenProjector iteration (G step + D step, pix2pix_model.py:92-141 / model_trainer.py:34-50) on the same GPU:
the reference's ops (CPU restatement run on cuda: grid_sample + conv2d, torch autograd, torch.optim.Adam), fp32, TF32 off, train-mode
SPADE statistics and one spectral-norm power iteration per wrapped conv.  Usage: python tools/eager_gan_step.py [--ngf 64 --ndf 64 --batch 4]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch
from oracle import genprojector_oracle as GO
from train_genprojector_synthetic import synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--ngf", type=int, default=64)
ap.add_argument("--ndf", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--device", default="cuda:0")
a = ap.parse_args()
dev = torch.device(a.device)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def leaves(sd):
    out = {}
    for k, v in sd.items():
        t = v.to(dev)
        if t.is_floating_point() and not (k.endswith("_u") or k.endswith("_v") or "running_" in k):
            t.requires_grad_(True)
        out[k] = t
    return out


sdg, sdd = leaves(GO.init_generator_state_dict(1, a.ngf)), leaves(GO.init_discriminator_state_dict(2, a.ndf))
sdv = {k: v.to(dev) for k, v in GO.init_vgg_state_dict(3).items()}
pg = [t for t in sdg.values() if t.requires_grad]
pd = [t for t in sdd.values() if t.requires_grad]
og = torch.optim.Adam(pg, lr=1e-4, betas=(0.0, 0.9))
od = torch.optim.Adam(pd, lr=4e-4, betas=(0.0, 0.9))
data = synthetic_batch(a.batch, torch.Generator().manual_seed(1), dev)


def write_back(sd, upd):
    with torch.no_grad():
        for k, v in upd.items():
            if k in sd and not sd[k].requires_grad:
                sd[k].copy_(v)


def iteration():
    og.zero_grad(set_to_none=True); od.zero_grad(set_to_none=True)
    upd = {}
    fake = GO.generator_forward(sdg, data["input"], data["crop"], a.ngf, upd=upd)
    gl = GO.generator_losses(sdd, sdv, data["input"], fake, data["warped"], data["map"])
    sum(gl.values()).backward()
    og.step(); write_back(sdg, upd)
    od.zero_grad(set_to_none=True)
    with torch.no_grad():
        upd = {}
        fake = GO.generator_forward(sdg, data["input"], data["crop"], a.ngf, upd=upd)
    write_back(sdg, upd)
    dl = GO.discriminator_losses(sdd, data["input"], fake.detach(), data["warped"])
    sum(dl.values()).backward()
    od.step()
    return float(sum(gl.values())), float(sum(dl.values()))


for _ in range(2):
    losses = iteration()
if dev.type == "cuda":
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        iteration()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"workload": "GenProjector G step + D step, PyTorch eager (reference ops) on cuda:0", "ngf": a.ngf, "batch": a.batch,
                      "ms_per_iteration": ms, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30, "losses": losses}))
else:
    print(json.dumps({"losses": losses}))
