"""Where one GenProjector training iteration (G step + D step, gp_train.py) spends its device time: CUDA events around every C-ABI
call (the pattern of tools/bench_generator.py --profile), aggregated per entry point; the remainder of the iteration is the torch
bookkeeping (slicing, permutes, optimiser).   python tools/profile_gan_step.py --batch 4 --ngf 64 --ndf 64"""
import argparse, collections, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch
import emlight_b200 as E
from emlight_b200 import _lib
from train_genprojector_synthetic import synthetic_batch

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--ngf", type=int, default=64)
ap.add_argument("--ndf", type=int, default=64)
a = ap.parse_args()
dev = torch.device("cuda:0")
opt = argparse.Namespace(ngf=a.ngf, ndf=a.ndf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", norm_D="spectralinstance",
                         semantic_nc=3, label_nc=3, output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0,
                         num_D=2, n_layers_D=4, netD_subarch="n_layer", no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0],
                         isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
torch.manual_seed(0)
model = E.Pix2PixModel(opt)
model.train()
model.autograd = True
og, od = model.create_optimizers(opt)
data = synthetic_batch(a.batch, torch.Generator().manual_seed(1), dev)


def iteration():
    og.zero_grad(); gl, _ = model(data, "generator"); sum(gl.values()).mean().backward(); og.step()
    od.zero_grad(); dl = model(data, "discriminator"); sum(dl.values()).mean().backward(); od.step()


iteration()
torch.cuda.synchronize()
lib = _lib.load()
rec = []


class Wrap:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = self.fn(*args); e1.record(); rec.append((self.name, e0, e1)); return r


names = [n for n in _lib.SIGNATURES if n not in ("eml_version", "eml_error_string", "eml_device_ok", "eml_conv_wpack_bytes", "eml_sinkhorn_workspace_bytes")]
orig = {n: getattr(lib, n) for n in names}
for n in names: setattr(lib, n, Wrap(n, orig[n]))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s0.record(); iteration(); s1.record(); host = time.perf_counter() - t0
torch.cuda.synchronize()
for n in names: setattr(lib, n, orig[n])
agg, cnt = collections.Counter(), collections.Counter()
for n, e0, e1 in rec: agg[n] += e0.elapsed_time(e1); cnt[n] += 1
total = s0.elapsed_time(s1)
print("one iteration: %.1f ms on the device, %.1f ms of host time to enqueue, %d C-ABI calls, %.1f ms inside them" % (total, host * 1e3, len(rec), sum(agg.values())))
for n, v in agg.most_common(): print("  %-28s x%5d %9.3f ms" % (n, cnt[n], v))
print(json.dumps({"ms_per_iteration": total, "abi_ms": sum(agg.values()), "batch": a.batch, "ngf": a.ngf, "peak_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
# clean timing (no per-call events): 3 iterations back to back
iteration()
torch.cuda.synchronize()
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); c0.record()
for _ in range(3):
    iteration()
c1.record(); host = (time.perf_counter() - t0) / 3
torch.cuda.synchronize()
print(json.dumps({"clean_ms_per_iteration": c0.elapsed_time(c1) / 3, "host_enqueue_ms": host * 1e3}))
if os.environ.get("EML_CPROFILE") == "1":
    # where the HOST time of one iteration goes (the step is enqueue-bound): top functions by own time and by cumulative time
    import cProfile, pstats, io
    pr = cProfile.Profile()
    pr.enable()
    iteration()
    pr.disable()
    torch.cuda.synchronize()
    for key in ("tottime", "cumtime"):
        buf = io.StringIO()
        pstats.Stats(pr, stream=buf).sort_stats(key).print_stats(28)
        print("\n".join(l[:170] for l in buf.getvalue().splitlines()[4:]))
