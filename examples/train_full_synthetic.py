#!/usr/bin/env python
"""BASELINE configs[3] as one program: the full EMLight training iteration on synthetic data, one process per GPU --

  1. regression step   DenseNet forward, Sinkhorn EMD + 4 MSE terms, backward, Adam            (RegressionNetwork/train.py:79-102)
  2. guide assembly    predicted parameters -> Gaussian-map panorama                           (GenProjector/data.py:86-102)
  3. generator step    SPADE generator, discriminator, GAN / feature-matching / VGG / cosine losses, backward, Adam
  4. discriminator step                                                                         (pix2pix_model.py:92-141, trainers)

with one bucketed NCCL all-reduce of each network's gradients in front of its optimiser step when launched under torchrun.
The reference runs 1 and 3-4 as two programs with pickles in between; here the guide of step 3 comes from step 1's predictions.

    python examples/train_full_synthetic.py --steps 2 --batch 4 --ngf 16 --ndf 16
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_full_synthetic.py --batch 32      # configs[3]: 256 global

The GenProjector backward is the opt-in tape of emlight_b200/gp_train.py.  Prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "examples"))
import torch
import torch.nn as nn

import emlight_b200 as E
from emlight_b200 import parallel
import train_regression_synthetic as reg


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--batch", type=int, default=4, help="per GPU")
    ap.add_argument("--anchors", type=int, default=128)
    ap.add_argument("--ngf", type=int, default=64)
    ap.add_argument("--ndf", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = parallel.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init("nccl", dev)
    torch.manual_seed(0)
    ln, B = args.anchors, args.batch
    net = E.DenseNet(n_anchors=ln).to(dev).train()
    opt_R = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.999))
    l2 = nn.MSELoss().to(dev)
    sam_loss = E.SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)
    opt = argparse.Namespace(ngf=args.ngf, ndf=args.ndf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", norm_D="spectralinstance",
                             semantic_nc=3, label_nc=3, output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0,
                             num_D=2, n_layers_D=4, netD_subarch="n_layer", no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0],
                             isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
    model = E.Pix2PixModel(opt)
    model.train()
    model.autograd = True
    opt_G, opt_D = model.create_optimizers(opt)
    gen = torch.Generator().manual_seed(1234 + rank)
    batch = reg.synthetic_batch(B, ln, gen, dev)
    warped = (torch.rand(B, 3, 128, 256, generator=gen) * 5).to(dev)
    mask = (torch.rand(B, 1, 128, 256, generator=gen) < 0.1).float().to(dev)

    def sync_grads(params):
        if world > 1:
            parallel.allreduce_mean_([p.grad for p in params if p.grad is not None])

    def iteration():
        loss, _ = reg.train_step(net, sam_loss, l2, opt_R, batch, ln, world)                       # 1
        with torch.no_grad():                                                                      # 2
            net.eval()
            pred = net(batch[0])
            net.train()
            dist = torch.softmax(pred["distribution"], 1)
            rgb = nn.functional.normalize(pred["rgb_ratio"].abs() + 1e-3, dim=1)
            guide = E.genprojector_guide(dist, pred["intensity"].abs() * 500.0, rgb, pred["ambient"].abs() * (128 * 256))
        data = {"input": guide, "crop": batch[0][:, :, :, 32:224].contiguous(), "warped": warped, "map": mask}
        opt_G.zero_grad()                                                                          # 3
        g_losses, _ = model(data, "generator")
        sum(g_losses.values()).mean().backward()
        sync_grads(model.netG.parameters())
        opt_G.step()
        opt_D.zero_grad()                                                                          # 4
        d_losses = model(data, "discriminator")
        sum(d_losses.values()).mean().backward()
        sync_grads(model.netD.parameters())
        opt_D.step()
        return float(loss), {k: float(v.sum()) for k, v in list(g_losses.items()) + list(d_losses.items())}

    first = iteration()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    if rank == 0:
        print(json.dumps({"workload": "BASELINE configs[3]: regression step + guide + GenProjector G step + D step", "n_gpus": world,
                          "batch_per_gpu": B, "global_batch": B * world, "ngf": args.ngf, "ms_per_iteration": ms,
                          "maps_per_s": B * world / ms * 1e3, "regression_loss": first[0], "gan_losses": first[1]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
