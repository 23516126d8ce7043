#!/usr/bin/env python
"""The reference's GenProjector training iteration (GenProjector/train.py + trainers: `run_generator_one_step`, then
`run_discriminator_one_step`; pix2pix_model.py:40-141) on synthetic data, running on the sm_100a drop-in modules with the
tape-based backward of `emlight_b200/gp_train.py` (opt-in: `model.autograd = True`).

Single GPU:   python examples/train_genprojector_synthetic.py --steps 3 --ngf 16 --ndf 16
Multi GPU :   torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 examples/train_genprojector_synthetic.py
              (one process per GPU; the batch is sharded; SPADE's batch statistics are all-reduced inside the forward and backward,
              and each optimiser step is preceded by ONE bucketed NCCL all-reduce of that network's gradients)

Adam betas (0, 0.9) and the TTUR learning rates of options/train_options.py:27-38.  Prints one JSON line with the step time.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import emlight_b200 as E
from emlight_b200 import parallel


def synthetic_batch(B, gen, dev):
    """Shapes / distributions of SURVEY.md section 8d (guide ~ rendered Gaussian map, warped = guide x log-normal, sparse mask)."""
    guide = torch.rand(B, 3, 128, 256, generator=gen) * 2
    crop = torch.rand(B, 3, 128, 128, generator=gen)
    warped = (guide * torch.exp(0.5 * torch.randn(B, 3, 128, 256, generator=gen))).clamp_min(0)
    mask = (torch.rand(B, 1, 128, 256, generator=gen) < 0.1).float()
    return {"input": guide.to(dev), "crop": crop.to(dev), "warped": warped.to(dev), "map": mask.to(dev)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2, help="per GPU (train_laval.sh uses 16 over 2 GPUs)")
    ap.add_argument("--ngf", type=int, default=64)
    ap.add_argument("--ndf", type=int, default=64)
    args = ap.parse_args()
    rank, world, local = parallel.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    parallel.init("nccl", dev)
    opt = argparse.Namespace(ngf=args.ngf, ndf=args.ndf, norm_G="spectralspadesyncbatch3x3", norm_E="spectralinstance", norm_D="spectralinstance",
                             semantic_nc=3, label_nc=3, output_nc=3, num_upsampling_layers="normal", crop_size=256, aspect_ratio=2.0,
                             num_D=2, n_layers_D=4, netD_subarch="n_layer", no_ganFeat_loss=False, no_vgg_loss=False, gpu_ids=[0],
                             isTrain=True, gan_mode="hinge", lr=0.0002, beta1=0.0, beta2=0.9, no_TTUR=False)
    torch.manual_seed(0)                                        # identical initial weights on every rank
    model = E.Pix2PixModel(opt)
    model.train()
    model.autograd = True
    opt_G, opt_D = model.create_optimizers(opt)
    gen = torch.Generator().manual_seed(1234 + rank)
    data = synthetic_batch(args.batch, gen, dev)

    def one_iteration():
        opt_G.zero_grad()
        g_losses, _ = model(data, "generator")                 # model_trainer.run_generator_one_step
        g_loss = sum(g_losses.values()).mean()
        g_loss.backward()
        if world > 1:
            parallel.allreduce_mean_([p.grad for p in model.netG.parameters() if p.grad is not None])
        opt_G.step()
        opt_D.zero_grad()
        d_losses = model(data, "discriminator")                # model_trainer.run_discriminator_one_step
        d_loss = sum(d_losses.values()).mean()
        d_loss.backward()
        if world > 1:
            parallel.allreduce_mean_([p.grad for p in model.netD.parameters() if p.grad is not None])
        opt_D.step()
        return {k: float(v.sum()) for k, v in list(g_losses.items()) + list(d_losses.items())}

    losses = one_iteration()                                    # warm-up (LUT construction, allocator)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = parallel.max_over_ranks(e0.elapsed_time(e1), dev) / args.steps
    if rank == 0:
        print(json.dumps({"workload": "GenProjector G step + D step (fwd + bwd + Adam)", "ngf": args.ngf, "ndf": args.ndf, "n_gpus": world,
                          "batch_per_gpu": args.batch, "ms_per_iteration": ms, "maps_per_s": world * args.batch / ms * 1e3,
                          "first_iteration_losses": losses}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
