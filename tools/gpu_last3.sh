mkdir -p gpurun_out
timeout 150 python tools/eager_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/gan_eager_b4.log 2>&1; tail -2 gpurun_out/gan_eager_b4.log | cut -c1-400
