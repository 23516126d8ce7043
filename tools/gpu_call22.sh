mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_genprojector_gpu.py tests/test_gp_train_gpu.py tests/test_discriminator_gpu.py tests/test_handlers_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c22.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c22.log; grep -E "^E  " gpurun_out/pytest_c22.log | head -8 | cut -c1-300
timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/profile_gan_step_b4_ngf64_v4.log 2>&1; echo "gan exit $?"; tail -3 gpurun_out/profile_gan_step_b4_ngf64_v4.log | cut -c1-200
EML_TORCH_SPECTRAL=1 timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 2>&1 | tail -1
timeout 600 python tools/bench_generator.py --batch 16 > gpurun_out/gen_b16_v6.log 2>&1; tail -1 gpurun_out/gen_b16_v6.log
