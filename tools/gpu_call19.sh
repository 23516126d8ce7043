mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_genprojector_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c19.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c19.log; grep -E "^E  " gpurun_out/pytest_c19.log | head -8 | cut -c1-300
timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/profile_gan_step_b4_ngf64.log 2>&1; echo "gan exit $?"; tail -40 gpurun_out/profile_gan_step_b4_ngf64.log
timeout 600 python tools/bench_generator.py --batch 16 > gpurun_out/gen_b16_v4.log 2>&1; tail -1 gpurun_out/gen_b16_v4.log
