mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gemm_gpu.py tests/test_genprojector_gpu.py tests/test_gp_train_gpu.py tests/test_discriminator_gpu.py tests/test_densenet_gpu.py tests/test_conv_gpu.py tests/test_dense_layer_gpu.py tests/test_needlets_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c21.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c21.log; grep -E "^E  " gpurun_out/pytest_c21.log | head -8 | cut -c1-300
timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/profile_gan_step_b4_ngf64_v3.log 2>&1; echo "gan exit $?"; tail -34 gpurun_out/profile_gan_step_b4_ngf64_v3.log | cut -c1-200
timeout 600 python tools/bench_generator.py --batch 16 > gpurun_out/gen_b16_v5.log 2>&1; tail -1 gpurun_out/gen_b16_v5.log
