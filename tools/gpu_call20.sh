mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_genprojector_gpu.py tests/test_gp_train_gpu.py tests/test_discriminator_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c20.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c20.log; grep -E "^E  " gpurun_out/pytest_c20.log | head -8 | cut -c1-300
timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/profile_gan_step_b4_ngf64_v2.log 2>&1; echo "gan exit $?"; tail -34 gpurun_out/profile_gan_step_b4_ngf64_v2.log | cut -c1-200
EML_IM2COL_GENERAL=1 timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 2>&1 | tail -1
