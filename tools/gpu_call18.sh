mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_genprojector_gpu.py tests/test_gp_train_gpu.py tests/test_discriminator_gpu.py tests/test_handlers_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c18.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_c18.log; grep -E "^E  " gpurun_out/pytest_c18.log | head -8 | cut -c1-300
timeout 600 python tools/bench_generator.py --batch 16 --profile > gpurun_out/gen_profile_b16_v3.log 2>&1; echo "gen exit $?"; tail -45 gpurun_out/gen_profile_b16_v3.log
timeout 600 python tools/bench_generator.py --batch 16 --precision bf16 > gpurun_out/gen_b16_bf16_v3.log 2>&1; tail -1 gpurun_out/gen_b16_bf16_v3.log
SAN_PARTS="initcheck" bash tools/sanitize.sh
