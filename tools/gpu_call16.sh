mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_genprojector_gpu.py tests/test_gp_train_gpu.py tests/test_discriminator_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c16.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_c16.log; grep -E "^E  " gpurun_out/pytest_c16.log | head -6 | cut -c1-300
timeout 600 python tools/bench_generator.py --batch 16 --profile > gpurun_out/gen_profile_b16_v2.log 2>&1; echo "gen exit $?"; tail -60 gpurun_out/gen_profile_b16_v2.log
SAN_PARTS="memcheck initcheck" bash tools/sanitize.sh
