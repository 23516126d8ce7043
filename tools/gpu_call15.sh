mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_bwd1_gpu.py tests/test_densenet_gpu.py tests/test_training_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c15.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_c15.log; grep -E "^E  " gpurun_out/pytest_c15.log | head -6 | cut -c1-300
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64_v2.log 2>&1; echo "profile exit $?"; head -12 gpurun_out/profile_train_b64_v2.log; grep wgrad_3x3 gpurun_out/profile_train_b64_v2.log
EML_WGRAD3X3_V1=1 timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64_v1.log 2>&1; echo "profile v1 exit $?"; head -2 gpurun_out/profile_train_b64_v1.log; grep wgrad_3x3 gpurun_out/profile_train_b64_v1.log
for s in 64 256 64 256; do echo "fc slice $s: $(EML_FC_SLICE=$s timeout 300 python tools/fwd_time.py 256 2>&1 | tail -1)"; done
timeout 600 python tools/bench_generator.py --batch 16 --profile > gpurun_out/gen_profile_b16.log 2>&1; echo "gen exit $?"; tail -25 gpurun_out/gen_profile_b16.log
timeout 600 python tools/bench_generator.py --batch 16 --precision bf16 > gpurun_out/gen_b16_bf16.log 2>&1; tail -1 gpurun_out/gen_b16_bf16.log
