mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_training_gpu.py tests/test_dense_bwd1_gpu.py tests/test_densenet_gpu.py tests/test_conv_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "train pytest exit $?"; tail -5 gpurun_out/pytest_train.log; grep -E "^E  " gpurun_out/pytest_train.log | head -6 | cut -c1-300
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64.log 2>&1; echo "profile exit $?"; head -16 gpurun_out/profile_train_b64.log
