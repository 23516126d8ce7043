mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_densenet_gpu.py tests/test_training_gpu.py tests/test_dense_bwd1_gpu.py tests/test_conv_gpu.py tests/test_dense_layer_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c25.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c25.log; grep -E "^E  " gpurun_out/pytest_c25.log | head -8 | cut -c1-300
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64_v5.log 2>&1; echo "profile exit $?"; head -14 gpurun_out/profile_train_b64_v5.log; grep bn_bwd gpurun_out/profile_train_b64_v5.log
