mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_conv_gpu.py tests/test_densenet_gpu.py tests/test_training_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_bwd.log
python tools/profile_train.py 64 2>&1 | tail -22
