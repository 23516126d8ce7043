mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_training_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_train.log
