mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_densenet_gpu.py tests/test_training_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "backward or training" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_bwd.log | cut -c1-220
python tools/profile_train.py 64 2>&1 | tail -22
