# round 2 call 1: GAN backward per-parameter error table (both precisions), the three re-calibrated tests, the train-step profile
mkdir -p gpurun_out
timeout 600 python tests/debug_gp_bwd.py 4 2 > gpurun_out/gp_bwd_debug.log 2>&1; echo "debug exit $?"; tail -3 gpurun_out/gp_bwd_debug.log
timeout 900 python -m pytest tests/test_gp_train_gpu.py tests/test_handlers_gpu.py tests/test_needlets_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_call1.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_call1.log
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64.log 2>&1; echo "profile exit $?"; cat gpurun_out/profile_train_b64.log
