mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_dense_bwd1_gpu.py tests/test_gp_train_gpu.py tests/test_genprojector_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_c29.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c29.log; grep -E "^E  " gpurun_out/pytest_c29.log | head -8 | cut -c1-300
