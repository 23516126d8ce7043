mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_genprojector_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gen.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_gen.log
