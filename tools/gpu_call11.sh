mkdir -p gpurun_out
tools/micro/record_rw_bw > gpurun_out/record_rw_bw.log 2>&1; cat gpurun_out/record_rw_bw.log
timeout 600 python tools/layer_ab.py > gpurun_out/layer_ab.log 2>&1; echo "layer_ab exit $?"; cat gpurun_out/layer_ab.log
timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense.log 2>&1; echo "dense pytest exit $?"; tail -4 gpurun_out/pytest_dense.log
