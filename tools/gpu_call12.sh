mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense.log 2>&1; echo "dense pytest exit $?"; tail -4 gpurun_out/pytest_dense.log
EML_DENSE_RS_MAX_C=0 timeout 600 python -m pytest tests/test_dense_layer_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense_z.log 2>&1; echo "dense zstencil pytest exit $?"; tail -3 gpurun_out/pytest_dense_z.log
timeout 600 python tools/layer_ab.py > gpurun_out/layer_ab.log 2>&1; echo "layer_ab exit $?"; cat gpurun_out/layer_ab.log
