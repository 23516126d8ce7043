mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_densenet_gpu.py tests/test_dense_layer_gpu.py -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_c27.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_c27.log; grep -E "^E  " gpurun_out/pytest_c27.log | head -8 | cut -c1-300
for v in 1 0 1 0; do echo "planes=$v: $(EML_DENSE_PLANES=$v timeout 300 python tools/fwd_time.py 256 2>&1 | tail -1)"; done
EML_DENSE_PLANES=1 timeout 300 python tools/layer_times.py 256 > gpurun_out/layer_times_planes.log 2>&1; head -20 gpurun_out/layer_times_planes.log; tail -1 gpurun_out/layer_times_planes.log
