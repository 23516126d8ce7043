mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense.log 2>&1; echo "dense pytest exit $?"; tail -6 gpurun_out/pytest_dense.log; grep -E "^E  " gpurun_out/pytest_dense.log | head -6 | cut -c1-300
EML_DENSE_CW=16 timeout 600 python -m pytest tests/test_dense_layer_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense_cw16.log 2>&1; echo "dense cw16 pytest exit $?"; tail -3 gpurun_out/pytest_dense_cw16.log
EML_DENSE_SMEM_A=1 timeout 600 python -m pytest tests/test_dense_layer_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense_smema.log 2>&1; echo "dense smem-A pytest exit $?"; tail -3 gpurun_out/pytest_dense_smema.log
for flags in "EML_DENSE_SMEM_A=1" "EML_DENSE_CW=16" "EML_DENSE_NO_STASH=1" ""; do
  env $flags timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('[$flags]', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']); print({k: (v['ms_per_step'], v['GBps']) for k, v in d['roofline']['families'].items()})"
done
timeout 300 python tools/layer_times.py 256 > gpurun_out/layer_times_rs.log 2>&1; echo "layer times exit $?"; grep -E "dense_layer|sum" gpurun_out/layer_times_rs.log | awk '{print $2, $5}' | tr '\n' ' '
echo
