mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py tests/test_dense_bwd1_gpu.py tests/test_training_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dense.log 2>&1; echo "dense pytest exit $?"; tail -6 gpurun_out/pytest_dense.log; grep -E "^E  " gpurun_out/pytest_dense.log | head -8 | cut -c1-300
EML_DENSE_CW=16 timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense_cw16.log 2>&1; echo "dense cw16 pytest exit $?"; tail -3 gpurun_out/pytest_dense_cw16.log
for flags in "EML_DENSE_SMEM_A=1" "EML_DENSE_CW=16" ""; do
  env $flags python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('[$flags]', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']); print({k: (v['ms_per_step'], v['GBps']) for k, v in d['roofline']['families'].items()})"
done
python tools/layer_times.py 256 > gpurun_out/layer_times_units.log 2>&1; echo "layer times exit $?"; grep -E "dense_layer|sum" gpurun_out/layer_times_units.log | awk '{print $2, $5}' | tr '\n' ' '
echo
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64.log 2>&1; echo "profile exit $?"; head -24 gpurun_out/profile_train_b64.log
