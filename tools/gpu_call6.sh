mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_bwd1_gpu.py tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py tests/test_training_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_dense.log 2>&1; echo "dense pytest exit $?"; tail -6 gpurun_out/pytest_dense.log; grep -E "^E  " gpurun_out/pytest_dense.log | head -8 | cut -c1-300
EML_DENSE_CW=8 timeout 600 python -m pytest tests/test_dense_layer_gpu.py tests/test_densenet_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_dense_cw8.log 2>&1; echo "dense cw8 pytest exit $?"; tail -3 gpurun_out/pytest_dense_cw8.log
for flags in "EML_DENSE_SMEM_A=1" "" "EML_DENSE_CW=8"; do
  env $flags python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('[$flags]', d['ms_per_step'], d['value'], d['roofline']['frac'], d['clocks']); print({k: (v['ms_per_step'], v['GBps']) for k, v in d['roofline']['families'].items()})"
done
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_all.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').readline())
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
print(json.dumps(d['roofline']['families'], indent=1))
print(json.dumps(d['other_workloads'], indent=1))
print(d['cpu_baseline'])
PY
