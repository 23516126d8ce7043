mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_discriminator_gpu.py tests/test_genprojector_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_disc.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_disc.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; tail -2 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step']); print(d['other_workloads'])
PY
