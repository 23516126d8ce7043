mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'cpu',d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
print(json.dumps(d['roofline']['families'])); print(d['other_workloads'], d['clocks'], d['roofline']['frac'])
PY
