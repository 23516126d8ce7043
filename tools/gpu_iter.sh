# quick iteration: gpu tests + bench + ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'cpu',d['cpu_baseline'])
print(json.dumps(d['roofline']['families']))
PY
KF='regex:conv|stem_kernel|head_pool|linear_kernel|sg_render'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -s 312 -c 104 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
