mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'])
ow=d['other_workloads']
for k in ow:
    if k.startswith('config1_train') or k.startswith('config3'): print(k, {a:b for a,b in ow[k].items() if 'ms' in a})
PY
