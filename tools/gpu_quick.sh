mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_layer_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_fused.log 2>&1; echo "pytest fused exit $?"; tail -5 gpurun_out/pytest_fused.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step']); print(json.dumps(d['roofline']['families']))
PY
KF='regex:dense_layer|conv|stem_kernel|head_pool|linear_kernel|sg_render'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k "$KF" -s 216 -c 72 --csv --log-file gpurun_out/launches_fused.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
