mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().split('\n')[-1])
print('N=2 value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'n_gpus',d['n_gpus'], d['clocks'])
print(json.dumps(d['other_workloads'], indent=1))
PY
tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-sample 2 2>&1 | tail -1 | cut -c1-200
NCCL_DEBUG=INFO python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 examples/train_regression_synthetic.py --steps 2 --batch 8 2>&1 | grep -E "NVLS|P2P|via|loss" | head -8
