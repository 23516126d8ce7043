# usage (under gpurun --gpus N): bash tools/gpu_ngpu.sh N
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench$N exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('N=$N value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],2),'n_gpus',d['n_gpus'], d['clocks'])
t=d['other_workloads']['config1_train_step_fwd_bwd_allreduce_adam_b64_per_gpu']; g=d['other_workloads']['config3_genprojector_G_step_plus_D_step_b4_per_gpu']
print('train',t); print('gan',g)
PY
tail -3 gpurun_out/bench_n$N.err
