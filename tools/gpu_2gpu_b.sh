mkdir -p gpurun_out
for v in default 4 default 4; do
  if [ "$v" = "default" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2_$v.json 2> gpurun_out/bench_n2_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_$v.json').read().strip().split('\n')[-1])
t=d['other_workloads']['config1_train_step_fwd_bwd_allreduce_adam_b64_per_gpu']; g=d['other_workloads']['config3_genprojector_G_step_plus_D_step_b4_per_gpu']
print('NCCL_MAX_CTAS=$v value',round(d['value']),'train',t['ms_per_step'],'no-comm',t.get('ms_per_step_without_allreduce'),'alone',t['allreduce_alone_ms'],'wait',t.get('allreduce_exposed_wait_ms'),'| gan',g.get('ms_per_iteration'),'alone',g.get('allreduce_alone_ms'))
PY
done
