# First B200 call of the next round: run the GPU tests that were written after round 1's GPU budget was spent (gated by
# EML_PENDING_GPU), the new examples, the A/B of the prepared switches and a per-entry-point profile of the GAN step (≈25 min of box time).
# Usage:  gpurun --timeout 2400 -- 'bash tools/gpu_pending.sh'     (the regular suite + evidence run stay tools/gpu_final.sh)
mkdir -p gpurun_out
EML_PENDING_GPU=1 timeout 1200 python -m pytest tests/test_gp_train_gpu.py tests/test_handlers_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_pending.log 2>&1
echo "pending exit $?"; tail -30 gpurun_out/pytest_pending.log
timeout 300 python examples/predict_exr.py --out gpurun_out/predict > gpurun_out/predict.log 2>&1; echo "predict_exr exit $?"; tail -2 gpurun_out/predict.log
timeout 900 python examples/train_genprojector_synthetic.py --steps 2 --ngf 16 --ndf 16 > gpurun_out/train_gan.log 2>&1; echo "train_genprojector exit $?"; tail -3 gpurun_out/train_gan.log
timeout 900 python examples/train_full_synthetic.py --steps 1 --batch 2 --ngf 16 --ndf 16 > gpurun_out/train_full.log 2>&1; echo "train_full exit $?"; tail -3 gpurun_out/train_full.log
EML_PENDING_GPU=1 timeout 900 python -m pytest tests/test_experiments_gpu.py -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_experiments.log 2>&1; echo "experiments exit $?"; tail -5 gpurun_out/pytest_experiments.log
# A/B of the prepared switches on the headline workload (compare ms_per_step; none of these lines is a bench value of record)
for flags in "" "EML_STEM_V2=1" "EML_FC_SPLITK=1" "EML_STEM_V2=1 EML_FC_SPLITK=1"; do
  env $flags python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print('$flags', d['ms_per_step'], d['value'], d['clocks'])"
done
timeout 300 python examples/make_gt_pickles.py --out-dir gpurun_out/pkl > gpurun_out/make_gt.log 2>&1; echo "make_gt_pickles exit $?"; tail -1 gpurun_out/make_gt.log
timeout 900 python tools/profile_gan_step.py --batch 2 --ngf 32 --ndf 32 > gpurun_out/profile_gan.log 2>&1; echo "profile_gan exit $?"; tail -25 gpurun_out/profile_gan.log
EML_BENCH_GAN=1 python bench.py --steps 5 --warmup 3 2> /dev/null | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(json.dumps(d[\"other_workloads\"], indent=1))"
