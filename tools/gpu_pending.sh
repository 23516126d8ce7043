# First B200 call of the next round: run the GPU tests that were written after round 1's GPU budget was spent (gated by
# EML_PENDING_GPU), then the regular suite.  Usage:  gpurun --timeout 1500 -- 'bash tools/gpu_pending.sh'
mkdir -p gpurun_out
EML_PENDING_GPU=1 timeout 1200 python -m pytest tests/test_gp_train_gpu.py tests/test_handlers_gpu.py -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_pending.log 2>&1
echo "pending exit $?"; tail -30 gpurun_out/pytest_pending.log
timeout 300 python examples/predict_exr.py --out gpurun_out/predict > gpurun_out/predict.log 2>&1; echo "predict_exr exit $?"; tail -2 gpurun_out/predict.log
timeout 900 python examples/train_genprojector_synthetic.py --steps 2 --ngf 16 --ndf 16 > gpurun_out/train_gan.log 2>&1; echo "train_genprojector exit $?"; tail -3 gpurun_out/train_gan.log
timeout 900 python examples/train_full_synthetic.py --steps 1 --batch 2 --ngf 16 --ndf 16 > gpurun_out/train_full.log 2>&1; echo "train_full exit $?"; tail -3 gpurun_out/train_full.log
