mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_genprojector_gpu.py -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gen.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_gen.log
python tools/bench_generator.py --batch 4 --precision bf16x3 2>&1 | tail -1
python tools/bench_generator.py --batch 4 --precision bf16 2>&1 | tail -1
python tools/bench_generator.py --batch 16 --precision bf16x3 2>&1 | tail -1
