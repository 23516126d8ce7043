mkdir -p gpurun_out
python tools/bench_generator.py --batch 4 --precision bf16x3 --profile 2>&1 | tail -14
python tools/bench_generator.py --batch 4 --precision bf16x3 --graph 2>&1 | tail -1
python tools/bench_generator.py --batch 16 --precision bf16x3 --graph 2>&1 | tail -1
python tools/bench_generator.py --batch 16 --precision bf16x3 --profile 2>&1 | grep -A8 "profile:"
