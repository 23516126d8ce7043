mkdir -p gpurun_out
timeout 200 python tools/bench_generator.py --batch 16 --eager > gpurun_out/gen_b16_eager.log 2>&1; tail -2 gpurun_out/gen_b16_eager.log | cut -c1-400
