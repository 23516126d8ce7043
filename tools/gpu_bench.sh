# bench + ncu evidence on one B200 (run under gpurun).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
KF='regex:conv|stem_kernel|head_pool|linear_kernel|sg_render'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -s 312 -c 104 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 317 -c 2 -o gpurun_out/prof_conv -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
