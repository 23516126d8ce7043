# session-2 baseline: full gpu tests, bench, launch list, ncu --set full of the two dominant kernels (run under gpurun)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider --durations=15 > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
tail -c 2500 gpurun_out/bench_n1.json
KF='regex:conv|stem_kernel|head_pool|linear_kernel|sg_render'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -s 312 -c 104 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv1x1_persist -s 298 -c 1 -o gpurun_out/prof_c1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1; echo "ncu c1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_roll -s 200 -c 1 -o gpurun_out/prof_c2 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 exit $?"
ls -la gpurun_out
