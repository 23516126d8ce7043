mkdir -p gpurun_out
# one mid-block-1 conv1 (3 chunks) and one conv2 launch, full sections + source
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv1x1_persist -s 298 -c 1 -o gpurun_out/prof_c1 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c1.log 2>&1; echo "ncu c1 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_rows_persist -s 200 -c 1 -o gpurun_out/prof_c2 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c2.log 2>&1; echo "ncu c2 exit $?"
ls -la gpurun_out/*.ncu-rep
