mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_dense_rs_c144 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 111 -c 1 -o gpurun_out/prof_dense_rs_c204 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full204.log 2>&1; echo "ncu full c204 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 96 -c 1 -o gpurun_out/prof_dense_rs_c24 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full24.log 2>&1; echo "ncu full c24 exit $?"
ls -la gpurun_out/*rs*.ncu-rep
