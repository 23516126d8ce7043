mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 96 -c 1 -o gpurun_out/prof_h0 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_h0.log 2>&1; echo "ncu h0 exit $?"
