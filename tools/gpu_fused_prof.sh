mkdir -p gpurun_out
# dense-layer launches inside one bench step: -s index counts only kernels matching -k.  warmup 3 + roofline pass... capture from the timed step
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 96 -c 1 -o gpurun_out/prof_g0 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g0.log 2>&1; echo "ncu g0 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_g10 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_g10.log 2>&1; echo "ncu g10 exit $?"
ls -la gpurun_out/*.ncu-rep
