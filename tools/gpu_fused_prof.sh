mkdir -p gpurun_out
KF='regex:dense_layer|conv|stem_kernel|head_pool|linear_kernel|sg_render'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -s 216 -c 72 --csv --log-file gpurun_out/launches_fused.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
# layer index 10 of block 1 (C_in = 144, 3 chunks) and layer 2 (C_in = 48, 1 chunk)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_f10 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f10.log 2>&1; echo "ncu f10 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 98 -c 1 -o gpurun_out/prof_f2 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1; echo "ncu f2 exit $?"
ls -la gpurun_out
