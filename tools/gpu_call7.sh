mkdir -p gpurun_out
python tools/layer_times.py 256 > gpurun_out/layer_times_ts.log 2>&1; echo "layer times exit $?"; cat gpurun_out/layer_times_ts.log
EML_DENSE_CW=8 python tools/layer_times.py 256 > gpurun_out/layer_times_cw8.log 2>&1; echo "layer times cw8 exit $?"; grep -E "dense_layer|sum" gpurun_out/layer_times_cw8.log | awk '{print $2, $5}' | tr '\n' ' '
echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_dense_ts_c144 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
EML_DENSE_CW=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_dense_ts8_c144 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full8.log 2>&1; echo "ncu full cw8 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 96 -c 1 -o gpurun_out/prof_dense_ts_c24 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full24.log 2>&1; echo "ncu full c24 exit $?"
ls -la gpurun_out/*.ncu-rep
