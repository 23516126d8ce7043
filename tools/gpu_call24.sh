mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread
timeout 1200 ncu --metrics $M --clock-control none -k regex:"bn_bwd|wgrad1x1|dense_bwd1_kernel|conv1x1_persist|conv3x3_roll|wgrad3x3" -s 700 -c 330 --csv --log-file gpurun_out/train_kernels_metrics.csv \
    python tools/profile_train.py 64 > gpurun_out/ncu_train2.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_train2.log
wc -l gpurun_out/train_kernels_metrics.csv
