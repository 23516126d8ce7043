mkdir -p gpurun_out
for P in 0 1 3; do
  EML_DENSE_DBG=$P timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dense_layer -s 96 -c 32 --csv --log-file gpurun_out/dbg_$P.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/dbg_$P.log 2>&1; echo "dbg $P exit $?"
done
