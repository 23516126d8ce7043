mkdir -p gpurun_out
KF='regex:dense_layer'
for P in 0 1 2 3; do
  EML_DENSE_L2PROMO=$P timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k "$KF" -s 96 -c 32 --csv --log-file gpurun_out/promo_$P.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/promo_$P.log 2>&1; echo "promo $P exit $?"
done
