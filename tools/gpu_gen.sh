mkdir -p gpurun_out
python tools/bench_generator.py --batch 4 --precision bf16x3 > gpurun_out/gen_bench.log 2>&1; tail -2 gpurun_out/gen_bench.log
python tools/bench_generator.py --batch 4 --precision bf16 >> gpurun_out/gen_bench.log 2>&1; tail -1 gpurun_out/gen_bench.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gen_launches.csv python tools/bench_generator.py --batch 2 --steps 1 > /dev/null 2>&1; echo ncu $?
python - <<'PY'
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/gen_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    n=re.sub(r'\(.*','',r[ki])[:60]; agg[n]+=float(r[vi].replace(',',''))/1e6; cnt[n]+=1
tot=sum(agg.values())
for n,v in agg.most_common(12): print('%-62s x%5d %9.2f ms %5.1f%%'%(n,cnt[n],v,100*v/tot))
print('total',tot)
PY
