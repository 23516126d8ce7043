mkdir -p gpurun_out
timeout 120 tools/micro/ts_mma_test > gpurun_out/ts_mma_test.log 2>&1; echo "ts_mma exit $?"; cat gpurun_out/ts_mma_test.log
timeout 600 python tests/debug_gp_bwd.py 4 2 > gpurun_out/gp_bwd_debug2.log 2>&1; echo "debug exit $?"
python - <<'PY'
import re
sec=None; worst={}
for line in open('gpurun_out/gp_bwd_debug2.log'):
    if line.startswith('===='):
        sec=line.strip(); worst[sec]=[]; print(sec); continue
    m=re.search(r'^(\S+).*err_vs_fp64 (\S+)', line)
    if m and sec and float(m.group(2))<0.9: worst[sec].append((float(m.group(2)), m.group(1)))
for s,v in worst.items():
    v.sort(reverse=True); print(s); print('   worst:', v[:4]); print('   median: %.3e' % v[len(v)//2][0])
PY
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_all.log
