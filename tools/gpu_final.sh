# full evidence run (round 2): gpu tests, smoke, bench (+cpu baseline, eager baseline, training workloads), reference arm, the ncu launch list
# with DRAM bytes of one headline step and of one training step, ncu --set full of the dominant kernel
R=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider > gpurun_out/pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -2 gpurun_out/bench_n1.err
tail -c 4000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 700 gpurun_out/bench_ref.json
KF='regex:dense_layer|conv|stem_kernel|head_pool|linear_kernel|sg_render|gemm_tma|split_bf16'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KF" -s 160 -c 170 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python tools/launch_traffic.py gpurun_out/launches.csv gpurun_out/${R}_traffic.json > /dev/null; echo "traffic exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_layer -s 106 -c 1 -o gpurun_out/prof_dense_c144 -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
# one training step (B = 64): per-entry-point device time and the ncu launch list
timeout 600 python tools/profile_train.py 64 > gpurun_out/profile_train_b64.log 2>&1; echo "profile_train exit $?"; head -30 gpurun_out/profile_train_b64.log
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/launches_train_b64.csv \
    python tools/profile_train.py 64 > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit $?"
ls -la gpurun_out | head -40
# GenProjector: generator forward (eval) per-shape profile, and the G + D iteration
timeout 600 python tools/bench_generator.py --batch 16 --profile > gpurun_out/profile_generator_fwd_b16.log 2>&1; echo "gen exit $?"; tail -1 gpurun_out/profile_generator_fwd_b16.log
timeout 600 python tools/bench_generator.py --batch 16 --precision bf16 > gpurun_out/generator_fwd_b16_bf16.log 2>&1; tail -1 gpurun_out/generator_fwd_b16_bf16.log
timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/profile_gan_step_b4_ngf64.log 2>&1; echo "gan exit $?"; tail -2 gpurun_out/profile_gan_step_b4_ngf64.log
