mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_kernel -s 2 -c 1 -o gpurun_out/prof_sinkhorn -f python tools/profile_other_kernels.py --what sinkhorn > gpurun_out/ncu_o1.log 2>&1; echo "sinkhorn exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sg_render_fwd -s 2 -c 1 -o gpurun_out/prof_render -f python tools/profile_other_kernels.py --what render > gpurun_out/ncu_o2.log 2>&1; echo "render exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 5 -c 1 -o gpurun_out/prof_needlet_gemm -f python tools/profile_other_kernels.py --what needlets > gpurun_out/ncu_o3.log 2>&1; echo "needlets exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 150 -c 1 -o gpurun_out/prof_generator_gemm -f python tools/profile_other_kernels.py --what generator > gpurun_out/ncu_o4.log 2>&1; echo "generator exit $?"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; tail -c 900 gpurun_out/bench_quick.json
ls -la gpurun_out/*.ncu-rep
