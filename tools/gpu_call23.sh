mkdir -p gpurun_out
EML_CPROFILE=1 timeout 600 python tools/profile_gan_step.py --ngf 64 --ndf 64 --batch 4 > gpurun_out/gan_cprofile.log 2>&1; echo "exit $?"; tail -90 gpurun_out/gan_cprofile.log | cut -c1-180
for s in 256 256; do echo "fwd: $(timeout 300 python tools/fwd_time.py 256 2>&1 | tail -1)"; done
echo "serial fc: $(EML_FC_SERIAL=1 timeout 300 python tools/fwd_time.py 256 2>&1 | tail -1)"
timeout 600 python -m pytest tests/test_densenet_gpu.py tests/test_experiments_gpu.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -3
