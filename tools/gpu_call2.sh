mkdir -p gpurun_out
timeout 300 python tools/sanitize_driver.py generator_train --poison > gpurun_out/poison_generator.log 2>&1; echo "poison generator exit $?"; grep -v "finite$" gpurun_out/poison_generator.log | tail -40
timeout 300 python tools/sanitize_driver.py gan_step --poison > gpurun_out/poison_gan.log 2>&1; echo "poison gan exit $?"; tail -12 gpurun_out/poison_gan.log
timeout 300 python tools/sanitize_driver.py densenet_train --poison > gpurun_out/poison_dn.log 2>&1; echo "poison densenet_train exit $?"; tail -5 gpurun_out/poison_dn.log
timeout 300 python tools/sanitize_driver.py densenet_eval --poison > gpurun_out/poison_dne.log 2>&1; echo "poison densenet_eval exit $?"; tail -5 gpurun_out/poison_dne.log
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 900 compute-sanitizer --tool initcheck --print-limit 30 python tools/sanitize_driver.py generator_train > gpurun_out/sanitize_initcheck_generator_train.log 2>&1; echo "initcheck exit $?"
grep -E "Uninitialized|at .*\(|ERROR SUMMARY" gpurun_out/sanitize_initcheck_generator_train.log | head -40
