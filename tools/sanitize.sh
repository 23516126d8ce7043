# compute-sanitizer over the hot kernels at small shapes.  Usage under gpurun:  bash tools/sanitize.sh
#   memcheck   : out-of-bounds / misaligned global + shared accesses               (all sections)
#   racecheck  : shared-memory hazards of the warp-specialised kernels             (dense layer / convs / GEMM / Sinkhorn)
#   initcheck  : reads of global memory nobody wrote (caching allocator off, so every tensor is its own cudaMalloc)
# Logs: gpurun_out/sanitize_<tool>_<section>.log; the summary lines are collected into gpurun_out/sanitize_summary.txt
mkdir -p gpurun_out
S=gpurun_out/sanitize_summary.txt; : > $S
run() {   # tool section timeout extra-env
  local tool=$1 what=$2 to=$3; shift 3
  local log=gpurun_out/sanitize_${tool}_${what}.log
  # --report-api-errors no: the first launch out of our libcudart makes the runtime probe cuKernelGetFunction with a handle of torch's
  # own runtime instance (CUDA_ERROR_INVALID_HANDLE, handled inside cudaLaunchKernel, the launch succeeds); device-side errors stay on
  env "$@" timeout $to compute-sanitizer --tool $tool --report-api-errors no --print-limit 100 --error-exitcode 9 python tools/sanitize_driver.py $what > $log 2>&1
  local rc=$?
  echo "$tool $what rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|driver done' $log | tr '\n' ' ')" | tee -a $S
}
PARTS=${SAN_PARTS:-memcheck racecheck initcheck poison}
case " $PARTS " in *" memcheck "*) for w in densenet_eval densenet_train sinkhorn gemm generator_train; do run memcheck $w 400 X=1; done;; esac
case " $PARTS " in *" racecheck "*) for w in densenet_eval sinkhorn gemm; do run racecheck $w 500 X=1; done;; esac
case " $PARTS " in *" initcheck "*) for w in densenet_eval generator_train densenet_train; do run initcheck $w 1200 PYTORCH_NO_CUDA_MEMORY_CACHING=1; done;; esac
case " $PARTS " in *" poison "*) ;; *) exit 0;; esac
for w in densenet_eval densenet_train generator_train gan_step; do
  timeout 300 python tools/sanitize_driver.py $w --poison > gpurun_out/poison_$w.log 2>&1; echo "poison $w rc=$? : $(tail -1 gpurun_out/poison_$w.log)" | tee -a $S
done
