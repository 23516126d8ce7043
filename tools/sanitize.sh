# compute-sanitizer over the hot kernels at small shapes.  Usage under gpurun:  bash tools/sanitize.sh [round tag]
#   memcheck   : out-of-bounds / misaligned global + shared accesses               (all sections)
#   racecheck  : shared-memory hazards of the warp-specialised kernels             (dense layer / convs / GEMM / Sinkhorn)
#   initcheck  : reads of global memory nobody wrote (caching allocator off, so every tensor is its own cudaMalloc)
# Logs: gpurun_out/sanitize_<tool>_<section>.log; the summary lines are collected into gpurun_out/sanitize_summary.txt
mkdir -p gpurun_out
S=gpurun_out/sanitize_summary.txt; : > $S
run() {   # tool section timeout extra-env
  local tool=$1 what=$2 to=$3; shift 3
  local log=gpurun_out/sanitize_${tool}_${what}.log
  env "$@" timeout $to compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_driver.py $what > $log 2>&1
  local rc=$?
  echo "$tool $what rc=$rc : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|driver done' $log | tr '\n' ' ')" | tee -a $S
}
for w in densenet_eval densenet_train sinkhorn gemm generator_train gan_step; do run memcheck $w 900; done
for w in densenet_eval sinkhorn gemm densenet_train; do run racecheck $w 1200; done
for w in generator_train densenet_train densenet_eval gemm; do run initcheck $w 900 PYTORCH_NO_CUDA_MEMORY_CACHING=1; done
