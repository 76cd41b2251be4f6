#!/bin/bash
# Sanitizer tier: run the single-GPU kernel tests under compute-sanitizer.
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'        -> gpurun_out/sanitize_<tool>.log (+ a one-line summary each)
# memcheck: out-of-bounds / misaligned accesses (TMA boxes, epilogue tails); racecheck: shared-memory hazards between
# the producer / MMA / epilogue warps; synccheck: barrier misuse.  tcgen05/TMA traffic goes through the async proxy,
# which the tools only partly model: a clean run is necessary, not sufficient.
set -u
mkdir -p gpurun_out
TESTS=(tests/test_gpu_kernels.py -k "gemm or xent or optimizer or small or conv")
# the fused NN kernels (plain bandwidth kernels: the tools model them fully) -- their hardware tests are still gated
NN_TESTS=(tests/test_gpu_nn_fused.py -k "fused_bn or pooling or im2col")
for tool in memcheck racecheck synccheck; do
  DTF_TEST_UNVALIDATED=1 timeout 200 compute-sanitizer --tool ${tool} --error-exitcode 3 --launch-timeout 60 \
      python -m pytest "${NN_TESTS[@]}" -x -q > gpurun_out/sanitize_nn_${tool}.log 2>&1
  echo "nn ${tool}: rc=$? $(tail -1 gpurun_out/sanitize_nn_${tool}.log)" | tee -a gpurun_out/sanitize_summary.txt
  log=gpurun_out/sanitize_${tool}.log
  timeout 280 compute-sanitizer --tool ${tool} --error-exitcode 3 --launch-timeout 60 \
      python -m pytest "${TESTS[@]}" -x -q > ${log} 2>&1
  rc=$?
  errs=$(grep -c "========= .*error\|========= Invalid\|========= Race" ${log} || true)
  echo "${tool}: rc=${rc} reported=${errs} $(tail -1 ${log})" | tee -a gpurun_out/sanitize_summary.txt
done
