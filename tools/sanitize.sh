#!/bin/bash
# Sanitizer tier: the single-GPU kernel tests under compute-sanitizer.
#   gpurun --timeout 1200 -- 'bash tools/sanitize.sh'        -> gpurun_out/sanitize_<tool>_<suite>.log + sanitize_summary.txt
# memcheck: out-of-bounds / misaligned accesses (TMA boxes, epilogue tails, bulk stores); racecheck: shared-memory hazards
# between the producer / MMA / epilogue warps and inside the fused head; synccheck: barrier misuse.  tcgen05 / TMA traffic
# goes through the async proxy, which the tools only partly model: a clean run is necessary, not sufficient.
set -u
mkdir -p gpurun_out
: > gpurun_out/sanitize_summary.txt
run() {  # tool suite timeout pytest-args...
  local tool=$1 suite=$2 t=$3; shift 3
  local log=gpurun_out/sanitize_${tool}_${suite}.log
  timeout "$t" compute-sanitizer --tool "${tool}" --error-exitcode 3 --launch-timeout 60 python -m pytest "$@" -x -q > "${log}" 2>&1
  local rc=$?
  local errs=$(grep -c "========= .*[Ee]rror\|========= Invalid\|========= Race\|========= Hazard" "${log}" || true)
  echo "${tool} ${suite}: rc=${rc} reported=${errs} | $(grep -E "passed|failed" "${log}" | tail -1) | $(grep "ERROR SUMMARY\|RACECHECK SUMMARY" "${log}" | tail -1)" | tee -a gpurun_out/sanitize_summary.txt
}
for tool in memcheck racecheck synccheck; do
  run ${tool} step 240 tests/test_gpu_step_kernel.py -m gpu -k "100-784 or two_consecutive"
  run ${tool} gemm 240 tests/test_gpu_kernels.py -m gpu -k "(all_operand_majors and 100-100-784) or xent or optimizer or tf32_persistent"
  run ${tool} nn 200 tests/test_gpu_nn_fused.py -m gpu -k "fused_bn and 1000 or pooling"
done
run memcheck engine 240 tests/test_gpu_engine.py -m gpu -k "tf32 and sgd"
