step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-600)"; }
step t_gpu_sel 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn_fused.py tests/test_gpu_auto_fabric.py tests/test_gpu_engine.py -m gpu -x -q
step resnet1_splitk 300 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
step racecheck_gemm 240 compute-sanitizer --tool racecheck --error-exitcode 3 --launch-timeout 60 python -m pytest tests/test_gpu_kernels.py -m gpu -k "(all_operand_majors and 100-100-784) or tf32_persistent" -x -q
tail -1 gpurun_out/resnet1_splitk.log > gpurun_out/resnet1_splitk.json
