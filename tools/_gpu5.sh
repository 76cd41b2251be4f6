# final 1-GPU battery: full gpu tier (PDL on), bench with / without PDL, ResNet, conv perf + ncu captures, sanitizer
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-420)"; }
step t_gpu_full 500 python -m pytest tests -m gpu -x -q
step bench1_pdl 200 python bench.py
DTF_PDL=0 step bench1_nopdl 200 python bench.py --baseline 0
step resnet1 200 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
step conv_perf 200 python tools/ncu_conv.py
step ncu_conv 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05_kernel --launch-skip 0 -c 3 -f -o gpurun_out/prof_conv python tools/ncu_conv.py
step ncu_step 300 ncu --set full --clock-control none --import-source on -k regex:"mlp_step|ps_apply" --launch-skip 12 -c 4 -f -o gpurun_out/prof_step3 python tools/ncu_step.py 12
step sanitize_memcheck_step 240 compute-sanitizer --tool memcheck --error-exitcode 3 --launch-timeout 60 python -m pytest tests/test_gpu_step_kernel.py tests/test_gpu_conv_implicit.py -m gpu -k "100-784 or two_consecutive or 8-32-32 or 8-4-4" -x -q
for f in bench1_pdl bench1_nopdl resnet1; do tail -1 gpurun_out/$f.log > gpurun_out/$f.json; done
