# one-GPU battery: full gpu test tier, ncu of the step / apply kernels, ResNet launch list, sanitizer engine rerun
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-300)"; }
step t_gpu_full 400 python -m pytest tests -m gpu -x -q
step ncu_step 300 ncu --set full --clock-control none --import-source on -k regex:"mlp_step|ps_apply" --launch-skip 12 -c 4 -f -o gpurun_out/prof_step2 python tools/ncu_step.py 12
step resnet_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/resnet_launches.csv -c 4000 python bench.py --model resnet18 --steps 1 --warmup 1 --graph-step 0 --baseline 0 --e2e-steps 0 --min-ms 0 --max-reps 1
step sanitize_memcheck_engine 240 compute-sanitizer --tool memcheck --error-exitcode 3 --launch-timeout 60 python -m pytest tests/test_gpu_engine.py -m gpu -k "tf32 and sgd" -x -q
step bench1 200 python bench.py
tail -1 gpurun_out/bench1.log > gpurun_out/bench1_full.json
