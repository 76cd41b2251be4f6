step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-420)"; }
step t_gpu_full 500 python -m pytest tests -m gpu -q
step conv_perf 200 python tools/ncu_conv.py
cp gpurun_out/conv_perf.json gpurun_out/conv_perf_v2.json
step resnet1 200 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
step bench1 200 python bench.py
for f in bench1 resnet1; do tail -1 gpurun_out/$f.log > gpurun_out/$f.json; done
