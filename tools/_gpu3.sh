step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-600)"; }
step t_conv 240 python -m pytest tests/test_gpu_conv_implicit.py -m gpu -q
step t_nn 300 python -m pytest tests/test_gpu_nn_fused.py tests/test_gpu_auto_fabric.py -m gpu -x -q -k "resnet or twin"
step resnet1_igemm 300 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
tail -1 gpurun_out/resnet1_igemm.log > gpurun_out/resnet1_igemm.json
