# 2-GPU battery
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-500)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
step t_recovery 500 python -m pytest tests/test_gpu_fabric_recovery.py tests/test_gpu_auto_fabric.py -m gpu -x -q -k "recovers or (True and twin)"
step t_multi 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -m gpu -x -q
step bench2 200 $TR bench.py --gpus 2
step resnet2_ps2 300 $TR bench.py --gpus 2 --model resnet18 --steps 10 --warmup 4 --graph-step 1
step resnet2_ps1 300 $TR bench.py --gpus 2 --model resnet18 --steps 10 --warmup 4 --graph-step 1 --num-ps 1 --baseline 0
for f in bench2 resnet2_ps2 resnet2_ps1; do tail -1 gpurun_out/$f.log > gpurun_out/$f.json; done
