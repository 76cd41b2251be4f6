N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-400)"; }
step bench8_final 200 $TR bench.py --gpus 8
step resnet8_final 200 $TR bench.py --gpus 8 --model resnet18 --steps 10 --warmup 4 --graph-step 1 --baseline 0
for f in bench8_final resnet8_final; do tail -1 gpurun_out/$f.log > gpurun_out/$f.json; done
