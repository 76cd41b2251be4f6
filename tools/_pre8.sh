TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-700)"; }
step resnet1_base 200 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
step ingraph2_tf32 200 python bench.py --gpus 2 --in-graph --baseline 0
step ingraph2_bf16_adam 200 python bench.py --gpus 2 --in-graph --precision bf16 --optimizer adam --ps-only-task 1 --baseline 0
step async2 200 $TR bench.py --gpus 2 --mode async --baseline 0 --e2e-steps 0
step async2_psonly 200 $TR bench.py --gpus 2 --mode async --ps-only-task 1 --baseline 0 --e2e-steps 0
