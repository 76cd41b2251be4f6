N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-400)"; }
step bench8_full 300 $TR bench.py --gpus 8
step nvls_check8 200 $TR tools/nvls_check.py --iters 10
DTF_PS_ON_WORKERS=1 DTF_NVLS=1 step mp8_pow_nvls_tf32 150 $TR tools/mp_check.py
step bench8_async 200 $TR bench.py --gpus 8 --mode async --baseline 0 --e2e-steps 0
step bench8_psonly 200 $TR bench.py --gpus 8 --ps-only-task 1 --optimizer sgd --baseline 0 --e2e-steps 0
step bench8_ingraph_2ps_adam_bf16 240 python bench.py --gpus 8 --in-graph --num-ps 2 --ps-only-task 1 --optimizer adam --precision bf16 --baseline 0
step resnet8 300 $TR bench.py --gpus 8 --model resnet18 --steps 10 --warmup 4 --graph-step 1
step mp_trace8 150 $TR tools/mp_trace.py
for f in bench8_full nvls_check8 bench8_async bench8_psonly bench8_ingraph_2ps_adam_bf16 resnet8; do tail -1 gpurun_out/$f.log > gpurun_out/$f.json; done
