N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-1200)"; }
step nvls_check${N} 200 $TR tools/nvls_check.py --iters 10
DTF_PS_ON_WORKERS=1 DTF_NVLS=1 step mp_pow_nvls_tf32 120 $TR tools/mp_check.py
DTF_PS_ON_WORKERS=1 DTF_NVLS=1 DTF_PRECISION=bf16 step mp_pow_nvls_bf16 120 $TR tools/mp_check.py
step bench${N}_nvls 200 $TR bench.py --gpus $N --nvls on --baseline 0 --e2e-steps 0
step resnet${N} 200 $TR bench.py --gpus $N --model resnet18 --steps 10 --warmup 4 --graph-step 1
