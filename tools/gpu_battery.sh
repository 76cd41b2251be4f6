#!/bin/bash
# Measurement batteries for `gpurun` (one call = one fresh B200 box; batch the work, results land in gpurun_out/):
#   gpurun --timeout 1500 -- 'bash tools/gpu_battery.sh single'            # 1 GPU : gpu test tier, MNIST + ResNet bench, conv table
#   gpurun --timeout 1500 -- 'bash tools/gpu_battery.sh profile'           # 1 GPU : ncu --set full of the hot kernels + ResNet launch list
#   gpurun --timeout 1500 -- 'bash tools/gpu_battery.sh sanitize'          # 1 GPU : compute-sanitizer tier (tools/sanitize.sh)
#   gpurun --timeout 900  -- 'bash tools/gpu_battery.sh recovery'          # 1 GPU : tf.train programs on the fabric, kill-and-restart recovery
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_battery.sh multi N'  # N GPUs: oracle checks, benches, fabric GB/s, per-rank step trace
# Every step runs under its own `timeout` (a hung kernel must not take the box down: strikes close gpurun) and prints one line.
# Afterwards, here: python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep ; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
mode=${1:-single}
N=${2:-2}
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-400)"; }
keep_json() { for f in "$@"; do tail -1 "gpurun_out/$f.log" > "gpurun_out/$f.json"; done; }

case "$mode" in
single)
  step t_gpu_full 600 python -m pytest tests -m gpu -q
  step bench1 200 python bench.py
  DTF_PDL=0 step bench1_nopdl 200 python bench.py --baseline 0
  step bench1_bf16 200 python bench.py --precision bf16 --baseline 0
  step resnet1 300 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
  step conv_perf 200 python tools/ncu_conv.py
  keep_json bench1 bench1_nopdl bench1_bf16 resnet1
  ;;
profile)
  step ncu_step 300 ncu --set full --clock-control none --import-source on -k regex:"mlp_step|ps_apply" --launch-skip 12 -c 4 -f -o gpurun_out/prof_step python tools/ncu_step.py 12
  step ncu_conv 300 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tcgen05 --launch-skip 0 -c 6 -f -o gpurun_out/prof_conv python tools/ncu_conv.py
  step ncu_gemm 300 ncu --set full --clock-control none --import-source on -k regex:persistent --launch-skip 2 -c 1 -f -o gpurun_out/prof_gemm_pair python tools/ncu_gemm.py
  step ncu_nn 300 ncu --set full --clock-control none --import-source on -k regex:"bn_|pool" -c 12 -f -o gpurun_out/prof_nn python tools/nn_perf.py --ncu
  step resnet_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/resnet_launches.csv -c 4000 python bench.py --model resnet18 --steps 1 --warmup 1 --graph-step 0 --baseline 0 --min-ms 0 --max-reps 1
  step step_trace 120 python tools/step_trace.py
  step gemm_perf 300 python tools/gemm_perf.py
  ;;
sanitize)
  bash tools/sanitize.sh
  ;;
recovery)
  step t_fabric_programs 600 python -m pytest tests/test_gpu_auto_fabric.py tests/test_gpu_fabric_recovery.py -m gpu -q
  ;;
multi)
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
  step t_multi 500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -m gpu -q
  step bench${N} 300 $TR bench.py --gpus $N
  step resnet${N} 300 $TR bench.py --gpus $N --model resnet18 --steps 10 --warmup 4 --graph-step 1
  step nvls_check${N} 200 $TR tools/nvls_check.py --iters 10
  DTF_PS_ON_WORKERS=1 DTF_NVLS=1 step mp${N}_pow_nvls_tf32 150 $TR tools/mp_check.py
  DTF_PS_ON_WORKERS=1 DTF_NVLS=1 DTF_PRECISION=bf16 step mp${N}_pow_nvls_bf16 150 $TR tools/mp_check.py
  step mp_trace${N} 150 $TR tools/mp_trace.py
  step bench${N}_async 200 $TR bench.py --gpus $N --mode async --baseline 0 --e2e-steps 0
  step bench${N}_psonly 200 $TR bench.py --gpus $N --ps-only-task 1 --baseline 0 --e2e-steps 0
  step bench${N}_ingraph_2ps_adam_bf16 240 python bench.py --gpus $N --in-graph --num-ps 2 --ps-only-task 1 --optimizer adam --precision bf16 --baseline 0
  keep_json bench${N} resnet${N} nvls_check${N} bench${N}_async bench${N}_psonly bench${N}_ingraph_2ps_adam_bf16
  ;;
*)
  echo "usage: $0 single | profile | sanitize | recovery | multi N" >&2
  exit 2
  ;;
esac
