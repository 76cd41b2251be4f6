#!/bin/bash
# Standard measurement battery for one gpurun call; everything lands in gpurun_out/ (copy what matters to profiles/).
#   gpurun --timeout 900 -- 'bash tools/gpu_battery.sh single'            # 1 GPU : tests, bench, GEMM table, ncu of the pair GEMM
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_battery.sh multi 2'  # N GPUs: engine vs oracle, fabric GB/s, bench (NVLS + unicast)
#   gpurun --timeout 600 -- 'bash tools/gpu_battery.sh pending'           # 1 GPU : first hardware run of everything written blind
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_battery.sh pending-multi 2'   # N GPUs: ps_on_workers topology vs the oracle + bench
# Each step has its own timeout and logs to its own file, so one failure does not hide the others.
set -u
mode=${1:-single}
N=${2:-2}
mkdir -p gpurun_out
step() { local name=$1 t=$2; shift 2; timeout "$t" "$@" > "gpurun_out/${name}.log" 2>&1; echo "${name}: rc=$? $(tail -1 gpurun_out/${name}.log | cut -c1-300)"; }
if [ "$mode" = "single" ]; then
  step t_gpu 300 python -m pytest tests -x -q -m gpu
  step bench_n1 120 python bench.py
  step gemm_perf 240 python tools/gemm_perf.py
  step phase_trace 120 python tools/phase_trace.py
  step ncu_pair 120 ncu --set full --clock-control none --import-source on -k regex:persistent --launch-skip 2 -c 1 -f \
      -o gpurun_out/prof_gemm_pair python tools/ncu_gemm.py
elif [ "$mode" = "pending" ]; then
  # kernels / paths written after round 1's GPU budget was spent (validated under the host emulation only)
  DTF_TEST_UNVALIDATED=1 step t_nn_fused 300 python -m pytest tests/test_gpu_nn_fused.py -q -m gpu
  step t_pipeline 120 python -m pytest tests/test_gpu_pipeline.py -q -m gpu
  step t_gpu 300 python -m pytest tests -x -q -m gpu
  step bench_n1 180 python bench.py
  step resnet18_eager 200 python bench.py --model resnet18 --steps 10 --warmup 3
  DTF_FUSED_NN=1 step resnet18_fused 200 python bench.py --model resnet18 --steps 10 --warmup 3
  step nn_perf 200 python tools/nn_perf.py
  step ncu_nn 240 ncu --set full --clock-control none --import-source on -k "regex:bn_|im2col_nhwc_vec8|col2im_nhwc_vec4|pool" -c 12 -f \
      -o gpurun_out/prof_nn python tools/nn_perf.py --ncu
  step resnet18_graph 200 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
  DTF_FUSED_NN=1 step resnet18_fused_graph 200 python bench.py --model resnet18 --steps 10 --warmup 4 --graph-step 1
elif [ "$mode" = "pending-multi" ]; then
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
  DTF_PS_ON_WORKERS=1 step mp_check_pow 120 $TR tools/mp_check.py
  DTF_PS_ON_WORKERS=1 DTF_NVLS=1 step mp_check_pow_nvls 120 $TR tools/mp_check.py
  step bench_pow_nvls 150 $TR bench.py --gpus $N --nvls on --ps-on-workers 1
  step bench_pow_unicast 150 $TR bench.py --gpus $N --nvls off --ps-on-workers 1
  step bench_ref 150 $TR bench.py --gpus $N
  step bench_e2e_pipe 150 $TR bench.py --gpus $N --e2e-pipeline 2
  step resnet18_pow 200 $TR bench.py --gpus $N --model resnet18 --steps 10 --warmup 3 --ps-on-workers 1
  DTF_FUSED_NN=1 step resnet18_pow_fused_graph 200 $TR bench.py --gpus $N --model resnet18 --steps 10 --warmup 4 --ps-on-workers 1 --graph-step 1
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
  DTF_NVLS=1 step mp_check_nvls 120 $TR tools/mp_check.py
  step mp_check 120 $TR tools/mp_check.py
  step nvls_check 120 $TR tools/nvls_check.py --iters 10
  step bench_nvls 150 $TR bench.py --gpus $N --nvls on
  step bench_unicast 150 $TR bench.py --gpus $N --nvls off
  step resnet18 200 $TR bench.py --gpus $N --model resnet18 --steps 10 --warmup 3
fi
