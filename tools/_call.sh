mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "persistent" > gpurun_out/t_persist.log 2>&1; echo "persist rc=$?"; tail -6 gpurun_out/t_persist.log
timeout 200 python tools/gemm_perf.py > gpurun_out/gemm_perf.log 2>&1; echo "perf rc=$?"; tail -8 gpurun_out/gemm_perf.log
timeout 200 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "engine rc=$?"; tail -6 gpurun_out/t_engine.log
timeout 120 python bench.py --gpus 1 --steps 1000 --warmup 10 --e2e-steps 500 > gpurun_out/b1.log 2>&1; echo "b1 rc=$?"; tail -1 gpurun_out/b1.log
