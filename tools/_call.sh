mkdir -p gpurun_out
timeout 80 python -m pytest tests -x -q -m gpu > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/t_gpu.log
timeout 60 python bench.py > gpurun_out/b1_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/b1_default.log
timeout 60 ncu --set full --clock-control none --import-source on -k regex:persistent --launch-skip 2 -c 1 -f -o gpurun_out/prof_gemm_pair python tools/ncu_gemm.py > gpurun_out/ncu_pair.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_pair.log
