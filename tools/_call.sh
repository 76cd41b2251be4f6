mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 120 python tools/nvls_check.py --gpus 2 > gpurun_out/nvls_sp.log 2>&1; echo "nvls_sp rc=$?"; tail -3 gpurun_out/nvls_sp.log
timeout 120 $TR tools/nvls_check.py > gpurun_out/nvls_mp.log 2>&1; echo "nvls_mp rc=$?"; tail -3 gpurun_out/nvls_mp.log
DTF_NVLS=1 timeout 120 $TR tools/mp_check.py > gpurun_out/mpcheck_nvls.log 2>&1; echo "mpcheck_nvls rc=$?"; grep MP_CHECK gpurun_out/mpcheck_nvls.log || tail -5 gpurun_out/mpcheck_nvls.log
timeout 120 $TR bench.py --gpus 2 --steps 1000 --warmup 10 --e2e-steps 300 > gpurun_out/b2_off.log 2>&1; echo "b2_off rc=$?"; tail -1 gpurun_out/b2_off.log
timeout 120 $TR bench.py --gpus 2 --steps 1000 --warmup 10 --e2e-steps 300 --nvls on > gpurun_out/b2_nvls.log 2>&1; echo "b2_nvls rc=$?"; tail -1 gpurun_out/b2_nvls.log
timeout 120 python bench.py --gpus 1 --steps 1000 --warmup 10 --e2e-steps 300 > gpurun_out/b1.log 2>&1; echo "b1 rc=$?"; tail -1 gpurun_out/b1.log
