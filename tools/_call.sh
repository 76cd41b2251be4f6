mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 70 $TR bench.py --gpus 8 --steps 2000 --warmup 10 --e2e-steps 400 --nvls on > gpurun_out/b8_nvls.log 2>&1; echo "b8_nvls rc=$?"; tail -1 gpurun_out/b8_nvls.log
timeout 70 $TR bench.py --gpus 8 --steps 2000 --warmup 10 --e2e-steps 400 > gpurun_out/b8_off.log 2>&1; echo "b8_off rc=$?"; tail -1 gpurun_out/b8_off.log
timeout 50 $TR tools/nvls_check.py --iters 10 > gpurun_out/nvls8_mp.log 2>&1; echo "nvls8 rc=$?"; tail -1 gpurun_out/nvls8_mp.log
