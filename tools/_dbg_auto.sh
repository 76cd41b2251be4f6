cd /root/repo
export DTF_GPU_INDEX=0 DTF_FABRIC=1 DTF_FABRIC_PORT_OFFSET=1500 DTF_DEBUG=1 PYTHONPATH=/root/repo
H="--ps_hosts=127.0.0.1:22420 --worker_hosts=127.0.0.1:22421,127.0.0.1:22422"
A="--train_steps=600 --validate_every=200 --train_dir=/tmp/dbg_ckpt --issync=True --log_every=200"
rm -rf /tmp/dbg_ckpt
timeout 150 python -u examples/distributed_mnist.py --job_name=ps --task_index=0 $H $A > gpurun_out/dbg_ps.log 2>&1 &
timeout 150 python -u examples/distributed_mnist.py --job_name=worker --task_index=1 $H $A > gpurun_out/dbg_w1.log 2>&1 &
timeout 140 python -u examples/distributed_mnist.py --job_name=worker --task_index=0 $H $A > gpurun_out/dbg_w0.log 2>&1
echo "w0 rc=$?"
sleep 2
for f in ps w0 w1; do echo "=== $f"; tail -25 gpurun_out/dbg_$f.log | cut -c1-300; done
