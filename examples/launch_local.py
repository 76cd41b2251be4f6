"""Launch a whole ps/worker cluster of one example script on this machine (one process per task).

    python examples/launch_local.py examples/distributed_mnist.py --num_ps 1 --num_workers 2 -- --issync=True --train_steps=200

Each task gets ``--job_name/--task_index/--ps_hosts/--worker_hosts``; with GPUs present
task n is bound to GPU ``n % num_gpus`` through ``DTF_GPU_INDEX`` (ps tasks first).
Workers' exit ends the run; ps processes are then terminated (they ``join()`` forever).  In-graph programs
(``example_in_graph.py``, ``example_distributed_server.py``) have ONE client, worker 0: run them with ``--wait first``.
"""
import argparse
import os
import signal
import socket
import subprocess
import sys
import time


def free_ports(n):
    socks, ports = [], []
    for _ in range(n):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        ports.append(s.getsockname()[1])
        socks.append(s)
    for s in socks:
        s.close()
    return ports


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--num_ps", type=int, default=1)
    ap.add_argument("--num_workers", type=int, default=2)
    ap.add_argument("--timeout", type=float, default=600)
    ap.add_argument("--gpus", type=int, default=-1, help="GPUs to spread tasks over (-1: all visible, 0: none)")
    ap.add_argument("--wait", default="all", choices=["all", "first"],
                    help="all: the run ends when every worker exits (between-graph); first: when worker 0 exits "
                         "(in-graph replication: worker 0 is the only client, the other workers just serve)")
    # launcher options may come before or after the script; everything after a literal "--" goes to the script
    argv = sys.argv[1:]
    extra = []
    if "--" in argv:
        k = argv.index("--")
        argv, extra = argv[:k], argv[k + 1:]
    a = ap.parse_args(argv)
    ports = free_ports(a.num_ps + a.num_workers)
    ps_hosts = ",".join("127.0.0.1:%d" % p for p in ports[:a.num_ps])
    wk_hosts = ",".join("127.0.0.1:%d" % p for p in ports[a.num_ps:])
    ngpu = a.gpus
    if ngpu < 0:
        try:
            import torch
            ngpu = torch.cuda.device_count()
        except Exception:
            ngpu = 0
    procs = []
    n = 0
    for job, cnt in (("ps", a.num_ps), ("worker", a.num_workers)):
        for i in range(cnt):
            env = dict(os.environ)
            if ngpu > 0:
                env["DTF_GPU_INDEX"] = str(n % ngpu)
            cmd = [sys.executable, "-u", a.script, "--job_name=%s" % job, "--task_index=%d" % i,
                   "--ps_hosts=%s" % ps_hosts, "--worker_hosts=%s" % wk_hosts] + extra
            procs.append((job, i, subprocess.Popen(cmd, env=env)))
            n += 1
    rc = 0
    deadline = time.time() + a.timeout
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(143))      # killed launcher still reaps its tasks (finally below)
    try:
        for job, i, p in procs:
            if job != "worker" or (a.wait == "first" and i != 0):
                continue
            left = max(1.0, deadline - time.time())
            try:
                r = p.wait(timeout=left)
            except subprocess.TimeoutExpired:
                r = 124
            rc = rc or r
    finally:
        for job, i, p in procs:
            if p.poll() is None:
                p.terminate()
        for job, i, p in procs:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()
    sys.exit(rc)


if __name__ == "__main__":
    main()
