"""Backup workers on the fabric tier, multi-GPU (sorted after the other GPU files on purpose: it is the one hardware test of the
round that has not had its first hardware run -- the round's GPU budget was spent when it was written)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("workers,rta", [(2, 1), (4, 3)])
def test_backup_workers_on_the_fabric(workers, rta):
    """``SyncReplicasOptimizer(replicas_to_aggregate < total_num_replicas)`` on the device protocol (reference
    distributed_mnist.py:120-122 allows it): a straggler never overwrites a push the ps may still read (``consumed`` handshake),
    nothing deadlocks, one global step and ``rta`` gradients per aggregate.  Written after the round's GPU budget was spent: the
    same protocol runs under the host emulation with real concurrency (test_concurrent_backup_worker_with_a_straggler); this
    is its hardware twin, first run pending."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < workers:
        pytest.skip("needs %d GPUs" % workers)
    env = dict(os.environ, DTF_NVLS="auto", DTF_PS_ON_WORKERS="1", DTF_RTA=str(rta), MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(workers), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "mp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    rep = json.loads([l for l in r.stdout.splitlines() if l.startswith("MP_CHECK ")][-1][len("MP_CHECK "):])["backup_workers"]
    assert rep["ok"], rep
