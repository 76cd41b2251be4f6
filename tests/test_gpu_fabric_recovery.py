"""Failure handling on the fabric tier, end to end (reference example_between_graph.py:99: "MonitoredTrainingSession ...
handles AbortedError in case of preempted PS"; here the peer that goes away is a sync replica -- the harder case, because
every other replica is blocked on the GPU waiting for a token that will never come).

Three processes share GPU 0 (ps, worker 0 = chief, worker 1) and train the MNIST MLP with SyncReplicasOptimizer on the
fused path.  Worker 1 is SIGKILLed mid-training and started again.  Expected: worker 0's device-side token wait times out
(DTF_FABRIC_STEP_TIMEOUT), the strategy raises AbortedError, MonitoredTrainingSession recovers, the fabric re-forms as
generation 1 (new rendezvous, new buffers; the variables keep their values on the ps) once the restarted worker is back,
and both finish at the stop step."""
import os
import re
import signal
import subprocess
import sys
import time

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _wait_for(path, pattern, timeout):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if os.path.exists(path):
            m = re.findall(pattern, open(path).read())
            if m:
                return m
        time.sleep(0.25)
    raise AssertionError("no %r in %s after %d s:\n%s" % (pattern, path, timeout, open(path).read()[-3000:] if os.path.exists(path) else ""))


def test_sync_replica_killed_and_restarted_job_recovers(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    base = 22700
    hosts = ["--ps_hosts=127.0.0.1:%d" % base, "--worker_hosts=127.0.0.1:%d,127.0.0.1:%d" % (base + 1, base + 2)]
    env = dict(os.environ, DTF_GPU_INDEX="0", DTF_FABRIC="1", DTF_HDFS_ROOT=str(tmp_path / "hdfs"), DTF_FABRIC_PORT_OFFSET="1600",
               DTF_FABRIC_STEP_TIMEOUT="3", DTF_FABRIC_TIMEOUT="150", DTF_FABRIC_LIVENESS_SECS="0.5")
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    steps = 6000

    def start(job, idx, log):
        cmd = [sys.executable, "-u", os.path.join(ROOT, "examples", "distributed_mnist.py"), "--train_steps=%d" % steps,
               "--train_dir=%s" % (tmp_path / "ckpt"), "--validate_every=100000", "--job_name=%s" % job, "--task_index=%d" % idx,
               "--issync=True"] + hosts
        f = open(tmp_path / log, "w")
        return subprocess.Popen(cmd, env=env, cwd=str(tmp_path), stdout=f, stderr=subprocess.STDOUT), f
    procs = {}
    try:
        procs["ps"] = start("ps", 0, "ps.log")
        procs["w0"] = start("worker", 0, "w0.log")
        procs["w1"] = start("worker", 1, "w1.log")
        _wait_for(tmp_path / "w1.log", r"training step:1500 ", 300)
        procs["w1"][0].send_signal(signal.SIGKILL)
        procs["w1"][0].wait(30)
        killed_at = int(re.findall(r"global step:(\d+)", open(tmp_path / "w0.log").read())[-1])
        procs["w1b"] = start("worker", 1, "w1b.log")
        for n in ("w0", "w1b"):
            rc = procs[n][0].wait(timeout=400)
            assert rc == 0, "%s exited with %s\n" % (n, rc) + "".join(
                "== %s\n%s\n" % (l, open(tmp_path / l).read()[-2500:]) for l in ("ps.log", "w0.log", "w1b.log"))
    finally:
        for p, f in procs.values():
            if p.poll() is None:
                p.kill()
            f.close()
    w0 = open(tmp_path / "w0.log").read()
    w1b = open(tmp_path / "w1b.log").read()
    assert re.search(r"dtf.fabric: fabric generation 0 of job \w+ aborted on worker 0: .*timed out", w0), w0[-3000:]
    assert "AbortedError" in w0 or "retrying" in w0 or "recover" in w0.lower() or w0.count("stop hook armed") >= 2, w0[-3000:]
    assert "routed onto the NVLink fabric" in w1b and "Training elapsed time" in w0 and "Training elapsed time" in w1b
    gs = [int(v) for v in re.findall(r"global step:(\d+)", w0)]
    assert max(gs) >= steps - 1 and killed_at < steps - 500                    # training went on well past the failure ...
    after = [g for g in gs if g > killed_at]
    assert len(after) > 500                                                     # ... on worker 0, in the new generation
    losses = [float(v) for v in re.findall(r"\| loss: ([0-9.eE+-]+)", w0)]
    assert sum(losses[-50:]) / 50 < 0.5 * sum(losses[:50]) / 50                # and it kept what it had learnt
    assert re.search(r"ran \d+ steps on the fused MLP step", w1b)
