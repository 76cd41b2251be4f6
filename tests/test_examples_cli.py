"""Example programs run the way a user of the reference runs them: one ``python X.py --job_name ... --task_index ...``
process per task on localhost ports (the only no-cluster technique the reference has).
CPU tier -- every script mirrors one reference script (S1-S20)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
ENV = dict(os.environ, CUDA_VISIBLE_DEVICES="", DTF_FABRIC="0")


def _run(args, timeout=240, cwd=None):
    r = subprocess.run([sys.executable] + args, capture_output=True, text=True, timeout=timeout, env=ENV, cwd=cwd or ROOT)
    assert r.returncode == 0, "exit %d\nSTDOUT:\n%s\nSTDERR:\n%s" % (r.returncode, r.stdout[-3000:], r.stderr[-3000:])
    return r.stdout


def test_standalone_one_gpu_golden_values():
    out = _run([os.path.join(EX, "standalone.py"), "--one_gpu"])
    flat = out.replace("\n", " ")
    assert "2." in flat and "7." in flat and "14." in flat            # [[2,3],[6,7]] and [[5],[14]]


def test_standalone_towers_train_and_report_time():
    out = _run([os.path.join(EX, "standalone.py"), "--iters", "30", "--batch_size", "200", "--num_gpu", "2"])
    assert "2 towers:" in out and "for 30 iterations" in out and "final tower loss" in out
    assert "affine_last/w" in out                                      # variables allocated once, shared by both towers


def test_between_graph_linear_regression_cluster_async(tmp_path):
    """example_between_graph.py: 1 ps + 2 workers, async SGD on y = 2x + 10, checkpoints in ckpt_dir."""
    out = _run([os.path.join(EX, "launch_local.py"), os.path.join(EX, "example_between_graph.py"), "--num_ps", "1",
                "--num_workers", "2", "--gpus", "0", "--timeout", "200", "--", "--num_steps=400", "--steps_to_validate=100",
                "--ckpt_dir=%s" % (tmp_path / "ckpt"), "--save_checkpoint_secs=1"], timeout=300)
    assert "weight" in out.lower() or "step" in out.lower()
    assert (tmp_path / "ckpt" / "checkpoint").exists()


def test_in_graph_example_golden_result_and_timeline(tmp_path):
    """example_in_graph.py on 1 ps + 2 workers: scatter on the ps, one matmul per worker, gather -> [[9],[21],[33],[45]]."""
    out = _run([os.path.join(EX, "launch_local.py"), os.path.join(EX, "example_in_graph.py"), "--num_ps", "1",
                "--num_workers", "2", "--gpus", "0", "--timeout", "120", "--wait", "first", "--", "--out_dir=%s" % tmp_path], timeout=200)
    flat = out.replace("\n", " ")
    for v in ("9.", "21.", "33.", "45."):
        assert v in flat, out
    tl = tmp_path / "timeline_client.json"
    assert tl.exists()
    ev = json.load(open(tl))["traceEvents"]
    assert any(e.get("ph") == "X" for e in ev)
    assert (tmp_path / "logs").exists() and os.listdir(tmp_path / "logs")


def test_distributed_server_plus_pure_client(tmp_path):
    """example_distributed_server.py x3 (ps, worker 0, worker 1 only serve) + example_distributed_client.py: a pure
    client with no ClusterSpec connects to worker 0's master and gets the same golden result."""
    import socket
    import time
    socks, ports = [], []
    for _ in range(3):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        ports.append(s.getsockname()[1])
        socks.append(s)
    for s in socks:
        s.close()
    hosts = ["--ps_hosts=127.0.0.1:%d" % ports[0], "--worker_hosts=127.0.0.1:%d,127.0.0.1:%d" % (ports[1], ports[2])]
    servers = []
    try:
        for job, idx in (("ps", 0), ("worker", 0), ("worker", 1)):
            servers.append(subprocess.Popen([sys.executable, "-u", os.path.join(EX, "example_distributed_server.py"),
                                             "--job_name=%s" % job, "--task_index=%d" % idx] + hosts, env=ENV,
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
        for p in servers:                           # each prints one "serving ..." line when it is up
            line = p.stdout.readline()
            assert "serving" in line, line
        out = _run([os.path.join(EX, "example_distributed_client.py"), "--master=grpc://127.0.0.1:%d" % ports[1],
                    "--out_dir=%s" % tmp_path], timeout=120)
        flat = out.replace("\n", " ")
        for v in ("9.", "21.", "33.", "45."):
            assert v in flat, out
        assert (tmp_path / "timeline_client.json").exists()
    finally:
        for p in servers:
            p.terminate()
        for p in servers:
            try:
                p.wait(timeout=10)
            except subprocess.TimeoutExpired:
                p.kill()


def test_mnist_standalone_single_process_cpu(tmp_path):
    """BASELINE.json config 1: MNIST MLP, one process, CPU, world_size 1 -- trains and checkpoints."""
    out = _run([os.path.join(EX, "mnist_standalone.py"), "--train_steps=120", "--num_train=2000", "--train_dir=%s" % (tmp_path / "ck")])
    line = [l for l in out.splitlines() if l.startswith("standalone:")][-1]
    assert "120 steps" in line and "steps/s" in line
    assert float(line.rsplit(" ", 1)[1]) < 60.0                 # batch-sum loss of 100 examples starts near 230
    assert (tmp_path / "ck" / "checkpoint").exists()


def test_backup_worker_survives_a_killed_replica(tmp_path):
    """Fault injection: 1 ps + 3 sync workers with replicas_to_aggregate=2 (one backup replica);
    worker 2 is SIGKILLed mid-run.  The survivors keep training to the stop step and exit promptly: the dead
    client's pending token dequeue is cancelled on the ps (no token is lost) and shutdown does not wait for it."""
    import signal
    import socket
    import time
    socks, ports = [], []
    for _ in range(4):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        ports.append(s.getsockname()[1])
        socks.append(s)
    for s in socks:
        s.close()
    script = os.path.join(EX, "example_between_graph.py")
    hosts = ["--ps_hosts=127.0.0.1:%d" % ports[0], "--worker_hosts=" + ",".join("127.0.0.1:%d" % p for p in ports[1:])]
    args = hosts + ["--is_sync=True", "--replicas_to_aggregate=2", "--num_steps=1200", "--steps_to_validate=400",
                    "--ckpt_dir=%s" % (tmp_path / "ck")]

    def task(job, idx, out):
        return subprocess.Popen([sys.executable, "-u", script, "--job_name=%s" % job, "--task_index=%d" % idx] + args, env=ENV,
                                stdout=out, stderr=subprocess.STDOUT, text=True)
    logs = [open(tmp_path / ("w%d.log" % i), "w") for i in range(3)]
    ps = task("ps", 0, subprocess.DEVNULL)
    workers = [task("worker", i, logs[i]) for i in range(3)]
    try:
        deadline = time.time() + 120
        while time.time() < deadline:                      # wait until training is under way, then kill worker 2
            if "step: 400" in open(tmp_path / "w0.log").read():
                break
            time.sleep(0.2)
        workers[2].send_signal(signal.SIGKILL)
        t_kill = time.time()
        assert workers[0].wait(timeout=120) == 0 and workers[1].wait(timeout=60) == 0
        assert time.time() - t_kill < 60                   # no 30 s connect retries towards the dead task at shutdown
        out0 = open(tmp_path / "w0.log").read()
        last = [l for l in out0.splitlines() if "weight:" in l][-1]
        assert int(last.split("step:")[1].split(",")[0]) >= 800            # progress continued after the kill at ~400
        assert abs(float(last.split("weight:")[1].split(",")[0]) - 2.0) < 0.3
        final = [int(f.split("-")[1].split(".")[0]) for f in os.listdir(tmp_path / "ck") if f.startswith("model.ckpt-") and f.endswith(".index")]
        assert max(final) >= 1200                                          # the chief's final checkpoint is at the stop step
    finally:
        for p in workers + [ps]:
            if p.poll() is None:
                p.kill()
        for f in logs:
            f.close()


def test_job_restart_resumes_from_checkpoint(tmp_path):
    """Checkpoint / resume at job level: the same cluster program is run twice with the same
    checkpoint directory; the second run's chief restores variables + global_step and ``StopAtStepHook(num_steps)``
    counts from the restored step."""
    def run():
        return _run([os.path.join(EX, "launch_local.py"), os.path.join(EX, "example_between_graph.py"), "--num_ps", "1",
                     "--num_workers", "1", "--gpus", "0", "--timeout", "200", "--", "--num_steps=300", "--steps_to_validate=100",
                     "--ckpt_dir=%s" % (tmp_path / "ck")], timeout=300)

    def steps(out):
        return [int(l.split("step:")[1].split(",")[0]) for l in out.splitlines() if "weight:" in l]
    first, second = steps(run()), steps(run())
    assert max(first) <= 300 and min(second) >= 300 and max(second) >= 500, (first, second)     # 0..300, then 301..601
    saved = [int(f.split("-")[1].split(".")[0]) for f in os.listdir(tmp_path / "ck") if f.endswith(".index")]
    assert max(saved) >= 600


@pytest.mark.parametrize("sync", ["False", "True"])
def test_supervisor_style_mnist_job(tmp_path, sync):
    """examples/mnist_supervisor.py: the mnist_replica.py generation of ps programs (tf.train.Supervisor, chief queue
    runner + init tokens in sync mode) on 1 ps + 2 workers."""
    out = _run([os.path.join(EX, "launch_local.py"), os.path.join(EX, "mnist_supervisor.py"), "--num_ps", "1", "--num_workers", "2",
                "--gpus", "0", "--timeout", "200", "--", "--sync_replicas=%s" % sync, "--train_steps=150", "--num_train=2000",
                "--train_dir=%s" % (tmp_path / "sv")], timeout=300)
    assert out.count("Session initialization complete.") == 2
    import re                                                     # two unbuffered workers share one pipe: lines may interleave
    vals = [float(v) for v in re.findall(r"validation cross entropy = ([0-9][0-9.eE+-]*)", out)]
    assert len(vals) == 2 and max(vals) < 5000.0                # 5000 validation images, batch-sum loss: ~11500 untrained
    assert (tmp_path / "sv" / "checkpoint").exists()


def test_mnist_fed_from_tfrecord_shards(tmp_path):
    """examples/mnist_tfrecords.py: the training split converted to Example records in four TFRecord shards, read back through
    TFRecordDataset -> map(parse) -> shuffle / repeat / batch / prefetch -> get_next under a MonitoredTrainingSession."""
    out = _run([os.path.join(EX, "mnist_tfrecords.py"), "--data_dir", str(tmp_path / "rec"), "--train_steps", "120", "--num_train", "1200"])
    line = [l for l in out.splitlines() if l.startswith("tfrecords:")][0]
    assert "120 steps" in line and "(1200 records in 4 shards)" in line
    assert float(line.split("held-out accuracy ")[1].split()[0]) > 0.9
    assert len(list((tmp_path / "rec").glob("train-*.tfrecord"))) == 4 and (tmp_path / "rec" / "validation.tfrecord").exists()
