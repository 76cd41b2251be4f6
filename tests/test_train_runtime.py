"""Hooks, MonitoredTrainingSession, Saver, sync-replica state machine, distributed sessions."""
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.parallel.ps_state import ConditionalAccumulator, FIFOQueue

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------- accumulators / queues (A12 oracle)
def test_native_accumulator_keeps_dtype_of_non_fp32_gradients():
    """float64 / bf16 variables: the accumulated mean keeps the gradient's dtype (the native fp32 sum is bypassed), the stale
    drop and the accumulator's time step behave the same, a consumer already waiting is served."""
    from distributed_tensorflow_b200.utils import native_runtime
    if not native_runtime.available():
        pytest.skip("native runtime not built")
    acc = native_runtime.make_accumulator("acc64")
    acc.set_global_step(3)
    got = []
    t = threading.Thread(target=lambda: got.append(acc.take_grad(2, timeout=10.0)))
    t.start()
    time.sleep(0.1)
    big = torch.tensor([1.0 + 2.0 ** -40, 2.0], dtype=torch.float64)          # not representable in fp32
    assert not acc.apply_grad(big, 2)                                          # stamped before the carried time step: stale
    assert acc.apply_grad(big, 3) and acc.apply_grad(big, 3)
    t.join(10.0)
    assert got and got[0].dtype == torch.float64 and torch.equal(got[0], big)
    assert acc.global_step == 4 and acc.num_dropped == 1
    b = native_runtime.make_accumulator("acc16")
    assert b.apply_grad(torch.ones(4, dtype=torch.bfloat16), 0)
    assert b.take_grad(1).dtype == torch.bfloat16


@pytest.mark.parametrize("native", [False, True])
def test_accumulator_mean_stale_drop_and_backup_workers(native):
    if native:
        from distributed_tensorflow_b200.utils import native_runtime
        if not native_runtime.available():
            pytest.skip("native runtime not built")
        acc = native_runtime.make_accumulator("acc")
    else:
        acc = ConditionalAccumulator(name="acc")
    assert acc.apply_grad(torch.tensor([2.0, 4.0]), 0)
    assert acc.apply_grad(torch.tensor([4.0, 8.0]), 0)
    assert acc.num_accumulated() == 2
    mean = acc.take_grad(2)
    np.testing.assert_allclose(mean.numpy(), [3.0, 6.0])            # MEAN, not sum
    acc.set_global_step(1)
    assert not acc.apply_grad(torch.tensor([100.0, 100.0]), 0)      # stamped 0 < global step 1 -> stale, dropped
    assert acc.num_accumulated() == 0
    # backup workers: 3 replicas push, aggregate needs only 2 -> averages what has arrived
    for v in (1.0, 2.0, 6.0):
        assert acc.apply_grad(torch.tensor([v, v]), 1)
    np.testing.assert_allclose(acc.take_grad(2).numpy(), [3.0, 3.0])
    # take_grad blocks until enough gradients arrive
    got = []
    t = threading.Thread(target=lambda: got.append(acc.take_grad(1)))
    t.start()
    time.sleep(0.15)
    assert not got
    acc.set_global_step(2)
    acc.apply_grad(torch.tensor([5.0, 5.0]), 2)
    t.join(2)
    np.testing.assert_allclose(got[0].numpy(), [5.0, 5.0])
    # cancellation
    ev = threading.Event()
    ev.set()
    with pytest.raises(dtf.errors.CancelledError):
        acc.take_grad(1, cancel=ev)


@pytest.mark.parametrize("native", [False, True])
def test_token_queue_fifo_and_blocking(native):
    if native:
        from distributed_tensorflow_b200.utils import native_runtime
        if not native_runtime.available():
            pytest.skip("native runtime not built")
        q = native_runtime.make_queue("q")
    else:
        q = FIFOQueue(name="q")
    q.enqueue_many([1, 1, 2])
    assert q.size() == 3 and [q.dequeue(), q.dequeue(), q.dequeue()] == [1, 1, 2]
    got = []
    t = threading.Thread(target=lambda: got.append(q.dequeue()))
    t.start()
    time.sleep(0.1)
    assert not got
    q.enqueue(7)
    t.join(2)
    assert got == [7]
    with pytest.raises(dtf.errors.DeadlineExceededError):
        q.dequeue(timeout=0.1)
    q.close()
    with pytest.raises(dtf.errors.OutOfRangeError):
        q.dequeue()


# ---------------------------------------------------------------- hooks
class Recorder(dtf.train.SessionRunHook):
    def __init__(self, log, tag):
        self.log, self.tag = log, tag

    def begin(self): self.log.append((self.tag, "begin"))
    def after_create_session(self, s, c): self.log.append((self.tag, "after_create_session"))
    def before_run(self, ctx): self.log.append((self.tag, "before_run"))
    def after_run(self, ctx, vals): self.log.append((self.tag, "after_run"))
    def end(self, s): self.log.append((self.tag, "end"))


def _tiny_train():
    gs = dtf.train.get_or_create_global_step()
    w = dtf.Variable(dtf.constant([5.0]), name="w")
    loss = dtf.reduce_sum(dtf.square(w))
    return gs, w, loss, dtf.train.GradientDescentOptimizer(0.1).minimize(loss, global_step=gs)


def test_hook_ordering_and_stop_at_last_step():
    gs, w, loss, train = _tiny_train()
    log = []
    hooks = [Recorder(log, "a"), dtf.train.StopAtStepHook(last_step=3), Recorder(log, "b")]
    n = 0
    with dtf.train.MonitoredTrainingSession(hooks=hooks) as sess:
        while not sess.should_stop():
            sess.run(train)
            n += 1
    assert n == 3
    assert log[:4] == [("a", "begin"), ("b", "begin"), ("a", "after_create_session"), ("b", "after_create_session")]
    assert log[4:8] == [("a", "before_run"), ("b", "before_run"), ("a", "after_run"), ("b", "after_run")]
    assert log[-2:] == [("a", "end"), ("b", "end")]


def test_stop_at_step_num_steps_is_relative_and_args_validated(tmp_path):
    gs, w, loss, train = _tiny_train()
    d = str(tmp_path / "ck")
    with dtf.train.MonitoredTrainingSession(checkpoint_dir=d, hooks=[dtf.train.StopAtStepHook(num_steps=4)],
                                            save_checkpoint_secs=None, log_step_count_steps=None) as sess:
        while not sess.should_stop():
            sess.run(train)
        assert sess.run(gs) == 4
        dtf.train.Saver().save(sess, os.path.join(d, "model.ckpt"), global_step=gs)
    # resume: num_steps counts from the RESTORED step (SURVEY §5 checkpoint/resume)
    steps = []
    with dtf.train.MonitoredTrainingSession(checkpoint_dir=d, hooks=[dtf.train.StopAtStepHook(num_steps=2)],
                                            save_checkpoint_secs=None, log_step_count_steps=None) as sess:
        while not sess.should_stop():
            steps.append(int(sess.run([train, gs])[1]))
    assert steps == [5, 6]
    with pytest.raises(ValueError):
        dtf.train.StopAtStepHook()
    with pytest.raises(ValueError):
        dtf.train.StopAtStepHook(num_steps=1, last_step=2)


def test_subclassed_stop_hook_like_reference_runs():
    class MyStop(dtf.train.StopAtStepHook):
        def after_run(self, run_context, run_values):
            if run_values.results >= self._last_step:
                run_context.request_stop()
    gs, w, loss, train = _tiny_train()
    with dtf.train.MonitoredTrainingSession(hooks=[MyStop(last_step=2)]) as sess:
        n = 0
        while not sess.should_stop():
            sess.run([train, gs, loss])
            n += 1
    assert n == 2


# ---------------------------------------------------------------- saver
def test_checkpoint_roundtrip_partial_restore_and_max_to_keep(tmp_path):
    d = str(tmp_path / "ckpt")
    gs = dtf.train.get_or_create_global_step()
    a = dtf.Variable(dtf.constant([[1.0, 2.0], [3.0, 4.0]]), name="hid_w")
    b = dtf.Variable(dtf.constant([9.0]), name="extra/Adam")
    h = dtf.Variable(dtf.cast(dtf.constant([1.5, -2.5]), dtf.bfloat16), name="half")
    saver = dtf.train.Saver(max_to_keep=2)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        paths = []
        for step in (1, 2, 3):
            sess.run(dtf.assign(gs, dtf.constant(step, dtype=dtf.int64)))
            paths.append(saver.save(sess, os.path.join(d, "model.ckpt"), global_step=gs))
    st = dtf.train.get_checkpoint_state(d)
    assert st.model_checkpoint_path.endswith("model.ckpt-3")
    assert [os.path.basename(p) for p in st.all_model_checkpoint_paths] == ["model.ckpt-2", "model.ckpt-3"]
    assert not os.path.exists(paths[0] + ".index") and os.path.exists(paths[2] + ".index")
    assert dict(dtf.train.list_variables(d))["hid_w"] == [2, 2]
    assert os.path.exists(paths[2] + ".meta") and os.path.exists(paths[2] + ".data-00000-of-00001")
    # a different program restores ONLY hid_w by name (predict-style partial restore)
    dtf.reset_default_graph()
    a2 = dtf.Variable(dtf.zeros([2, 2]), name="hid_w")
    with dtf.Session() as sess:
        dtf.train.Saver().restore(sess, dtf.train.latest_checkpoint(d))
        np.testing.assert_allclose(sess.run(a2), [[1, 2], [3, 4]])
    # a name the checkpoint lacks -> NotFoundError; wrong shape -> InvalidArgumentError
    dtf.reset_default_graph()
    dtf.Variable(dtf.zeros([1]), name="missing")
    with dtf.Session() as sess, pytest.raises(dtf.errors.NotFoundError):
        dtf.train.Saver().restore(sess, dtf.train.latest_checkpoint(d))
    dtf.reset_default_graph()
    dtf.Variable(dtf.zeros([3, 3]), name="hid_w")
    with dtf.Session() as sess, pytest.raises(dtf.errors.InvalidArgumentError):
        dtf.train.Saver().restore(sess, dtf.train.latest_checkpoint(d))
    # bf16 survives the round trip bit-exactly
    r = dtf.train.NewCheckpointReader(dtf.train.latest_checkpoint(d))
    assert r.get_tensor("half").dtype == torch.bfloat16 and r.get_tensor("half").tolist() == [1.5, -2.5]
    assert dtf.train.get_checkpoint_state(str(tmp_path / "nope")) is None


def test_hdfs_url_maps_to_local_root(tmp_path, monkeypatch):
    monkeypatch.setenv("DTF_HDFS_ROOT", str(tmp_path))
    from distributed_tensorflow_b200.train.saver import resolve_path
    assert resolve_path("hdfs://phoenix-001.phoenix.com:8020/test/ckpt") == os.path.join(str(tmp_path), "test/ckpt")
    v = dtf.Variable(dtf.constant([1.0]), name="v")
    with dtf.Session() as sess:
        sess.run(v.initializer)
        dtf.train.Saver().save(sess, "hdfs://h:1/test/ckpt/model.ckpt", global_step=5)
    assert dtf.train.latest_checkpoint("hdfs://h:1/test/ckpt").endswith("model.ckpt-5")


def test_checkpoint_saver_hook_saves_on_create_steps_and_end(tmp_path):
    gs, w, loss, train = _tiny_train()
    d = str(tmp_path / "c")
    with dtf.train.MonitoredTrainingSession(checkpoint_dir=d, save_checkpoint_secs=None, save_checkpoint_steps=2,
                                            hooks=[dtf.train.StopAtStepHook(last_step=5)]) as sess:
        while not sess.should_stop():
            sess.run(train)
    names = sorted(f for f in os.listdir(d) if f.endswith(".index"))
    assert names == ["model.ckpt-0.index", "model.ckpt-2.index", "model.ckpt-4.index", "model.ckpt-5.index"]


# ---------------------------------------------------------------- timeline / summary
def test_timeline_and_graph_dump(tmp_path):
    a = dtf.constant([[1.0, 2.0]])
    with dtf.device("/cpu:0"):
        b = dtf.matmul(a, dtf.constant([[3.0], [4.0]]), name="mm")
    md = dtf.RunMetadata()
    with dtf.Session() as sess:
        sess.run(b, options=dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE), run_metadata=md)
        w = dtf.summary.FileWriter(str(tmp_path / "logs"), sess.graph)
        w.close()
    assert any(e["name"] == "mm" and e["op"] == "MatMul" for e in md.step_stats)
    tr = json.loads(dtf.Timeline(md.step_stats).generate_chrome_trace_format())
    assert any(e.get("ph") == "X" and e["name"] == "mm" for e in tr["traceEvents"])
    assert any(e.get("ph") == "M" for e in tr["traceEvents"])
    ev = dtf.summary.read_events(w.path)
    assert any("graph_def" in e and any(n["name"] == "mm" for n in e["graph_def"]["node"]) for e in ev)


# ---------------------------------------------------------------- distributed sessions, in-process tasks
def test_in_graph_replication_golden_result(cluster3, tmp_path):
    """example_in_graph.py -> [[9],[21],[33],[45]]; one pid per device in the trace."""
    cluster, servers = cluster3
    with dtf.device('/job:ps/task:0/cpu:0'):
        input_data = dtf.Variable([[1., 2., 3.], [4., 5., 6.], [7., 8., 9.], [10., 11., 12.]], name="input_data")
        b = dtf.Variable([[1.], [1.], [2.]], name="w")
    inputs = dtf.split(input_data, 2)
    outputs = []
    md = dtf.RunMetadata()
    with dtf.Session(servers[1].target) as sess:
        sess.run(dtf.global_variables_initializer())
        for i in range(2):
            with dtf.device("/job:worker/task:%d/gpu:0" % i):
                np.testing.assert_allclose(sess.run(inputs[i]), np.arange(1 + 6 * i, 7 + 6 * i).reshape(2, 3))
                outputs.append(dtf.matmul(inputs[i], b))            # graph grows between runs
        with dtf.device('/job:ps/task:0/cpu:0'):
            output = dtf.concat(outputs, axis=0)
        res = sess.run(output, options=dtf.RunOptions(trace_level=dtf.RunOptions.FULL_TRACE), run_metadata=md)
    np.testing.assert_allclose(res, [[9], [21], [33], [45]])
    tasks = {e["task"] for e in md.step_stats}
    assert tasks == {"/job:ps/task:0", "/job:worker/task:0", "/job:worker/task:1"}
    # variables live on the ps server and persist across client sessions
    assert servers[0].store.variable_names() == ["input_data", "w"]
    assert servers[1].store.variable_names() == []
    with dtf.Session(servers[2].target) as sess2:
        np.testing.assert_allclose(sess2.run(b), [[1.], [1.], [2.]])


def _between_graph_worker(cluster, task, is_sync, steps, results, num_workers=2, lr=0.05, backup=0):
    """One between-graph client in its own graph (thread = stand-in for a worker process)."""
    g = dtf.Graph()
    with g.as_default():
        server_target = "grpc://" + cluster.task_address("worker", task)
        with dtf.device(dtf.train.replica_device_setter(cluster=cluster,
                                                        worker_device="/job:worker/task:%d" % task)):
            gs = dtf.train.get_or_create_global_step()
            w = dtf.get_variable("weight", [1], initializer=dtf.constant_initializer(0.0))
            b = dtf.get_variable("biase", [1], initializer=dtf.constant_initializer(0.0))
            X, Y = dtf.placeholder(dtf.float32), dtf.placeholder(dtf.float32)
            loss = dtf.reduce_mean(dtf.square(Y - (X * w + b)))
            opt = dtf.train.GradientDescentOptimizer(lr)
            hooks = [dtf.train.StopAtStepHook(last_step=steps)]
            if is_sync:
                opt = dtf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=num_workers - backup,
                                                      total_num_replicas=num_workers)
                hooks.append(opt.make_session_run_hook(task == 0))
            train = opt.minimize(loss, global_step=gs)
        rng = np.random.RandomState(task)
        n = 0
        with dtf.train.MonitoredTrainingSession(master=server_target, is_chief=(task == 0), hooks=hooks) as sess:
            while not sess.should_stop():
                tx = rng.rand(32).astype(np.float32)
                _, step = sess.run([train, gs], {X: tx, Y: 2 * tx + 10})
                n += 1
            raw = sess.raw_session()
        results[task] = n


@pytest.mark.parametrize("is_sync", [False, True])
def test_between_graph_two_workers_share_ps_variables(cluster3, is_sync):
    cluster, servers = cluster3
    results = {}
    threads = [threading.Thread(target=_between_graph_worker, args=(cluster, t, is_sync, 60, results)) for t in (0, 1)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not any(t.is_alive() for t in threads)
    ps = servers[0].store
    assert int(ps.read("global_step")) >= 60
    assert set(ps.variable_names()) >= {"global_step", "weight", "biase"}
    assert "weight" not in servers[1].store.variable_names()         # parameters are on the ps only
    # both workers contributed; in sync mode each aggregate consumed one gradient from each
    assert results[0] > 0 and results[1] > 0
    if is_sync:
        # every aggregate consumes replicas_to_aggregate gradients and hands out as many tokens, so the two
        # workers together take ~2 steps per global step (a fast worker may contribute both gradients of an
        # aggregate while the other is still starting -- accumulators count gradients, not distinct workers)
        assert 2 * 60 - 6 <= results[0] + results[1] <= 2 * 60 + 6
        assert servers[1].store.variable_names() == ["sync_rep_local_step"]      # worker-local step counter


def test_sync_mean_equals_single_worker_on_concatenated_batch(cluster3):
    """Golden: mean of N per-worker gradients == gradient of the mean loss over the union batch."""
    cluster, servers = cluster3
    xs = [np.array([1.0, 2.0], np.float32), np.array([3.0, 5.0], np.float32)]
    g = dtf.Graph()
    with g.as_default():
        with dtf.device("/job:ps/task:0"):
            gs = dtf.train.get_or_create_global_step()
            w = dtf.Variable(dtf.constant([0.0]), name="w")
        grad_ph = [dtf.placeholder(dtf.float32, [1]) for _ in range(2)]
        opt = dtf.train.SyncReplicasOptimizer(dtf.train.GradientDescentOptimizer(1.0), 2, 2)
        # two "workers" share one client here: push both gradients, then run the chief aggregate once
        with dtf.device("/job:worker/task:0"):
            train0 = opt.apply_gradients([(grad_ph[0], w)], global_step=gs)
        with dtf.Session(servers[1].target) as sess:
            sess.run(dtf.global_variables_initializer())
            sess.run(opt.local_step_init_op)
            sess.run(opt.chief_init_op)
            sess.run(opt.get_init_tokens_op())
            push = [n for n in g.nodes if n.op_type == "AccumulatorApplyGrad"][0]
            sess.run(push, {grad_ph[0]: [2.0]})
            sess.run(push, {grad_ph[0]: [4.0]})
            sess.run(opt.sync_op)                 # take_grad(2) -> mean 3 -> w -= 1.0*3 ; step 1 ; 2 tokens
            assert sess.run(w)[0] == pytest.approx(-3.0)
            assert sess.run(gs) == 1


def test_recoverable_session_survives_ps_restart(ports, tmp_path):
    """A15: kill the ps mid-training; the chief restores from the last checkpoint and continues."""
    p = ports(2)
    spec = {"ps": ["127.0.0.1:%d" % p[0]], "worker": ["127.0.0.1:%d" % p[1]]}
    cluster = dtf.train.ClusterSpec(spec)
    ps = dtf.train.Server(cluster, "ps", 0)
    wk = dtf.train.Server(cluster, "worker", 0)
    d = str(tmp_path / "ck")
    try:
        with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
            gs, w, loss, train = _tiny_train()
        seen = []
        with dtf.train.MonitoredTrainingSession(master=wk.target, is_chief=True, checkpoint_dir=d,
                                                save_checkpoint_secs=None, save_checkpoint_steps=1,
                                                log_step_count_steps=None,
                                                hooks=[dtf.train.StopAtStepHook(last_step=8)]) as sess:
            while not sess.should_stop():
                _, step = sess.run([train, gs])      # the fetched global_step is the value read BEFORE this run's update
                seen.append(int(step))
                if len(seen) == 4:
                    ps.stop()                                   # fault injection: the ps dies ...
                    ps = dtf.train.Server(cluster, "ps", 0)     # ... and comes back empty
            assert sess.num_recoveries >= 1
        assert seen[:4] == [0, 1, 2, 3] and seen[-1] == 7
        assert seen[4] in (3, 4)              # resumed from the newest checkpoint (step 3 or 4), not from zero
        assert int(dtf.train.NewCheckpointReader(dtf.train.latest_checkpoint(d)).get_tensor("global_step")) == 8
    finally:
        ps.stop()
        wk.stop()


# ---------------------------------------------------------------- multi-process (real processes, gloo-free control plane)
@pytest.mark.parametrize("mode", ["async", "sync"])
def test_multiprocess_distributed_mnist_then_predict(tmp_path, mode):
    d = str(tmp_path / "ck")
    cmd = [sys.executable, os.path.join(ROOT, "examples/launch_local.py"), os.path.join(ROOT, "examples/distributed_mnist.py"),
           "--num_ps", "1", "--num_workers", "2", "--gpus", "0", "--timeout", "150", "--",
           "--train_steps=80", "--num_train=1500", "--log_every=20", "--train_dir=" + d, "--hidden_units=32"]
    if mode == "sync":
        cmd.append("--issync=True")
    else:
        cmd.append("--measure_staleness=True")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "global_step is" in r.stdout and "Training elapsed time" in r.stdout
    if mode == "async":
        assert "staleness mean" in r.stdout
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "examples/distributed_mnist_predict.py"),
                         "--checkpoint_dir=" + d, "--hidden_units=32"], capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    acc = float(r2.stdout.strip().splitlines()[-1].split()[-1])
    assert acc > 0.3        # chance is 0.1


# ---------------------------------------------------------------- tf.train.Supervisor (the mnist_replica.py generation)
@pytest.mark.parametrize("sync", [False, True])
def test_supervisor_driven_ps_training(cluster3, tmp_path, sync):
    """The protocol of TF r1.3's mnist_replica.py (the reference's ancestor, distributed_mnist.py:57): Supervisor chief
    initialises, the other worker waits; with SyncReplicasOptimizer the chief runs the init-tokens op and starts the
    optimizer's queue runner through ``sv.start_queue_runners``; the chief checkpoints into ``logdir``."""
    cluster, servers = cluster3
    results, errs = {}, []
    xs = np.random.RandomState(0).rand(64).astype(np.float32)
    ys = 3.0 * xs - 1.0

    def worker(task):
        try:
            g = dtf.Graph()
            with g.as_default():
                with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:%d" % task)):
                    gs = dtf.train.get_or_create_global_step()
                    w = dtf.get_variable("w", [1], initializer=dtf.zeros_initializer())
                    b = dtf.get_variable("b", [1], initializer=dtf.zeros_initializer())
                    x, y = dtf.placeholder(dtf.float32), dtf.placeholder(dtf.float32)
                    loss = dtf.reduce_mean(dtf.square(y - (x * w + b)))
                    opt = dtf.train.GradientDescentOptimizer(0.2)
                    if sync:
                        opt = dtf.train.SyncReplicasOptimizer(opt, replicas_to_aggregate=2, total_num_replicas=2)
                    train_op = opt.minimize(loss, global_step=gs)
                    init_op = dtf.global_variables_initializer()
                chief = task == 0
                kw = {}
                if sync:
                    kw = {"local_init_op": opt.local_step_init_op, "ready_for_local_init_op": opt.ready_for_local_init_op}
                    if chief:
                        kw["local_init_op"] = opt.chief_init_op
                sv = dtf.train.Supervisor(is_chief=chief, logdir=str(tmp_path / "sv"), init_op=init_op, recovery_wait_secs=0.2,
                                          global_step=gs, save_model_secs=1, **kw)
                sess = sv.prepare_or_wait_for_session(servers[1 + task].target)
                if sync and chief:
                    sess.run(opt.get_init_tokens_op())
                    sv.start_queue_runners(sess, [opt.get_chief_queue_runner()])
                step = 0
                # sync mode: the replicas may be a step apart (the init tokens let one run ahead), so only the chief
                # decides when training ends; its sv.stop() closes the token queue and the other worker's blocked
                # train_op ends with OutOfRangeError -- the clean end-of-training signal, as in TF
                while not sv.should_stop() and (step < 150 or (sync and not chief)):
                    try:
                        _, step = sess.run([train_op, gs], feed_dict={x: xs, y: ys})
                    except dtf.errors.OutOfRangeError:
                        assert sync and not chief
                        break
                results[task] = (int(step), sess.run([w, b]))
                sv.stop()
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
    ts = [threading.Thread(target=worker, args=(t,)) for t in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(120)
    assert not errs, errs
    assert not any(t.is_alive() for t in ts)
    step, (wv, bv) = results[0]
    assert step >= 150 and abs(float(wv[0]) - 3.0) < 0.3 and abs(float(bv[0]) + 1.0) < 0.2
    assert dtf.train.latest_checkpoint(str(tmp_path / "sv")) is not None


def test_learning_rate_schedule_across_tasks(cluster3):
    """The schedule is computed where global_step lives (ps 0) and consumed by apply ops on the variable's ps task."""
    cluster, servers = cluster3
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
        gs = dtf.train.get_or_create_global_step()
        w = dtf.Variable([4.0], name="w")
        lr = dtf.train.piecewise_constant(gs, [1], [0.5, 0.25])
        train = dtf.train.GradientDescentOptimizer(lr).minimize(dtf.reduce_sum(w * 1.0), global_step=gs)
    assert "/job:ps" in lr.device and "/job:ps" in w.device
    with dtf.Session(servers[1].target) as sess:
        sess.run(dtf.global_variables_initializer())
        vals = []
        for _ in range(4):
            sess.run(train)
            vals.append(float(sess.run(w)[0]))
    np.testing.assert_allclose(vals, [3.5, 3.0, 2.75, 2.5], rtol=1e-6)      # steps 0,1 use 0.5; later ones 0.25


def test_profiler_hook_traces_steps_inside_a_training_loop(cluster3, tmp_path):
    """ProfilerHook: a FULL_TRACE step every N global steps, chrome trace per traced step with one process per device
    (ps and worker tasks both appear)."""
    cluster, servers = cluster3
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
        gs = dtf.train.get_or_create_global_step()
        w = dtf.Variable([[1.0], [2.0]], name="w")
        x = dtf.placeholder(dtf.float32, [None, 2])
        loss = dtf.reduce_mean(dtf.square(dtf.matmul(x, w)))
        train = dtf.train.GradientDescentOptimizer(0.01).minimize(loss, global_step=gs)
    prof = dtf.train.ProfilerHook(save_steps=4, output_dir=str(tmp_path / "prof"))
    with dtf.train.MonitoredTrainingSession(master=servers[1].target, is_chief=True,
                                            hooks=[dtf.train.StopAtStepHook(last_step=10), prof]) as sess:
        while not sess.should_stop():
            sess.run(train, feed_dict={x: np.ones((4, 2), np.float32)})
    assert len(prof.files) >= 2 and all(os.path.exists(f) for f in prof.files)
    ev = json.load(open(prof.files[0]))["traceEvents"]
    names = {e["args"]["name"] for e in ev if e.get("ph") == "M"}
    assert any("/job:ps" in n for n in names) and any("/job:worker" in n for n in names)
    assert any(e.get("ph") == "X" and e["args"]["op"] == "ApplyGradientDescent" for e in ev)


def test_recoverable_session_in_sync_mode_survives_ps_restart(ports, tmp_path):
    """A15 + A12: with SyncReplicasOptimizer the ps also owns the accumulators and the token queue; after it is
    restarted empty, the recovered chief session re-creates them (chief_init_op: accumulator steps, a fresh token
    queue, initial tokens, a new queue-runner thread) and training continues from the last checkpoint."""
    p = ports(2)
    cluster = dtf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0]], "worker": ["127.0.0.1:%d" % p[1]]})
    ps = dtf.train.Server(cluster, "ps", 0)
    wk = dtf.train.Server(cluster, "worker", 0)
    try:
        with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
            gs = dtf.train.get_or_create_global_step()
            w = dtf.Variable(dtf.constant([5.0]), name="w")
            opt = dtf.train.SyncReplicasOptimizer(dtf.train.GradientDescentOptimizer(0.1), replicas_to_aggregate=1,
                                                  total_num_replicas=1)
            train = opt.minimize(dtf.reduce_sum(dtf.square(w)), global_step=gs)
        seen, killed = [], False
        with dtf.train.MonitoredTrainingSession(master=wk.target, is_chief=True, checkpoint_dir=str(tmp_path / "ck"),
                                                save_checkpoint_secs=None, save_checkpoint_steps=1, log_step_count_steps=None,
                                                hooks=[opt.make_session_run_hook(True), dtf.train.StopAtStepHook(last_step=9)]) as sess:
            while not sess.should_stop():
                _, step = sess.run([train, gs])
                seen.append(int(step))
                if step >= 4 and not killed:
                    killed = True
                    ps.stop()
                    ps = dtf.train.Server(cluster, "ps", 0)
            assert sess.num_recoveries >= 1
        assert killed and max(seen) >= 8 and seen[0] <= 1        # fetched step = value before the run's own update
        final = dtf.train.NewCheckpointReader(dtf.train.latest_checkpoint(str(tmp_path / "ck")))
        assert int(final.get_tensor("global_step")) >= 9
        assert abs(float(final.get_tensor("w")[0])) < 5.0 * 0.8 ** 8 * 1.5       # w shrinks by 0.8 per applied step
    finally:
        ps.stop()
        wk.stop()


def test_departing_sync_replica_leaves_farewell_tokens(cluster3):
    """End-of-training hazard of SyncReplicasOptimizer (present in TF): replicas run up to a step apart, so one can finish
    (StopAtStepHook) while another is inside a step whose aggregate needs the finished replica's gradient.  A non-chief
    replica that ends cleanly therefore enqueues ``total_num_replicas - 1`` tokens in its hook's ``end()``; a replica
    blocked on the token queue proceeds, sees the same global step and stops."""
    cluster, servers = cluster3
    g = dtf.Graph()
    with g.as_default():
        with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:1")):
            gs = dtf.train.get_or_create_global_step()
            w = dtf.get_variable("w", [1], initializer=dtf.zeros_initializer())
            opt = dtf.train.SyncReplicasOptimizer(dtf.train.GradientDescentOptimizer(0.1), replicas_to_aggregate=3, total_num_replicas=3)
            opt.minimize(dtf.reduce_sum(w * w), global_step=gs)
            size = g.create_node("QueueSize", [], {"queue_name": opt._sync_token_queue_name}, "q_size", device=gs.device)
        chief_hook, hook = opt.make_session_run_hook(True), opt.make_session_run_hook(False)
        chief_hook.begin()
        hook.begin()
        with dtf.Session(servers[2].target) as sess:
            sess.run(dtf.global_variables_initializer())
            sess.run(opt.chief_init_op)                       # creates the accumulators / token queue on the ps
            assert int(sess.run(size)) == 0
            hook.end(sess)                                    # a non-chief replica leaves cleanly
            assert int(sess.run(size)) == 2                   # one token for each of the two other replicas
            chief_hook._q_runner = opt.get_chief_queue_runner()
            chief_hook.end(sess)                              # the chief leaves: queue closed ...
            hook.end(sess)                                    # ... a later farewell is refused quietly


def test_native_cpu_apply_flushes_denormal_slots_and_restores_the_callers_float_mode():
    """Weights whose gradient is exactly zero every step (MNIST's border pixels) have Adam slots that decay into the denormal
    range, where x86 takes a micro-code assist per operand (measured: 40 us -> 1.2 ms per apply of the 784x100 matrix, config 1
    dropped from ~1000 to ~450 steps/s after ~800 steps).  The native apply runs with flush-to-zero / denormals-are-zero for
    its own duration, as TensorFlow's kernels do; the caller's MXCSR comes back."""
    import time
    import torch
    from distributed_tensorflow_b200.utils.native_runtime import cpu_optimizer_apply
    n = 78400
    var, g = torch.randn(n), torch.zeros(n)
    v0 = var.clone()

    def run(m0, v0_):
        best, m, v = 1e9, None, None
        for _ in range(7):
            m, v = torch.full((n,), m0), torch.full((n,), v0_)
            t = time.perf_counter()
            ok = cpu_optimizer_apply(2, var, m, v, g, 0.01, 0.0, False, 0.9, 0.999, 1e-8)
            best = min(best, time.perf_counter() - t)
            if not ok:
                pytest.skip("native runtime library not built")
        return best, m, v
    t_norm, _, _ = run(1e-3, 1e-6)
    t_den, m, v = run(1e-40, 1e-10)
    assert float(m.abs().max()) == 0.0 and float(v.max()) == pytest.approx(0.999e-10, rel=1e-4)      # denormal moments are zero now
    assert t_den < 6 * t_norm + 2e-4, (t_den, t_norm)            # (30x without the flush)
    assert float(torch.tensor([1e-40]) * 1.0) != 0.0             # the interpreter's own float mode is untouched
    assert torch.isfinite(var).all() and (var - v0).abs().max() < 1.0
