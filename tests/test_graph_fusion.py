"""Plan-time fusion of the graph tier (framework/fusion.py): which patterns are rewritten, when a rewrite is refused because it
could be observed, and that a fused plan computes what the node-by-node plan computes -- on the torch path bit for bit, on the
kernel path (host emulation) to rounding.  Reference program: /root/reference/distributed_mnist.py:106-126."""
import shutil

import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as tf
from distributed_tensorflow_b200.framework import fusion
from distributed_tensorflow_b200.ops import cuda_lib
from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist


_EMU_DIR = []


def _emu_dir():
    if not _EMU_DIR:
        import tempfile
        _EMU_DIR.append(tempfile.mkdtemp(prefix="dtf_emu_fusion_"))
    return _EMU_DIR[0]


def _mnist_graph(hidden=32):
    tf.set_random_seed(3)
    x, y_ = tf.placeholder(tf.float32, [None, 784]), tf.placeholder(tf.float32, [None, 10])
    w1 = tf.Variable(tf.truncated_normal([784, hidden], stddev=1.0 / 28), name="hid_w")
    b1 = tf.Variable(tf.zeros([hidden]), name="hid_b")
    w2 = tf.Variable(tf.truncated_normal([hidden, 10], stddev=1.0 / np.sqrt(hidden)), name="sm_w")
    b2 = tf.Variable(tf.zeros([10]), name="sm_b")
    hid_lin = tf.nn.xw_plus_b(x, w1, b1)
    hid = tf.nn.relu(hid_lin)
    y = tf.nn.softmax(tf.nn.xw_plus_b(hid, w2, b2))
    loss = -tf.reduce_sum(y_ * tf.log(tf.clip_by_value(y, 1e-10, 1.0)))
    gs = tf.train.get_or_create_global_step()
    train = tf.train.GradientDescentOptimizer(0.001).minimize(loss, global_step=gs)
    return dict(x=x, y_=y_, hid_lin=hid_lin, hid=hid, y=y, loss=loss, train=train, vars=[w1, b1, w2, b2])


def test_patterns_are_planned_only_when_unobservable():
    g = _mnist_graph()
    with tf.Session() as sess:
        feeds = {g["x"].id if hasattr(g["x"], "id") else 0}

        def plan(fetches, fed=()):
            nodes = [sess._resolve(f) for f in fetches]
            return sess._plan(nodes, set(sess._resolve(f).id for f in fed)).fusions
        p = plan([g["train"], g["loss"]])
        assert p is not None and len(p["relu"]) == 1 and len(p["xent"]) == 1
        neg, sm, logits, labels, lo, interior = p["xent"][0]
        assert neg == g["loss"].id and sm == g["y"].id and labels == g["y_"].id and lo == 1e-10 and len(interior) == 5
        assert p["relu"][0] == [g["hid_lin"].id, g["hid"].id]
        assert plan([g["loss"], g["y"]])["xent"] == []                      # the softmax output is fetched: keep the chain
        assert plan([g["loss"], g["hid_lin"]])["relu"] == []                # the pre-activation is fetched: keep XwPlusB + Relu
        assert plan([g["loss"]], fed=[g["y"]]) is None or plan([g["loss"]], fed=[g["y"]])["xent"] == []     # y is fed
        only_pred = plan([g["y"]])
        assert only_pred is not None and only_pred["xent"] == [] and len(only_pred["relu"]) == 1
        # a second consumer of the hidden pre-activation
        extra = tf.reduce_sum(g["hid_lin"])
        assert plan([g["loss"], extra])["relu"] == []
        # gradients with respect to an interior value
        gy = tf.gradients(g["loss"], [g["y"]])[0]
        p2 = plan([gy])
        assert p2 is None or p2["xent"] == []
    # other clip bounds / axis reductions are not the pattern
    tf.reset_default_graph()
    x = tf.placeholder(tf.float32, [None, 10])
    y_ = tf.placeholder(tf.float32, [None, 10])
    with tf.Session() as sess:
        for loss in (-tf.reduce_sum(y_ * tf.log(tf.clip_by_value(tf.nn.softmax(x), 1e-2, 1.0))),
                     -tf.reduce_sum(y_ * tf.log(tf.clip_by_value(tf.nn.softmax(x), 1e-10, 0.9))),
                     -tf.reduce_sum(y_ * tf.log(tf.clip_by_value(tf.nn.softmax(x), 1e-10, 1.0)), axis=1)):
            assert sess._plan([sess._resolve(loss)], set()).fusions is None


def _train(steps=5, fetch_extra=False):
    tf.reset_default_graph()
    g = _mnist_graph()
    xs, ys = synthetic_mnist(100 * steps, seed=2)
    out = []
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        for i in range(steps):
            fetches = [g["train"], g["loss"]] + ([g["y"], g["hid_lin"]] if fetch_extra else [])
            out.append(float(sess.run(fetches, {g["x"]: xs[i * 100:(i + 1) * 100], g["y_"]: ys[i * 100:(i + 1) * 100]})[1]))
        ws = sess.run(g["vars"])
        pred = sess.run(g["y"], {g["x"]: xs[:50]})                            # forward-only plan: ReLU fusion without the loss
    return out, ws, pred


def test_fused_plan_equals_node_by_node_plan_on_the_torch_path(monkeypatch):
    fused = _train()
    monkeypatch.setattr(fusion, "ENABLED", False)
    plain = _train()
    monkeypatch.setattr(fusion, "ENABLED", True)
    observed = _train(fetch_extra=True)                                    # fetching interior values switches the rewrites off per plan
    for other in (plain, observed):
        assert fused[0] == other[0]
        for a, b in zip(fused[1], other[1]):
            assert np.array_equal(a, b)
        assert np.array_equal(fused[2], other[2])
    assert fused[0][-1] < fused[0][0]


def test_fused_plan_runs_the_fused_kernels_under_the_emulation(tmp_path_factory, monkeypatch):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    cuda_lib.enable_emulation(str(tmp_path_factory.mktemp("emu_lib")))
    try:
        n0 = cuda_lib.launch_count()
        fused = _train(steps=3)
        n_fused = cuda_lib.launch_count() - n0
        monkeypatch.setattr(fusion, "ENABLED", False)
        n0 = cuda_lib.launch_count()
        plain = _train(steps=3)
        n_plain = cuda_lib.launch_count() - n0
    finally:
        cuda_lib.disable_emulation()
    assert n_fused <= n_plain - 3 * 5                                      # our element-wise launches the rewrites remove, per training step
    np.testing.assert_allclose(fused[0], plain[0], rtol=2e-4)
    for a, b in zip(fused[1], plain[1]):
        np.testing.assert_allclose(a, b, rtol=1e-3, atol=2e-5)
    np.testing.assert_allclose(fused[2], plain[2], rtol=1e-3, atol=1e-5)


def test_fusions_travel_to_remote_tasks(ports):
    """1 ps + 1 worker in this process: the worker's segment receives the planned rewrites through the run options."""
    p = ports(2)
    cluster = tf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0]], "worker": ["127.0.0.1:%d" % p[1]]})
    ps = tf.train.Server(cluster, job_name="ps", task_index=0)
    wk = tf.train.Server(cluster, job_name="worker", task_index=0)
    try:
        with tf.device(tf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
            g = _mnist_graph()
        xs, ys = synthetic_mnist(300, seed=2)
        seen = []
        orig = fusion.try_execute

        def spy(node, ctx, values, st, dev, want_grad):
            handled, out = orig(node, ctx, values, st, dev, want_grad)
            if handled:
                seen.append((ctx.task, node.op_type))
            return handled, out
        fusion.try_execute = spy
        try:
            with tf.Session(wk.target) as sess:
                sess.run(tf.global_variables_initializer())
                l0 = float(sess.run([g["train"], g["loss"]], {g["x"]: xs[:100], g["y_"]: ys[:100]})[1])
                l1 = float(sess.run([g["train"], g["loss"]], {g["x"]: xs[:100], g["y_"]: ys[:100]})[1])
        finally:
            fusion.try_execute = orig
        assert l1 < l0
        assert (("worker", 0), "Neg") in seen and (("worker", 0), "XwPlusB") in seen and (("worker", 0), "Softmax") in seen
    finally:
        wk.stop()
        ps.stop()


def _tower_graph():
    """The tower of /root/reference/standalone.py:46-63: matmul + scalar bias + relu per layer, mean squared error."""
    tf.set_random_seed(4)
    x, t = tf.placeholder(tf.float32, [None, 2]), tf.placeholder(tf.float32, [None, 1])
    h, ws = x, []
    for i, d in enumerate((16, 8)):
        w = tf.get_variable("affine%d/w" % i, [h.get_shape()[1], d], initializer=tf.truncated_normal_initializer(0, 1))
        b = tf.get_variable("affine%d/b" % i, [], initializer=tf.zeros_initializer)
        h = tf.nn.relu(tf.matmul(h, w) + b)
        ws += [w, b]
    w = tf.get_variable("affine_last/w", [h.get_shape()[1], 1], initializer=tf.constant_initializer(value=1))
    b = tf.get_variable("affine_last/b", [1], initializer=tf.zeros_initializer)          # a vector bias on the last layer
    y = tf.matmul(h, w) + b
    loss = tf.reduce_mean(tf.square(y - t))
    train = tf.train.GradientDescentOptimizer(0.01).minimize(loss)
    return dict(x=x, t=t, y=y, loss=loss, train=train, vars=ws + [w, b])


def _train_tower(steps=8):
    tf.reset_default_graph()
    g = _tower_graph()
    rng = np.random.RandomState(0)
    out = []
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        plan = sess._plan([sess._resolve(g["train"]), sess._resolve(g["loss"])], {g["x"].id, g["t"].id}).fusions
        for _ in range(steps):
            xs = rng.rand(64, 2).astype(np.float32)
            out.append(float(sess.run([g["train"], g["loss"]], {g["x"]: xs, g["t"]: xs.sum(1, keepdims=True)})[1]))
        ws = sess.run(g["vars"])
    return plan, out, ws


def test_matmul_bias_relu_towers_fuse_into_the_gemm_epilogue(monkeypatch, tmp_path_factory):
    plan, fused, fw = _train_tower()
    assert plan is not None and len(plan["affine"]) == 3
    assert [a[2] >= 0 for a in plan["affine"]] == [True, True, False]        # two layers end in a Relu, the last one does not
    monkeypatch.setattr(fusion, "ENABLED", False)
    plan0, plain, pw = _train_tower()
    assert plan0 is None
    np.testing.assert_allclose(fused, plain, rtol=1e-6)
    for a, b in zip(fw, pw):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
    assert fused[-1] < fused[0]
    if shutil.which("g++") is None:
        return
    monkeypatch.setattr(fusion, "ENABLED", True)
    cuda_lib.enable_emulation(str(tmp_path_factory.mktemp("emu_lib")))     # ... and on the kernel path
    try:
        _, emu, ew = _train_tower()
    finally:
        cuda_lib.disable_emulation()
    np.testing.assert_allclose(emu, plain, rtol=2e-3)
    for a, b in zip(ew, pw):
        np.testing.assert_allclose(a, b, rtol=5e-3, atol=5e-4)


def test_mean_squared_error_chain_is_one_fused_op(monkeypatch, tmp_path_factory):
    """example_between_graph.py:55-60: y = weight * x + biase; loss = reduce_mean(square(y_ - y))."""
    def run(steps=20):
        tf.reset_default_graph()
        tf.set_random_seed(1)
        x, y_ = tf.placeholder(tf.float32, [None]), tf.placeholder(tf.float32, [None])
        weight = tf.get_variable("weight", [1], tf.float32, initializer=tf.random_normal_initializer())
        biase = tf.get_variable("biase", [1], tf.float32, initializer=tf.random_normal_initializer())
        loss = tf.reduce_mean(tf.square(y_ - (tf.multiply(weight, x) + biase)))
        train = tf.train.GradientDescentOptimizer(0.1).minimize(loss)
        rng = np.random.RandomState(0)
        out = []
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            plan = sess._plan([sess._resolve(train), sess._resolve(loss)], {x.id, y_.id}).fusions
            for _ in range(steps):
                xs = rng.rand(16).astype(np.float32)
                out.append(sess.run([train, loss, weight, biase], {x: xs, y_: 2 * xs + 10})[1:])
        return plan, out
    plan, fused = run()
    assert plan is not None and len(plan["mse"]) == 1 and plan["affine"] == []
    monkeypatch.setattr(fusion, "ENABLED", False)
    _, plain = run()
    for (l1, w1, b1), (l2, w2, b2) in zip(fused, plain):
        np.testing.assert_allclose([l1, w1[0], b1[0]], [l2, w2[0], b2[0]], rtol=1e-6)
    if shutil.which("g++") is None:
        return
    cuda_lib.enable_emulation(str(tmp_path_factory.mktemp("emu_lib")))
    try:
        n0 = cuda_lib.launch_count()
        _, plain_emu = run()
        n_plain = cuda_lib.launch_count() - n0
        monkeypatch.setattr(fusion, "ENABLED", True)
        n0 = cuda_lib.launch_count()
        _, fused_emu = run()
        n_fused = cuda_lib.launch_count() - n0
    finally:
        cuda_lib.disable_emulation()
    assert n_fused <= n_plain - 20 * 2                                      # at least two launches fewer per step
    for (l1, w1, b1), (l2, w2, b2) in zip(fused_emu, plain):
        np.testing.assert_allclose([l1, w1[0], b1[0]], [l2, w2[0], b2[0]], rtol=2e-4, atol=1e-5)
    # a reduction over an axis, or a second consumer of the difference, is not the pattern
    tf.reset_default_graph()
    a, b = tf.placeholder(tf.float32, [None, 3]), tf.placeholder(tf.float32, [None, 3])
    with tf.Session() as sess:
        assert sess._plan([sess._resolve(tf.reduce_mean(tf.square(a - b), axis=1))], {a.id, b.id}).fusions is None
        d = a - b
        both = [sess._resolve(tf.reduce_mean(tf.square(d))), sess._resolve(tf.reduce_sum(d))]
        assert sess._plan(both, {a.id, b.id}).fusions is None


@pytest.mark.parametrize("seed", range(12))
def test_random_programs_fused_equals_unfused(seed, monkeypatch):
    """Randomly assembled programs from the op families the rewrites touch (affine layers in both spellings, scalar / vector biases,
    optional ReLU, either loss head) with random extra consumers and extra fetches of interior values: whatever the planner decides
    to fuse or to refuse, fetched values and gradients equal the node-by-node plan's."""
    rng = np.random.RandomState(100 + seed)

    def build_and_run():
        r = np.random.RandomState(100 + seed)                    # the same program both times
        tf.reset_default_graph()
        tf.set_random_seed(7)
        B, D = 12, int(r.randint(3, 9))
        x = tf.placeholder(tf.float32, [None, D])
        h, width, variables, extras = x, D, [], []
        for li in range(int(r.randint(1, 4))):
            out = int(r.randint(2, 7))
            w = tf.get_variable("w%d" % li, [width, out], initializer=tf.truncated_normal_initializer(0, 0.5))
            scalar_bias = bool(r.randint(0, 2))
            b = tf.get_variable("b%d" % li, [] if scalar_bias else [out], initializer=tf.constant_initializer(0.1))
            spelling = int(r.randint(0, 3))
            if spelling == 0 and not scalar_bias:
                pre = tf.nn.xw_plus_b(h, w, b)
            elif spelling == 1:
                pre = tf.matmul(h, w) + b
            else:
                pre = b + tf.matmul(h, w)                         # bias on the left
            if r.randint(0, 4) == 0:
                extras.append(tf.reduce_sum(pre))                 # a second consumer of the pre-activation
            h = tf.nn.relu(pre) if r.randint(0, 3) else pre
            if r.randint(0, 5) == 0:
                extras.append(h)                                  # the activation itself is fetched
            variables += [w, b]
            width = out
        if r.randint(0, 2):
            t = tf.placeholder(tf.float32, [None, width])
            diff_first = bool(r.randint(0, 2))
            loss = tf.reduce_mean(tf.square(h - t if diff_first else t - h))
        else:
            t = tf.placeholder(tf.float32, [None, width])
            y = tf.nn.softmax(h)
            loss = -tf.reduce_sum(t * tf.log(tf.clip_by_value(y, 1e-10, 1.0)))
            if r.randint(0, 3) == 0:
                extras.append(y)
        grads = tf.gradients(loss, variables)
        xs = r.rand(B, D).astype(np.float32)
        ts = r.rand(B, width).astype(np.float32)
        ts = ts / ts.sum(1, keepdims=True)
        with tf.Session() as sess:
            sess.run(tf.global_variables_initializer())
            fetches = [loss] + list(grads) + extras
            plan = sess._plan([sess._resolve(f) for f in fetches], {x.id, t.id}).fusions
            return plan, sess.run(fetches, {x: xs, t: ts})
    plan, fused = build_and_run()
    emu = None
    if seed % 3 == 0 and shutil.which("g++") is not None:        # every third program also on the kernel path (host emulation)
        cuda_lib.enable_emulation(_emu_dir())
        try:
            _, emu = build_and_run()
        finally:
            cuda_lib.disable_emulation()
    monkeypatch.setattr(fusion, "ENABLED", False)
    plan0, plain = build_and_run()
    assert plan0 is None and len(fused) == len(plain)
    for a, b in zip(fused, plain):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)
    if emu is not None:
        for a, b in zip(emu, plain):
            np.testing.assert_allclose(a, b, rtol=5e-3, atol=5e-4)
    test_random_programs_fused_equals_unfused.planned = getattr(test_random_programs_fused_equals_unfused, "planned", 0) + \
        (0 if plan is None else sum(len(v) for v in plan.values()))


def test_random_programs_exercised_the_planner():
    assert getattr(test_random_programs_fused_equals_unfused, "planned", 0) >= 8
