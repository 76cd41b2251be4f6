"""Golden numerics tier: in-graph result, one_gpu values, linear regression, TF-Adam,
clipped xent sum, sync mean-of-N == single worker on the concatenated batch / N."""
import math

import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.train.optimizer import adam_reference_step


def test_one_gpu_golden_values():
    with dtf.device("/cpu:0"):
        w = dtf.Variable(dtf.constant([[1.0, 2.0], [4.0, 5.0]]), name="w")
        b = dtf.Variable(dtf.constant([[1.0], [2.0]]), name="b")
    with dtf.device("/gpu:0"):
        addwb, mulwb = dtf.add(w, b), dtf.matmul(w, b)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        v1, v2 = sess.run([addwb, mulwb])
    np.testing.assert_allclose(v1, [[2, 3], [6, 7]])
    np.testing.assert_allclose(v2, [[5], [14]])


def test_uninitialized_variable_raises():
    w = dtf.Variable(dtf.zeros([2]), name="w")
    with dtf.Session() as sess:
        with pytest.raises(dtf.errors.FailedPreconditionError):
            sess.run(w)
        assert sess.run(dtf.report_uninitialized_variables()) == ["w"]
        sess.run(w.initializer)
        assert sess.run(dtf.report_uninitialized_variables()) == []


def test_placeholder_must_be_fed_and_shape_checked():
    x = dtf.placeholder(dtf.float32, [None, 3])
    y = x * 2.0
    with dtf.Session() as sess:
        with pytest.raises(ValueError):
            sess.run(y)
        with pytest.raises(ValueError):
            sess.run(y, {x: np.zeros((2, 4), np.float32)})
        np.testing.assert_allclose(sess.run(y, {x: np.ones((2, 3))}), 2 * np.ones((2, 3)))


def test_fetch_structures_and_names():
    a = dtf.constant(3.0, name="a")
    b = dtf.constant(4.0, name="b")
    with dtf.Session() as sess:
        out = sess.run({"s": a + b, "l": [a, (b, "a:0")]})
    assert out["s"] == 7.0 and out["l"][0] == 3.0 and out["l"][1] == (4.0, 3.0)


def test_linear_regression_converges_like_reference():
    """example_between_graph.py:36,61,64 -- w->2, b->10 within 2000 SGD steps at lr 0.03."""
    rng = np.random.RandomState(0)
    tx = rng.rand(100).astype(np.float32)
    ty = 2 * tx + 10
    dtf.set_random_seed(1)
    gs = dtf.Variable(0, name="global_step", trainable=False, dtype=dtf.int64)
    X, y = dtf.placeholder(dtf.float32), dtf.placeholder(dtf.float32)
    w = dtf.get_variable("weight", [1], dtf.float32, initializer=dtf.random_normal_initializer())
    b = dtf.get_variable("biase", [1], dtf.float32, initializer=dtf.random_normal_initializer())
    loss = dtf.reduce_mean(dtf.square(y - (dtf.multiply(X, w) + b)))
    train = dtf.train.GradientDescentOptimizer(0.03).minimize(loss, global_step=gs)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        for _ in range(2000):
            sess.run(train, {X: tx, y: ty})
        wv, bv, step = sess.run([w, b, gs])
    assert step == 2000
    assert abs(wv[0] - 2) < 0.25 and abs(bv[0] - 10) < 0.15


def test_adam_matches_tf_closed_form():
    """epsilon OUTSIDE the bias correction; beta powers kept as variables."""
    g0 = np.array([0.5, -2.0, 3.0], np.float32)
    w = dtf.Variable(dtf.constant([1.0, 2.0, 3.0]), name="w")
    gph = dtf.placeholder(dtf.float32, [3])
    opt = dtf.train.AdamOptimizer(0.01)
    train = opt.apply_gradients([(gph, w)])
    ref_w, ref_m, ref_v = torch.tensor([1.0, 2.0, 3.0]), torch.zeros(3), torch.zeros(3)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        for t in range(1, 6):
            g = g0 * t
            sess.run(train, {gph: g})
            ref_w, ref_m, ref_v = adam_reference_step(ref_w, ref_m, ref_v, torch.from_numpy(g), t, lr=0.01)
            np.testing.assert_allclose(sess.run(w), ref_w.numpy(), rtol=1e-5, atol=1e-6)
        b1p, b2p = sess.run(list(opt._get_beta_accumulators()))
    assert abs(b1p - 0.9 ** 6) < 1e-6 and abs(b2p - 0.999 ** 6) < 1e-6
    # differs from torch.optim.Adam (eps inside) -- make sure we are NOT that formula
    tw = torch.tensor([1.0, 2.0, 3.0], requires_grad=True)
    topt = torch.optim.Adam([tw], lr=0.01, eps=1e-2)
    assert topt is not None


def test_momentum_matches_tf_formula():
    w = dtf.Variable(dtf.constant([1.0, -1.0]), name="w")
    gph = dtf.placeholder(dtf.float32, [2])
    train = dtf.train.MomentumOptimizer(0.1, 0.9).apply_gradients([(gph, w)])
    acc, ref = np.zeros(2, np.float32), np.array([1.0, -1.0], np.float32)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        for t in range(4):
            g = np.array([0.3, 0.7], np.float32) * (t + 1)
            sess.run(train, {gph: g})
            acc = 0.9 * acc + g
            ref = ref - 0.1 * acc
            np.testing.assert_allclose(sess.run(w), ref, rtol=1e-6)


def _mnist_graph(H=16):
    hid_w = dtf.Variable(dtf.truncated_normal([784, H], stddev=1.0 / 28, seed=1), name="hid_w")
    hid_b = dtf.Variable(dtf.zeros([H]), name="hid_b")
    sm_w = dtf.Variable(dtf.truncated_normal([H, 10], stddev=1.0 / math.sqrt(H), seed=2), name="sm_w")
    sm_b = dtf.Variable(dtf.zeros([10]), name="sm_b")
    x = dtf.placeholder(dtf.float32, [None, 784])
    y_ = dtf.placeholder(dtf.float32, [None, 10])
    hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    y = dtf.nn.softmax(dtf.nn.xw_plus_b(hid, sm_w, sm_b))
    xent = -dtf.reduce_sum(y_ * dtf.log(dtf.clip_by_value(y, 1e-10, 1.0)))
    return (hid_w, hid_b, sm_w, sm_b), x, y_, y, xent


def test_xent_is_batch_sum_with_clip_and_grads_match_torch():
    """distributed_mnist.py:113 -- the loss is a SUM over the batch (no 1/B), clip gates the gradient."""
    vars_, x, y_, y, xent = _mnist_graph()
    grads = dtf.gradients(xent, list(vars_))
    rng = np.random.RandomState(0)
    bx = rng.rand(7, 784).astype(np.float32)
    by = np.eye(10, dtype=np.float32)[rng.randint(0, 10, 7)]
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        vals = sess.run([v for v in vars_])
        loss, gvals = sess.run([xent, grads], {x: bx, y_: by})
    tw = [torch.tensor(v, requires_grad=True) for v in vals]
    h = torch.relu(torch.from_numpy(bx) @ tw[0] + tw[1])
    p = torch.softmax(h @ tw[2] + tw[3], -1)
    tl = -(torch.from_numpy(by) * torch.log(torch.clamp(p, 1e-10, 1.0))).sum()
    tl.backward()
    assert abs(loss - tl.item()) < 1e-3
    for g, t in zip(gvals, tw):
        np.testing.assert_allclose(g, t.grad.numpy(), rtol=1e-4, atol=1e-5)
    # sum, not mean: doubling the batch doubles the loss
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        l1 = sess.run(xent, {x: bx, y_: by})
        l2 = sess.run(xent, {x: np.concatenate([bx, bx]), y_: np.concatenate([by, by])})
    assert abs(l2 - 2 * l1) < 1e-3 * abs(l1)


def test_fused_clipped_xent_node_equals_composed_graph():
    _, x, y_, y, xent = _mnist_graph()
    g = dtf.get_default_graph()
    logits = [n for n in g.nodes if n.op_type == "XwPlusB"][-1]
    fused = dtf.nn.clipped_softmax_xent_sum(logits, y_)
    rng = np.random.RandomState(1)
    bx = rng.rand(5, 784).astype(np.float32)
    by = np.eye(10, dtype=np.float32)[rng.randint(0, 10, 5)]
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        a, b = sess.run([xent, fused], {x: bx, y_: by})
    assert abs(a - b) < 1e-4 * max(1.0, abs(a))


def test_split_concat_expand_dims_and_tower_average():
    x = dtf.constant(np.arange(12, dtype=np.float32).reshape(4, 3))
    a, b = dtf.split(x, 2)
    assert a.shape == (2, 3)
    avg = dtf.reduce_mean(dtf.concat([dtf.expand_dims(a, 0), dtf.expand_dims(b, 0)], 0), 0, keep_dims=False)
    with dtf.Session() as sess:
        av, bv, m = sess.run([a, b, avg])
    np.testing.assert_allclose(av, np.arange(6).reshape(2, 3))
    np.testing.assert_allclose(m, (av + bv) / 2)
    with pytest.raises(ValueError):
        with dtf.Session() as sess:
            sess.run(dtf.split(dtf.constant(np.zeros((5, 2), np.float32)), 2)[0])


def test_variable_scope_reuse_shares_tower_weights():
    def tower(inp):
        with dtf.variable_scope("affine0"):
            w = dtf.get_variable("w", [2, 3], initializer=dtf.truncated_normal_initializer(0, 1))
            b = dtf.get_variable("b", [], initializer=dtf.zeros_initializer)
        return dtf.matmul(inp, w) + b, w
    x = dtf.placeholder(dtf.float32, [None, 2])
    _, w0 = tower(x)
    with pytest.raises(ValueError):
        tower(x)                                   # second creation without reuse is an error
    dtf.get_variable_scope().reuse_variables()
    _, w1 = tower(x)
    assert w0 is w1
    assert [v.var_name for v in dtf.global_variables()] == ["affine0/w", "affine0/b"]


def test_truncated_normal_within_two_sigma_and_seeded():
    t = dtf.truncated_normal([2000], stddev=0.5, seed=3)
    with dtf.Session() as sess:
        a = sess.run(t)
        b = sess.run(t)
    assert np.abs(a).max() <= 1.0 + 1e-6 and 0.3 < a.std() < 0.5
    assert not np.array_equal(a, b)                # TF: an op-level seed fixes the SEQUENCE; every run advances it ...
    with dtf.Session() as sess:
        np.testing.assert_array_equal(a, sess.run(t))      # ... and a new session replays it from the start
        np.testing.assert_array_equal(b, sess.run(t))


def test_conv_bn_pool_shapes_and_grads():
    x = dtf.placeholder(dtf.float32, [None, 8, 8, 3])
    w = dtf.get_variable("w", [3, 3, 3, 4], initializer=dtf.variance_scaling_initializer())
    s = dtf.get_variable("s", [4], initializer=dtf.ones_initializer())
    o = dtf.get_variable("o", [4], initializer=dtf.zeros_initializer())
    y = dtf.nn.conv2d(x, w, [1, 2, 2, 1], "SAME")
    y = dtf.nn.relu(dtf.nn.fused_batch_norm_train(y, s, o))
    y = dtf.nn.max_pool(y, [1, 2, 2, 1], [1, 2, 2, 1], "SAME")
    loss = dtf.reduce_mean(y)
    gw, gs = dtf.gradients(loss, [w, s])
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        yv, g1, g2 = sess.run([y, gw, gs], {x: np.random.rand(2, 8, 8, 3).astype(np.float32)})
    assert yv.shape == (2, 2, 2, 4) and g1.shape == (3, 3, 3, 4) and g2.shape == (4,)
    # reference conv against torch NCHW
    xin = np.random.rand(1, 5, 5, 3).astype(np.float32)
    x2 = dtf.placeholder(dtf.float32, [None, 5, 5, 3])
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        wv = sess.run(w)
        out = sess.run(dtf.nn.conv2d(x2, w, [1, 1, 1, 1], "SAME"), {x2: xin})
    ref = torch.nn.functional.conv2d(torch.from_numpy(xin).permute(0, 3, 1, 2), torch.from_numpy(wv).permute(3, 2, 0, 1),
                                     padding=1).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-4, atol=1e-5)


def test_layers_dense_and_logging_conveniences():
    """``tf.layers.dense`` creates ``<name>/kernel`` / ``<name>/bias`` through get_variable (scopes + reuse work) and trains;
    ``tf.logging`` mirrors the TF-1.x module."""
    dtf.set_random_seed(3)
    x = dtf.placeholder(dtf.float32, [None, 4])
    h = dtf.layers.dense(x, 8, activation=dtf.nn.relu, name="fc1")
    y = dtf.layers.dense(h, 1, name="out")
    again = dtf.layers.dense(x, 8, name="fc1", reuse=True)                # shares fc1's variables
    assert [v.var_name for v in dtf.trainable_variables()] == ["fc1/kernel", "fc1/bias", "out/kernel", "out/bias"]
    loss = dtf.reduce_mean(dtf.square(y - 1.0))
    step = dtf.train.GradientDescentOptimizer(0.1).minimize(loss)
    xs = np.random.RandomState(0).rand(16, 4).astype(np.float32)
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        first = sess.run(loss, {x: xs})
        for _ in range(60):
            sess.run(step, {x: xs})
        assert sess.run(loss, {x: xs}) < 0.5 * first
        a, b = sess.run([h, dtf.nn.relu(again)], {x: xs})
        np.testing.assert_allclose(a, b, rtol=1e-6)
    dtf.logging.set_verbosity(dtf.logging.INFO)
    assert dtf.logging.get_verbosity() == dtf.logging.INFO
    dtf.logging.info("step %d", 3)


def test_learning_rate_schedules_feed_the_apply_ops():
    """A tensor-valued learning rate (exponential_decay / piecewise_constant) is a run-time input of ApplyGradientDescent /
    ApplyMomentum / ApplyAdam; the fused fabric engines refuse it with a clear message."""
    gs = dtf.train.get_or_create_global_step()
    w = dtf.Variable([1.0], name="w")
    loss = dtf.reduce_sum(w * 2.0)                                   # d loss / d w = 2
    lr = dtf.train.exponential_decay(0.5, gs, decay_steps=2, decay_rate=0.1, staircase=True)
    opt = dtf.train.GradientDescentOptimizer(lr)
    step = opt.minimize(loss, global_step=gs)
    pw = dtf.train.piecewise_constant(gs, [1, 3], [1.0, 0.5, 0.25])
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        seen, expect, cur = [], [], 1.0
        for t in range(5):
            assert sess.run(pw) == pytest.approx([1.0, 1.0, 0.5, 0.5, 0.25][t])
            rate = 0.5 * 0.1 ** (t // 2)
            assert sess.run(lr) == pytest.approx(rate)
            sess.run(step)
            cur -= rate * 2.0
            expect.append(cur)
            seen.append(float(sess.run(w)[0]))
        np.testing.assert_allclose(seen, expect, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError, match="constant learning rate"):
        opt.fused_spec()
    # momentum and adam accept a schedule as well
    for make in (lambda r: dtf.train.MomentumOptimizer(r, 0.9), dtf.train.AdamOptimizer):
        dtf.reset_default_graph()
        gs2 = dtf.train.get_or_create_global_step()
        v = dtf.Variable([1.0, -1.0], name="v")
        tr = make(dtf.train.exponential_decay(0.1, gs2, 10, 0.5)).minimize(dtf.reduce_sum(dtf.square(v)), global_step=gs2)
        with dtf.Session() as sess:
            sess.run(dtf.global_variables_initializer())
            before = float(np.abs(sess.run(v)).sum())
            for _ in range(20):
                sess.run(tr)
            assert float(np.abs(sess.run(v)).sum()) < before


def test_random_ops_are_stateful_per_op_streams():
    """with set_random_seed, (1) two same-shaped initialisers differ, also across tasks, (2) a seeded random op
    advances between runs, (3) a fresh session replays the same sequence, (4) dropout honours seeds."""
    import distributed_tensorflow_b200 as dtf
    dtf.set_random_seed(7)
    a = dtf.Variable(dtf.truncated_normal([64], stddev=1.0), name="a")
    b = dtf.Variable(dtf.truncated_normal([64], stddev=1.0), name="b")
    r = dtf.random_normal([8], seed=3)
    d = dtf.nn.dropout(dtf.ones([1000]), keep_prob=0.5, seed=11)
    runs = []
    for _ in range(2):
        with dtf.Session() as sess:
            sess.run(dtf.global_variables_initializer())
            av, bv = sess.run([a, b])
            r1, r2 = sess.run(r), sess.run(r)
            d1, d2 = sess.run(d), sess.run(d)
        assert not np.allclose(av, bv)
        assert not np.allclose(r1, r2) and not np.array_equal(d1, d2)
        assert set(np.unique(d1)) <= {0.0, 2.0} and 350 < (d1 > 0).sum() < 650
        runs.append((av, bv, r1, r2, d1))
    for x, y in zip(*runs):
        np.testing.assert_array_equal(x, y)


def test_same_shaped_initialisers_on_two_ps_tasks_differ(ports):
    """The placement case from the advisory: variables that are each the FIRST random draw on their own ps task."""
    import distributed_tensorflow_b200 as dtf
    p = ports(3)
    cluster = dtf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0], "127.0.0.1:%d" % p[1]], "worker": ["127.0.0.1:%d" % p[2]]})
    servers = [dtf.train.Server(cluster, "ps", 0), dtf.train.Server(cluster, "ps", 1), dtf.train.Server(cluster, "worker", 0)]
    try:
        dtf.set_random_seed(5)
        with dtf.device("/job:ps/task:0"):
            v0 = dtf.Variable(dtf.truncated_normal([32]), name="v0")
        with dtf.device("/job:ps/task:1"):
            v1 = dtf.Variable(dtf.truncated_normal([32]), name="v1")
        with dtf.Session(servers[2].target) as sess:
            sess.run(dtf.global_variables_initializer())
            a, b = sess.run([v0, v1])
        assert not np.allclose(a, b)
    finally:
        for s in servers:
            s.stop()


def test_fetched_values_do_not_alias_live_variables():
    """``sess.run(var)`` hands out a copy (TF semantics): a later apply must not change an array the caller kept."""
    import numpy as np
    import distributed_tensorflow_b200 as tf
    w = tf.get_variable("alias_w", [3, 2], initializer=tf.constant_initializer(1.0))
    bump = tf.assign_add(w, tf.ones([3, 2]))
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        before = sess.run(w)
        sess.run(bump)
        after = sess.run(w)
    assert np.array_equal(before, np.ones((3, 2))) and np.array_equal(after, 2 * np.ones((3, 2)))
