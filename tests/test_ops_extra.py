"""framework/ops_extra.py + the Adagrad / RMSProp optimizers: ops a TF-1.x ps/worker program commonly uses next to the ones the
reference scripts call -- each against numpy / torch, gradients where they exist."""
import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as tf


def _run(fetches, feed=None):
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        return sess.run(fetches, feed or {})


def test_shape_and_selection_ops():
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    x = tf.constant(a)
    idx = tf.constant(np.array([2, 0], np.int64))
    cond = tf.greater_equal(x, 5.0)
    out = _run([tf.tile(x, [2, 1]), tf.gather(x, idx), tf.gather(x, idx, axis=1), tf.where(cond, x, -x), tf.where(cond),
                tf.size(x), tf.rank(x), tf.range(5), tf.range(2, 11, 3), tf.linspace(0.0, 1.0, 5), tf.cumsum(x, axis=1),
                tf.cumsum(x, axis=0, exclusive=True, reverse=True), tf.reverse(x, [1]), tf.pad(x, [[1, 0], [0, 2]]),
                tf.matrix_transpose(x), tf.nn.embedding_lookup(x, idx), tf.eye(3), tf.reduce_prod(x[1:2] if False else tf.constant(a[:, :2] + 1), axis=1)])
    assert np.array_equal(out[0], np.tile(a, (2, 1))) and np.array_equal(out[1], a[[2, 0]]) and np.array_equal(out[2], a[:, [2, 0]])
    assert np.array_equal(out[3], np.where(a >= 5, a, -a)) and np.array_equal(out[4], np.argwhere(a >= 5))
    assert out[5] == 12 and out[6] == 2 and out[7].tolist() == [0, 1, 2, 3, 4] and out[8].tolist() == [2, 5, 8]
    np.testing.assert_allclose(out[9], np.linspace(0, 1, 5))
    assert np.array_equal(out[10], np.cumsum(a, 1))
    assert np.array_equal(out[11], np.flip(np.cumsum(np.flip(a, 0), 0), 0) - a)
    assert np.array_equal(out[12], a[:, ::-1]) and np.array_equal(out[13], np.pad(a, [[1, 0], [0, 2]]))
    assert np.array_equal(out[14], a.T) and np.array_equal(out[15], a[[2, 0]]) and np.array_equal(out[16], np.eye(3))
    assert np.array_equal(out[17], np.prod(a[:, :2] + 1, axis=1))
    # a vector condition selects whole rows (TF's tf.where)
    rows = _run(tf.where(tf.constant([True, False, True]), x, tf.zeros([3, 4])))
    assert np.array_equal(rows, a * np.array([[1], [0], [1]], np.float32))


def test_elementwise_logic_and_rounding():
    v = np.array([-2.5, -0.4, 0.0, 0.5, 1.5, 7.2], np.float32)
    x = tf.constant(v)
    t = torch.from_numpy(v)
    out = _run([tf.floor(x), tf.ceil(x), tf.round(x), tf.sign(x), tf.nn.relu6(x), tf.nn.elu(x), tf.nn.leaky_relu(x, 0.1), tf.nn.softplus(x),
                tf.erf(x), tf.log1p(tf.abs(x)), tf.expm1(x), tf.logical_and(x > 0.0, x < 2.0), tf.logical_or(x < 0.0, x > 2.0),
                tf.logical_not(x > 0.0), tf.less_equal(x, 0.5), tf.not_equal(x, 0.0), tf.floordiv(x, 2.0), tf.mod(x, 2.0),
                tf.reduce_all(x > -3.0), tf.reduce_any(x > 7.0), tf.reduce_any(tf.constant([[True, False], [False, False]]), axis=1)])
    ref = [torch.floor(t), torch.ceil(t), torch.round(t), torch.sign(t), torch.clamp(t, 0, 6), torch.nn.functional.elu(t),
           torch.nn.functional.leaky_relu(t, 0.1), torch.nn.functional.softplus(t), torch.erf(t), torch.log1p(t.abs()), torch.expm1(t),
           (t > 0) & (t < 2), (t < 0) | (t > 2), ~(t > 0), t <= 0.5, t != 0, torch.floor(t / 2), torch.remainder(t, 2.0)]
    for got, want in zip(out, ref):
        np.testing.assert_allclose(got, want.numpy(), rtol=1e-6, atol=1e-6)
    assert out[18] and out[19] and out[20].tolist() == [True, False]


def test_norms_and_gradient_clipping_in_a_training_step():
    tf.set_random_seed(2)
    w = tf.get_variable("w", [4, 3], initializer=tf.truncated_normal_initializer(stddev=1.0))
    b = tf.get_variable("b", [3], initializer=tf.constant_initializer(0.5))
    x = tf.placeholder(tf.float32, [None, 4])
    loss = tf.reduce_sum(tf.square(tf.matmul(x, w) + b)) * 100.0
    opt = tf.train.GradientDescentOptimizer(0.1)
    gv = opt.compute_gradients(loss, [w, b])
    clipped, gn = tf.clip_by_global_norm([g for g, _ in gv], 1.0)
    train = opt.apply_gradients(list(zip(clipped, [w, b])))
    xs = np.random.RandomState(0).rand(8, 4).astype(np.float32)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        w0, b0 = sess.run([w, b])
        g0, g1, gnv, n2, n1, cn = sess.run([gv[0][0], gv[1][0], gn, tf.norm(w), tf.norm(w, ord=1), tf.clip_by_norm(w, 0.5)], {x: xs})
        sess.run(train, {x: xs})
        w1, b1 = sess.run([w, b])
    want_gn = np.sqrt((g0 ** 2).sum() + (g1 ** 2).sum())
    np.testing.assert_allclose(gnv, want_gn, rtol=1e-5)
    assert want_gn > 1.0
    np.testing.assert_allclose(w1, w0 - 0.1 * g0 / want_gn, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(b1, b0 - 0.1 * g1 / want_gn, rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(n2, np.sqrt((w0 ** 2).sum()), rtol=1e-5)
    np.testing.assert_allclose(n1, np.abs(w0).sum(), rtol=1e-5)
    np.testing.assert_allclose(np.sqrt((cn ** 2).sum()), 0.5, rtol=1e-5)


def test_nn_and_loss_helpers():
    rng = np.random.RandomState(1)
    z = rng.randn(6, 5).astype(np.float32)
    y = (rng.rand(6, 5) > 0.5).astype(np.float32)
    lab = rng.randint(0, 5, 6)
    zt = torch.from_numpy(z)
    logits = tf.constant(z)
    vals, idx = tf.nn.top_k(logits, 2)
    out = _run([tf.nn.sigmoid_cross_entropy_with_logits(labels=tf.constant(y), logits=logits), tf.nn.l2_normalize(logits, axis=1),
                tf.nn.in_top_k(logits, tf.constant(lab), 2), vals, idx,
                tf.losses.mean_squared_error(tf.constant(y), logits), tf.losses.softmax_cross_entropy(tf.constant(np.eye(5, dtype=np.float32)[lab]), logits),
                tf.losses.sparse_softmax_cross_entropy(tf.constant(lab), logits)])
    np.testing.assert_allclose(out[0], torch.nn.functional.binary_cross_entropy_with_logits(zt, torch.from_numpy(y), reduction="none").numpy(), rtol=1e-5)
    np.testing.assert_allclose(out[1], z / np.linalg.norm(z, axis=1, keepdims=True), rtol=1e-5)
    top2 = np.argsort(-z, axis=1)[:, :2]
    assert out[2].tolist() == [lab[i] in top2[i] for i in range(6)]
    np.testing.assert_allclose(out[3], -np.sort(-z, axis=1)[:, :2]) and None
    assert np.array_equal(out[4], top2)
    np.testing.assert_allclose(out[5], ((z - y) ** 2).mean(), rtol=1e-5)
    ce = torch.nn.functional.cross_entropy(zt, torch.from_numpy(lab)).item()
    np.testing.assert_allclose(out[6], ce, rtol=1e-5)
    np.testing.assert_allclose(out[7], ce, rtol=1e-5)
    # differentiable where it should be
    g = _run(tf.gradients(tf.reduce_sum(tf.nn.l2_normalize(logits, axis=1) * tf.constant(y)), [logits])[0])
    zt2 = zt.clone().requires_grad_()
    (torch.nn.functional.normalize(zt2, dim=1) * torch.from_numpy(y)).sum().backward()
    np.testing.assert_allclose(g, zt2.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_print_and_py_func(capfd):
    x = tf.constant(np.array([1.0, 2.0, 3.0, 4.0], np.float32))
    p = tf.Print(x, [x, tf.reduce_sum(x)], message="x and its sum: ", first_n=2, summarize=3)
    doubled = tf.py_func(lambda a: a * 2, [x], tf.float32)
    s, c = tf.py_func(lambda a: (a.sum(), np.int64(a.size)), [x], [tf.float32, tf.int64])
    with tf.Session() as sess:
        for _ in range(3):
            out = sess.run([p, doubled, s, c])
    assert out[0].tolist() == [1, 2, 3, 4] and out[1].tolist() == [2, 4, 6, 8] and out[2] == 10.0 and out[3] == 4
    err = capfd.readouterr().err
    assert err.count("x and its sum: [1 2 3...][10]") == 2                 # first_n


@pytest.mark.parametrize("which", ["adagrad", "rmsprop", "rmsprop_momentum_centered"])
def test_adagrad_and_rmsprop_follow_tensorflows_formulas(which):
    tf.set_random_seed(3)
    w = tf.get_variable("w", [5, 2], initializer=tf.truncated_normal_initializer(stddev=0.5))
    x = tf.placeholder(tf.float32, [None, 5])
    loss = tf.reduce_mean(tf.square(tf.matmul(x, w) - 1.0))
    gs = tf.train.get_or_create_global_step()
    if which == "adagrad":
        opt = tf.train.AdagradOptimizer(0.1, initial_accumulator_value=0.1)
    elif which == "rmsprop":
        opt = tf.train.RMSPropOptimizer(0.01, decay=0.9, epsilon=1e-10)
    else:
        opt = tf.train.RMSPropOptimizer(0.01, decay=0.8, momentum=0.5, epsilon=1e-6, centered=True)
    grad = tf.gradients(loss, [w])[0]
    train = opt.minimize(loss, global_step=gs)
    rng = np.random.RandomState(0)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        wv = sess.run(w).astype(np.float64)
        acc = np.full_like(wv, 0.1)
        ms, mom, mg = np.ones_like(wv), np.zeros_like(wv), np.zeros_like(wv)
        for _ in range(6):
            xs = rng.rand(7, 5).astype(np.float32)
            g = sess.run(grad, {x: xs}).astype(np.float64)
            sess.run(train, {x: xs})
            if which == "adagrad":
                acc += g * g
                wv -= 0.1 * g / np.sqrt(acc)
            elif which == "rmsprop":
                ms = 0.9 * ms + 0.1 * g * g
                mom = 0.0 * mom + 0.01 * g / np.sqrt(ms + 1e-10)
                wv -= mom
            else:
                ms = 0.8 * ms + 0.2 * g * g
                mg = 0.8 * mg + 0.2 * g
                mom = 0.5 * mom + 0.01 * g / np.sqrt(ms - mg * mg + 1e-6)
                wv -= mom
            np.testing.assert_allclose(sess.run(w), wv, rtol=2e-5, atol=1e-6)
        assert sess.run(gs) == 6
    assert sorted(opt.get_slot_names()) == (["accumulator"] if which == "adagrad" else
                                            (["momentum", "rms"] if which == "rmsprop" else ["mg", "momentum", "rms"]))


def test_streaming_metrics_accumulate_over_batches_and_reset():
    labels = tf.placeholder(tf.int64, [None])
    preds = tf.placeholder(tf.int64, [None])
    acc, acc_update = tf.metrics.accuracy(labels, preds)
    mean, mean_update = tf.metrics.mean(tf.cast(preds, tf.float32), name="mean_pred")
    assert len(tf.local_variables()) >= 4 and not any(v in tf.trainable_variables() for v in tf.local_variables())
    with tf.Session() as sess:
        sess.run(tf.local_variables_initializer())
        assert sess.run(acc) == 0.0                                      # nothing seen yet: 0, not NaN
        u1 = sess.run([acc_update, mean_update], {labels: [1, 2, 3, 4], preds: [1, 2, 0, 0]})
        u2 = sess.run([acc_update, mean_update], {labels: [5, 6], preds: [5, 6]})
        assert u1[0] == pytest.approx(0.5) and u2[0] == pytest.approx(4 / 6)
        assert sess.run(acc) == pytest.approx(4 / 6) and sess.run(mean) == pytest.approx((1 + 2 + 0 + 0 + 5 + 6) / 6)
        sess.run(tf.local_variables_initializer())
        assert sess.run(acc) == 0.0


def test_cond_runs_only_the_chosen_branch_and_is_differentiable(capfd):
    x = tf.placeholder(tf.float32, [None])
    flag = tf.placeholder(tf.bool, [])
    w = tf.get_variable("cw", [], initializer=tf.constant_initializer(3.0))
    calls = []

    def yes():
        return tf.reduce_sum(tf.Print(x, [x], message="true branch ran ") * w)

    def no():
        return tf.reduce_sum(tf.py_func(lambda a: (calls.append("f"), a)[1], [x], tf.float32)) - w * w
    y = tf.cond(flag, yes, no)
    gw = tf.gradients(y, [w])[0]
    pair = tf.cond(tf.reduce_sum(x) > 5.0, lambda: (x + 1.0, w), lambda: (x - 1.0, w * 2.0))
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        xs = np.array([1.0, 2.0, 4.0], np.float32)
        assert sess.run(y, {x: xs, flag: True}) == pytest.approx(21.0)
        assert calls == [] and "true branch ran" in capfd.readouterr().err          # the other branch did not execute
        assert sess.run(y, {x: xs, flag: False}) == pytest.approx(7.0 - 9.0) and calls == ["f"]
        assert "true branch ran" not in capfd.readouterr().err
        assert sess.run(gw, {x: xs, flag: True}) == pytest.approx(7.0) and sess.run(gw, {x: xs, flag: False}) == pytest.approx(-6.0)
        a, b = sess.run(pair, {x: xs})
        assert a.tolist() == [2.0, 3.0, 5.0] and b == 3.0
        a, b = sess.run(pair, {x: xs - 1.0})
        assert a.tolist() == [-1.0, 0.0, 2.0] and b == 6.0
    with pytest.raises(ValueError):
        tf.cond(flag, lambda: (x, x), lambda: x)


def test_learning_rate_schedules_follow_tensorflows_definitions():
    import math
    gs = tf.train.get_or_create_global_step()
    bump = tf.assign_add(gs, tf.constant(7, dtype=tf.int64))
    sch = {"inv": tf.train.inverse_time_decay(0.1, gs, 10, 0.5), "inv_s": tf.train.inverse_time_decay(0.1, gs, 10, 0.5, staircase=True),
           "nat": tf.train.natural_exp_decay(0.1, gs, 10, 0.5), "poly": tf.train.polynomial_decay(0.1, gs, 20, 0.01, power=2.0),
           "poly_c": tf.train.polynomial_decay(0.1, gs, 20, 0.01, cycle=True), "cos": tf.train.cosine_decay(0.1, gs, 20, alpha=0.1),
           "exp": tf.train.exponential_decay(0.1, gs, 10, 0.9)}
    # a schedule drives an optimizer like a constant would
    w = tf.get_variable("sw", [], initializer=tf.constant_initializer(1.0))
    step = tf.train.GradientDescentOptimizer(sch["cos"]).minimize(tf.square(w))
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        for s in (0, 7, 14, 21, 28):
            got = sess.run(sch)
            want = {"inv": 0.1 / (1 + 0.5 * s / 10), "inv_s": 0.1 / (1 + 0.5 * (s // 10)), "nat": 0.1 * math.exp(-0.5 * s / 10),
                    "poly": (0.1 - 0.01) * (1 - min(s, 20) / 20) ** 2 + 0.01,
                    "poly_c": (0.1 - 0.01) * (1 - s / (20 * max(1, math.ceil(s / 20)))) + 0.01,
                    "cos": 0.1 * (0.9 * 0.5 * (1 + math.cos(math.pi * min(s, 20) / 20)) + 0.1), "exp": 0.1 * 0.9 ** (s / 10)}
            for k in want:
                assert got[k] == pytest.approx(want[k], rel=1e-5), (k, s)
            if s == 0:
                w0 = sess.run(w)
                sess.run(step)
                assert sess.run(w) == pytest.approx(w0 - 0.1 * 2 * w0)
            sess.run(bump)


def test_layers_build_a_small_cnn_that_trains_and_batch_norm_tracks_moving_statistics():
    rng = np.random.RandomState(0)
    xs = rng.rand(16, 8, 8, 3).astype(np.float32)
    ys = np.eye(4, dtype=np.float32)[rng.randint(0, 4, 16)]
    tf.set_random_seed(5)

    def net(x, training):
        with tf.variable_scope("cnn", reuse=tf.AUTO_REUSE):
            h = tf.layers.conv2d(x, 6, 3, padding="same", activation=tf.nn.relu, name="c1")
            h = tf.layers.batch_normalization(h, training=training, momentum=0.9, name="bn1")
            h = tf.layers.max_pooling2d(h, 2, 2)
            h = tf.layers.dropout(h, rate=0.2, training=training, seed=1)
            h = tf.layers.flatten(h)
            return tf.layers.dense(h, 4, name="out")
    x, y_ = tf.placeholder(tf.float32, [None, 8, 8, 3]), tf.placeholder(tf.float32, [None, 4])
    logits = net(x, True)
    assert logits.get_shape()[-1] == 4
    loss = tf.losses.softmax_cross_entropy(y_, logits)
    updates = tf.get_collection(tf.GraphKeys.UPDATE_OPS)
    assert len(updates) == 2
    with tf.control_dependencies(updates):
        train = tf.train.AdamOptimizer(0.01).minimize(loss)
    eval_logits = net(x, False)                                    # same variables, moving statistics, no dropout
    names = sorted(v.var_name for v in tf.global_variables() if v.var_name.startswith("cnn/") and "Adam" not in v.var_name)
    assert names == ["cnn/bn1/beta", "cnn/bn1/gamma", "cnn/bn1/moving_mean", "cnn/bn1/moving_variance", "cnn/c1/bias", "cnn/c1/kernel",
                     "cnn/out/bias", "cnn/out/kernel"]
    assert sorted(v.var_name for v in tf.trainable_variables()) == [n for n in names if "moving" not in n]
    mm = [v for v in tf.global_variables() if v.var_name == "cnn/bn1/moving_mean"][0]
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        assert np.array_equal(sess.run(mm), np.zeros(6))
        losses = [sess.run([train, loss], {x: xs, y_: ys})[1] for _ in range(40)]
        moved = sess.run(mm)
        e1, e2 = sess.run(eval_logits, {x: xs}), sess.run(eval_logits, {x: xs})
    assert losses[-1] < 0.6 * losses[0] and np.abs(moved).max() > 0.01
    assert np.array_equal(e1, e2) and e1.shape == (16, 4)          # inference is deterministic (no dropout, fixed statistics)


def test_while_loop_iterates_captures_outer_values_and_is_differentiable():
    x = tf.placeholder(tf.float32, [])
    n = tf.placeholder(tf.int32, [])
    w = tf.get_variable("lw", [], initializer=tf.constant_initializer(1.5))
    # y = x * w^n by repeated multiplication; i counts the iterations
    i0, y0 = tf.constant(0), x
    i, y = tf.while_loop(lambda i, y: i < n, lambda i, y: (i + 1, y * w), [i0, y0])
    gw = tf.gradients(y, [w])[0]
    s = tf.while_loop(lambda v: v < 100.0, lambda v: v * 2.0, tf.constant(3.0))                   # single loop variable
    capped = tf.while_loop(lambda v: v < 1e9, lambda v: v + 1.0, tf.constant(0.0), maximum_iterations=5)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        iv, yv, g = sess.run([i, y, gw], {x: 2.0, n: 4})
        assert iv == 4 and yv == pytest.approx(2.0 * 1.5 ** 4) and g == pytest.approx(2.0 * 4 * 1.5 ** 3)
        iv, yv = sess.run([i, y], {x: 2.0, n: 0})
        assert iv == 0 and yv == 2.0                                   # the body never ran
        assert sess.run(s) == 192.0 and sess.run(capped) == 5.0
    with pytest.raises(ValueError):
        tf.while_loop(lambda a, b: a < 1, lambda a, b: a + 1, [tf.constant(0), tf.constant(0)])
