"""The second batch of TF-1.x surface next to the reference's own (framework/ops_more.py, train/extras.py): values against numpy /
torch, gradients where the op is differentiable, TF's update formulas for Adadelta and the moving average."""
import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as tf


@pytest.fixture(autouse=True)
def _fresh_graph():
    tf.reset_default_graph()
    yield
    tf.reset_default_graph()


def test_unstack_slice_argmin_einsum_tensordot_accumulate_n():
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    x = tf.constant(a)
    parts = tf.unstack(x, axis=1)
    sl = tf.slice(x, [0, 1, 1], [2, -1, 2])
    am = tf.argmin(tf.constant([[3.0, 1.0, 2.0], [0.5, 4.0, 0.1]]), axis=1)
    es = tf.einsum("bij,bjk->bik", x, tf.constant(a.transpose(0, 2, 1)))
    td = tf.tensordot(x, tf.constant(a), axes=[[1, 2], [1, 2]])
    acc = tf.accumulate_n([x, x, x])
    w = tf.Variable(tf.ones([3, 4]), name="w")
    g = tf.gradients(tf.reduce_sum(tf.unstack(x * w, axis=0)[1]), [w])[0]
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        p, s, m, e, t, c, gv = sess.run([parts, sl, am, es, td, acc, g])
    assert len(p) == 3 and np.array_equal(p[2], a[:, 2]) and np.array_equal(s, a[:, 1:, 1:3]) and m.tolist() == [1, 2]
    assert np.allclose(e, np.einsum("bij,bjk->bik", a, a.transpose(0, 2, 1))) and np.allclose(t, np.tensordot(a, a, axes=[[1, 2], [1, 2]]))
    assert np.array_equal(c, 3 * a) and np.array_equal(gv, a[1])
    assert [q.shape for q in parts] == [(2, 4)] * 3 and sl.shape == (2, 2, 2)
    with pytest.raises(ValueError):
        tf.unstack(tf.placeholder(tf.float32, [None, 3]))


def test_random_shuffle_is_a_permutation_that_changes_per_run_and_replays_under_a_seed():
    def orders(seed):
        tf.reset_default_graph()
        tf.set_random_seed(seed)
        sh = tf.random_shuffle(tf.constant(np.arange(40, dtype=np.float32).reshape(20, 2)))
        with tf.Session() as sess:
            return [sess.run(sh) for _ in range(3)]
    a, b = orders(5), orders(5)
    for r in a:
        assert sorted(r[:, 0].tolist()) == list(np.arange(0, 40, 2.0)) and np.array_equal(r[:, 1], r[:, 0] + 1)     # whole rows move
    assert not np.array_equal(a[0], a[1]) and all(np.array_equal(x, y) for x, y in zip(a, b))
    assert not np.array_equal(a[0], orders(6)[0])


def test_numeric_guards_and_assertions():
    x = tf.placeholder(tf.float32, [3])
    checked = tf.check_numerics(x * 2.0, "doubled input")
    flags = [tf.is_nan(x), tf.is_inf(x), tf.is_finite(x)]
    guard = tf.Assert(tf.reduce_all(x > 0.0), ["x must be positive:", x])
    with tf.control_dependencies([guard]):
        y = tf.identity(x) + 1.0
    eq = tf.assert_equal(tf.to_int32(x), tf.constant([1, 2, 3]), message="ints differ")
    with tf.Session() as sess:
        assert sess.run(checked, {x: [1, 2, 3]}).tolist() == [2.0, 4.0, 6.0]
        n, i, f = sess.run(flags, {x: [float("nan"), float("inf"), 1.0]})
        assert n.tolist() == [True, False, False] and i.tolist() == [False, True, False] and f.tolist() == [False, False, True]
        with pytest.raises(tf.errors.InvalidArgumentError, match="doubled input.*NaN"):
            sess.run(checked, {x: [1, float("nan"), 3]})
        with pytest.raises(tf.errors.InvalidArgumentError, match="doubled input.*Inf"):
            sess.run(checked, {x: [1, float("inf"), 3]})
        assert sess.run(y, {x: [1, 2, 3]}).tolist() == [2.0, 3.0, 4.0]
        with pytest.raises(tf.errors.InvalidArgumentError, match="assertion failed"):
            sess.run(y, {x: [1, -2, 3]})
        sess.run(eq, {x: [1, 2, 3]})
        with pytest.raises(tf.errors.InvalidArgumentError, match="ints differ"):
            sess.run(eq, {x: [1, 2, 4]})


def test_sparse_to_dense_and_the_v2_cross_entropy():
    onehot = tf.sparse_to_dense(tf.constant([[0, 2], [1, 0], [2, 1]]), [3, 3], 1.0, 0.0)
    vec = tf.sparse_to_dense(tf.constant([1, 3]), [5], tf.constant([7.0, 9.0]))
    logits = tf.Variable(np.array([[1.0, 2.0, 3.0], [1.0, 0.0, -1.0], [0.5, 0.5, 0.5]], np.float32))
    labels = tf.Variable(np.eye(3, dtype=np.float32)[[2, 0, 1]])
    ce = tf.nn.softmax_cross_entropy_with_logits_v2(labels=labels, logits=logits)
    gl, gy = tf.gradients(tf.reduce_sum(ce), [logits, labels])
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        o, v, c, g1, g2 = sess.run([onehot, vec, ce, gl, gy])
    assert np.array_equal(o, np.eye(3, dtype=np.float32)[[2, 0, 1]]) and v.tolist() == [0.0, 7.0, 0.0, 9.0, 0.0]
    lt = torch.tensor([[1.0, 2.0, 3.0], [1.0, 0.0, -1.0], [0.5, 0.5, 0.5]], requires_grad=True)
    yt = torch.eye(3)[[2, 0, 1]].clone().requires_grad_(True)
    want = -(yt * torch.log_softmax(lt, -1)).sum(-1)
    want.sum().backward()
    assert np.allclose(c, want.detach().numpy(), atol=1e-6) and np.allclose(g1, lt.grad.numpy(), atol=1e-6)
    assert np.allclose(g2, yt.grad.numpy(), atol=1e-6)                  # v2: the labels receive a gradient too


def test_adadelta_follows_tensorflows_update():
    w0 = np.array([1.0, -2.0, 3.0], np.float32)
    w = tf.Variable(w0, name="w")
    loss = tf.reduce_sum(tf.square(w))
    opt = tf.train.AdadeltaOptimizer(learning_rate=0.5, rho=0.9, epsilon=1e-6)
    step = opt.minimize(loss)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        for _ in range(5):
            sess.run(step)
        got = sess.run(w)
    var, acc, upd = w0.astype(np.float64), np.zeros(3), np.zeros(3)
    for _ in range(5):
        g = 2 * var
        acc = 0.9 * acc + 0.1 * g * g
        u = np.sqrt(upd + 1e-6) / np.sqrt(acc + 1e-6) * g
        upd = 0.9 * upd + 0.1 * u * u
        var = var - 0.5 * u
    assert np.allclose(got, var, rtol=1e-5) and sorted(opt.get_slot_names()) == ["accum", "accum_update"]


def test_exponential_moving_average_tracks_the_variable_and_restores_under_the_shadow_names(tmp_path):
    w = tf.Variable(np.array([1.0, 2.0], np.float32), name="w")
    gs = tf.train.get_or_create_global_step()
    bump = tf.group(tf.assign_add(w, [1.0, 1.0]), tf.assign_add(gs, 1))
    ema = tf.train.ExponentialMovingAverage(0.5)
    with tf.control_dependencies([bump]):
        train = ema.apply([w])
    shadow = ema.average(w)
    assert ema.average_name(w) == "w/ExponentialMovingAverage" == shadow.var_name and not shadow.trainable
    saver = tf.train.Saver()
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        assert sess.run(shadow).tolist() == [1.0, 2.0]                   # starts at the variable's initial value
        want = np.array([1.0, 2.0])
        for k in range(3):
            sess.run(train)
            want = want - 0.5 * (want - (np.array([1.0, 2.0]) + k + 1))
        assert np.allclose(sess.run(shadow), want) and sess.run(w).tolist() == [4.0, 5.0]
        assert tf.train.global_step(sess, gs) == 3
        path = saver.save(sess, str(tmp_path / "m"), global_step=gs)
    assert np.allclose(tf.train.load_variable(str(tmp_path), "w/ExponentialMovingAverage"), want)
    # evaluation graph: the variable is restored FROM its average
    tf.reset_default_graph()
    w2 = tf.Variable(np.zeros(2, np.float32), name="w")
    tf.train.get_or_create_global_step()
    ema2 = tf.train.ExponentialMovingAverage(0.5)
    mapping = ema2.variables_to_restore([w2])
    assert set(mapping) == {"w/ExponentialMovingAverage", "global_step"}
    with tf.Session() as sess:
        tf.train.Saver(mapping).restore(sess, path)
        assert np.allclose(sess.run(w2), want)
    # num_updates ramps the decay: min(decay, (1 + n) / (10 + n))
    tf.reset_default_graph()
    v = tf.Variable(10.0, name="v")
    e = tf.train.ExponentialMovingAverage(0.99, num_updates=tf.constant(0))
    up = e.apply([v])
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        sess.run(tf.assign(v, 20.0))
        sess.run(up)
        assert sess.run(e.average(v)) == pytest.approx(10.0 - (1 - 0.1) * (10.0 - 20.0))


def test_init_from_checkpoint_warm_starts_through_the_ordinary_init_op_and_write_graph(tmp_path):
    a = tf.Variable(np.array([1.0, 2.0, 3.0], np.float32), name="enc/a")
    b = tf.Variable(np.array([[4.0]], np.float32), name="enc/b")
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        tf.train.Saver().save(sess, str(tmp_path / "src"))
    tf.reset_default_graph()
    a2 = tf.Variable(tf.zeros([3]), name="model/a")
    b2 = tf.Variable(tf.zeros([1, 1]), name="model/b")
    c2 = tf.Variable(tf.ones([2]), name="head/c")
    tf.train.init_from_checkpoint(str(tmp_path), {"enc/": "model/"})
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        assert sess.run(a2).tolist() == [1.0, 2.0, 3.0] and sess.run(b2).tolist() == [[4.0]] and sess.run(c2).tolist() == [1.0, 1.0]
    with pytest.raises(ValueError, match="shape"):
        tf.train.init_from_checkpoint(str(tmp_path), {"enc/a": c2})
    p = tf.train.write_graph(tf.get_default_graph(), str(tmp_path / "g"), "graph.pbtxt")
    text = open(p).read()
    assert "name: 'model/a'" in text and "op: 'VariableV2'" in text


def test_gfile_follows_the_checkpoint_path_mapping(tmp_path, monkeypatch):
    """``tf.gfile`` resolves URL-style paths the way the Saver does (the reference's ``hdfs://`` checkpoint directories,
    ``distributed_mnist.py:127``): what ``MonitoredTrainingSession(checkpoint_dir=...)`` writes, ``tf.gfile`` sees."""
    monkeypatch.setenv("DTF_HDFS_ROOT", str(tmp_path / "hdfs"))
    d = "hdfs://namenode:9000/user/ckpt"
    assert not tf.gfile.Exists(d)
    tf.gfile.MakeDirs(d)
    assert tf.gfile.IsDirectory(d) and (tmp_path / "hdfs" / "user" / "ckpt").is_dir()
    w = tf.Variable([1.0, 2.0], name="w")
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        tf.train.Saver().save(sess, d + "/model", global_step=3)
    assert "checkpoint" in tf.gfile.ListDirectory(d) and tf.gfile.Glob(d + "/model-3.*")
    with tf.gfile.GFile(d + "/notes.txt", "w") as f:
        f.write("line1\nline2\n")
    with tf.gfile.Open(d + "/notes.txt") as f:
        assert f.readline() == "line1\n" and list(f) == ["line2\n"] and f.size() == 12
    tf.gfile.Copy(d + "/notes.txt", d + "/copy.txt")
    with pytest.raises(tf.errors.OpError):
        tf.gfile.Copy(d + "/notes.txt", d + "/copy.txt")
    tf.gfile.Rename(d + "/copy.txt", d + "/moved.txt")
    assert tf.gfile.Stat(d + "/moved.txt").length == 12 and not tf.gfile.Exists(d + "/copy.txt")
    tf.gfile.Remove(d + "/moved.txt")
    with pytest.raises(tf.errors.NotFoundError):
        tf.gfile.Remove(d + "/moved.txt")
    assert [r for r, _, fs in tf.gfile.Walk(d) if "notes.txt" in fs]
    tf.gfile.DeleteRecursively(d)
    assert not tf.gfile.Exists(d)


def test_small_helpers_logsumexp_truncatediv_check_all_numerics_and_namespaces():
    x = tf.Variable(np.array([[1000.0, 1000.5, -5.0], [0.1, 0.2, 0.3]], np.float32))
    lse = tf.reduce_logsumexp(x, axis=1)
    lse_all = tf.reduce_logsumexp(x, keepdims=True)
    g = tf.gradients(tf.reduce_sum(lse), [x])[0]
    td = tf.truncatediv(tf.constant([7, -7, 9]), tf.constant([2, 2, -4]))
    y = tf.placeholder(tf.float32, [2])
    z = tf.log(y) * 2.0
    guard = tf.add_check_numerics_ops()
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        a, b, gv, t = sess.run([lse, lse_all, g, td])
        xt = torch.tensor([[1000.0, 1000.5, -5.0], [0.1, 0.2, 0.3]], requires_grad=True)
        want = torch.logsumexp(xt, 1)
        want.sum().backward()
        assert np.allclose(a, want.detach().numpy(), rtol=1e-6) and np.isfinite(a).all() and b.shape == (1, 1)
        assert np.allclose(gv, xt.grad.numpy(), atol=2e-5) and t.tolist() == [3, -3, -2]      # (fp32 rounding: ours is the closer one)
        sess.run([z, guard], {y: [1.0, 2.0]})
        with pytest.raises(tf.errors.InvalidArgumentError):
            sess.run([z, guard], {y: [1.0, -2.0]})                     # log of a negative number: NaN somewhere in the graph
        assert abs(float(sess.run(tf.timestamp())) - __import__("time").time()) < 5.0
    assert tf.compat.as_bytes("ab") == b"ab" and tf.compat.as_str(b"ab") == "ab" and tf.VERSION == tf.__version__
    assert tf.is_tensor(x) and tf.is_tensor(lse) and not tf.is_tensor(np.zeros(2))
    assert tf.test.gpu_device_name() in ("", "/device:GPU:0") and tf.initializers.zeros is tf.zeros_initializer
