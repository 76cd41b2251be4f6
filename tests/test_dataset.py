"""utils/dataset.py: the minimal ``tf.data`` -- sources, transformations, iterators, end-of-sequence, a training loop without
feed_dict, one iterator per task in a ps/worker cluster."""
import numpy as np
import pytest

import distributed_tensorflow_b200 as tf
from distributed_tensorflow_b200.framework import errors


def _drain(get_next, sess, limit=10000):
    out = []
    try:
        for _ in range(limit):
            out.append(sess.run(get_next))
    except errors.OutOfRangeError:
        pass
    return out


def test_slices_map_batch_and_end_of_sequence():
    x = np.arange(10, dtype=np.float32).reshape(10, 1) * np.ones((1, 3), np.float32)
    y = np.arange(10, dtype=np.int64)
    ds = tf.data.Dataset.from_tensor_slices((x, y)).map(lambda a, b: (a * 2, b + 100)).batch(4)
    xb, yb = ds.make_one_shot_iterator().get_next()
    assert xb.shape == (None, 3) and yb.dtype == tf.int64
    with tf.Session() as sess:
        got = _drain([xb, yb], sess)
        with pytest.raises(errors.OutOfRangeError):
            sess.run(xb)
    assert [g[1].tolist() for g in got] == [[100, 101, 102, 103], [104, 105, 106, 107], [108, 109]]
    assert np.array_equal(got[0][0], x[:4] * 2)
    ds2 = tf.data.Dataset.from_tensor_slices((x, y)).batch(4, drop_remainder=True)
    with tf.Session() as sess:
        assert len(_drain(ds2.make_one_shot_iterator().get_next()[1], sess)) == 2


def test_shuffle_repeat_take_skip_filter_range_and_dicts():
    ds = tf.data.Dataset.range(20).shuffle(100, seed=3)
    nxt = ds.make_initializable_iterator()
    e = nxt.get_next()
    with tf.Session() as sess:
        with pytest.raises(errors.FailedPreconditionError):
            sess.run(e)
        sess.run(nxt.initializer)
        first = [int(v) for v in _drain(e, sess)]
        sess.run(nxt.initializer)
        second = [int(v) for v in _drain(e, sess)]
    assert sorted(first) == list(range(20)) and sorted(second) == list(range(20)) and first != list(range(20)) and first != second
    ds = tf.data.Dataset.range(5).repeat(2).skip(1).take(6).filter(lambda v: v % 2 == 0)
    with tf.Session() as sess:
        assert [int(v) for v in _drain(ds.make_one_shot_iterator().get_next(), sess)] == [2, 4, 0]
    ds = tf.data.Dataset.from_tensor_slices({"img": np.ones((6, 2), np.float32), "lab": np.arange(6)}).batch(3).prefetch(2)
    nx = ds.make_one_shot_iterator().get_next()
    with tf.Session() as sess:
        out = _drain(nx, sess)
    assert len(out) == 2 and out[1]["lab"].tolist() == [3, 4, 5] and out[0]["img"].shape == (3, 2)


def test_training_loop_without_feed_dict():
    from distributed_tensorflow_b200.utils.mnist_data import synthetic_mnist
    xs, ys = synthetic_mnist(2000, seed=3)
    ds = tf.data.Dataset.from_tensor_slices((xs, ys)).shuffle(2000, seed=1).repeat().batch(100)
    x, y_ = ds.make_one_shot_iterator().get_next()
    tf.set_random_seed(1)
    w = tf.get_variable("w", [784, 10], initializer=tf.zeros_initializer())
    b = tf.get_variable("b", [10], initializer=tf.zeros_initializer())
    loss = tf.losses.softmax_cross_entropy(y_, tf.nn.xw_plus_b(x, w, b))
    train = tf.train.GradientDescentOptimizer(0.5).minimize(loss)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        losses = [sess.run([train, loss])[1] for _ in range(60)]
    assert losses[-1] < 0.3 * losses[0]


def test_every_task_advances_its_own_iterator(ports):
    """Between-graph replication: the iterator op runs on the worker that built it, variables on the ps."""
    p = ports(2)
    cluster = tf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0]], "worker": ["127.0.0.1:%d" % p[1]]})
    ps = tf.train.Server(cluster, job_name="ps", task_index=0)
    wk = tf.train.Server(cluster, job_name="worker", task_index=0)
    try:
        with tf.device(tf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
            total = tf.get_variable("total", [], initializer=tf.zeros_initializer())
            v = tf.data.Dataset.range(1, 6).make_one_shot_iterator().get_next()
            add = tf.assign_add(total, tf.cast(v, tf.float32))
        assert "worker" in v.device and "ps" in total.device
        with tf.Session(wk.target) as sess:
            sess.run(tf.global_variables_initializer())
            for _ in range(5):
                sess.run(add)
            assert sess.run(total) == 15.0
            with pytest.raises(errors.OutOfRangeError):
                sess.run(add)
    finally:
        wk.stop()
        ps.stop()
