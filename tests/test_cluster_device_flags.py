"""Unit tier: ClusterSpec, device strings, replica_device_setter, flags."""
import pytest

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.framework.device import DeviceSpec
from distributed_tensorflow_b200.utils import flags as F


def test_cluster_spec_queries():
    c = dtf.train.ClusterSpec({"ps": ["h:1"], "worker": ["h:2", " h:3"]})   # note the space (reference default)
    assert c.jobs == ["ps", "worker"]
    assert c.num_tasks("worker") == 2
    assert c.task_address("worker", 1) == "h:3"
    assert c.as_dict() == {"ps": ["h:1"], "worker": ["h:2", "h:3"]}
    assert c.find_task("h:3") == ("worker", 1)
    assert [t[:2] for t in c.all_tasks()] == [("ps", 0), ("worker", 0), ("worker", 1)]
    assert c == dtf.train.ClusterSpec(c)
    with pytest.raises(ValueError):
        c.task_address("worker", 5)
    with pytest.raises(ValueError):
        c.num_tasks("chief")


def test_cluster_sparse_and_device_map():
    c = dtf.train.ClusterSpec({"ps": ["a:1", "a:2"], "worker": {0: "a:3", 2: "a:5"}})
    assert c.task_indices("worker") == [0, 2]
    assert c.as_dict()["worker"] == {0: "a:3", 2: "a:5"}
    dm = c.device_map(8)
    assert dm[("ps", 0)] == 0 and dm[("ps", 1)] == 1 and dm[("worker", 0)] == 2 and dm[("worker", 2)] == 3
    assert set(c.device_map(0).values()) == {-1}


def test_device_spec_parse_and_merge():
    d = DeviceSpec.from_string("/job:ps/task:1/cpu:0")
    assert (d.job, d.task, d.device_type, d.device_index) == ("ps", 1, "CPU", 0)
    assert DeviceSpec.from_string("/device:GPU:3").device_index == 3
    assert DeviceSpec.from_string("/gpu:0").to_string() == "/device:GPU:0"
    outer = DeviceSpec.from_string("/job:worker/task:3/cpu:0")
    inner = DeviceSpec.from_string("/gpu:1")
    m = outer.merge_from(inner)
    assert m.to_string() == "/job:worker/task:3/device:GPU:1"
    with pytest.raises(ValueError):
        DeviceSpec.from_string("/job:ps/bogus")


def test_device_scopes_nest_inner_wins():
    with dtf.device("/job:worker/task:1"):
        with dtf.device("/gpu:0"):
            a = dtf.constant(1.0)
        with dtf.device("/job:ps/task:0/cpu:0"):
            b = dtf.constant(1.0)
        with dtf.device(None):
            c = dtf.constant(1.0)
    assert a.device == "/job:worker/task:1/device:GPU:0"
    assert b.device == "/job:ps/task:0/device:CPU:0"
    assert c.device == ""


def test_replica_device_setter_round_robin_mnist_two_ps():
    """global_step->ps0, hid_w->ps1, hid_b->ps0, sm_w->ps1, sm_b->ps0; compute on the worker."""
    cluster = dtf.train.ClusterSpec({"ps": ["a:1", "a:2"], "worker": ["a:3"]})
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0/cpu:0")):
        gs = dtf.train.get_or_create_global_step()
        hid_w = dtf.Variable(dtf.truncated_normal([784, 100]), name="hid_w")
        hid_b = dtf.Variable(dtf.zeros([100]), name="hid_b")
        sm_w = dtf.Variable(dtf.truncated_normal([100, 10]), name="sm_w")
        sm_b = dtf.Variable(dtf.zeros([10]), name="sm_b")
        x = dtf.placeholder(dtf.float32, [None, 784])
        h = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    got = [DeviceSpec.from_string(v.device).task for v in (gs, hid_w, hid_b, sm_w, sm_b)]
    assert got == [0, 1, 0, 1, 0]
    assert all(DeviceSpec.from_string(v.device).job == "ps" for v in (gs, hid_w, hid_b, sm_w, sm_b))
    assert h.device == "/job:worker/task:0/device:CPU:0"
    # slots are colocated with their variable
    opt = dtf.train.AdamOptimizer(0.01)
    loss = dtf.reduce_sum(h)
    opt.minimize(loss, global_step=gs)
    assert opt.get_slot(hid_w, "m").device == hid_w.device
    assert opt.get_slot(sm_b, "v").device == sm_b.device


def test_replica_device_setter_without_ps_is_noop():
    assert dtf.train.replica_device_setter(ps_tasks=0) is None


def test_flags_lazy_parse_and_types(monkeypatch):
    monkeypatch.setattr("sys.argv", ["prog", "--job_name=ps", "--task_index", "3", "--issync", "--learning_rate=0.5",
                                     "--unknown=1", "positional"])
    F.DEFINE_string("job_name", "worker", "")
    F.DEFINE_integer("task_index", 0, "")
    F.DEFINE_bool("issync", False, "")
    F.DEFINE_float("learning_rate", 0.01, "")
    F.DEFINE_integer("train_steps", 5000, "")
    FLAGS = F.FLAGS
    assert FLAGS.job_name == "ps" and FLAGS.task_index == 3 and FLAGS.issync is True
    assert FLAGS.learning_rate == 0.5 and FLAGS.train_steps == 5000
    assert FLAGS.unparsed == ["--unknown=1", "positional"]
    with pytest.raises(AttributeError):
        FLAGS.nope
    FLAGS.train_steps = 7
    assert FLAGS.train_steps == 7


def test_flags_bool_spellings(monkeypatch):
    F.DEFINE_bool("a", True, "")
    F.DEFINE_bool("b", False, "")
    F.DEFINE_bool("c", False, "")
    monkeypatch.setattr("sys.argv", ["prog", "--noa", "--b=True", "--c", "false"])
    assert (F.FLAGS.a, F.FLAGS.b, F.FLAGS.c) == (False, True, False)


def test_log_device_placement_prints_each_node_once(capsys):
    import distributed_tensorflow_b200 as tf
    a = tf.constant([1.0, 2.0], name="a")
    with tf.device("/cpu:0"):
        b = tf.add(a, a, name="b")
    with tf.Session(config=tf.ConfigProto(log_device_placement=True)) as sess:
        sess.run(b)
        sess.run(b)
    err = capsys.readouterr().err
    assert err.count("a: (Const): ") == 1 and err.count("b: (Add): /device:CPU:0") == 1
