"""Automatic routing of tf.train programs onto the fabric (parallel/auto_fabric.py): the graph pattern matcher and the
conditions under which ``Optimizer.minimize`` keeps the ordinary graph-tier update (no GPU needed for either)."""
import math

import pytest

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.parallel.auto_fabric import match_reference_mlp, maybe_route_minimize


def _reference_model(hidden=100, classes=10, clip=1e-10, extra_var=False, mean=False):
    x = dtf.placeholder(dtf.float32, [None, 784])
    y_ = dtf.placeholder(dtf.float32, [None, classes])
    hid_w = dtf.Variable(dtf.truncated_normal([784, hidden], stddev=1.0 / 28), name="hid_w")
    hid_b = dtf.Variable(dtf.zeros([hidden]), name="hid_b")
    sm_w = dtf.Variable(dtf.truncated_normal([hidden, classes], stddev=1.0 / math.sqrt(hidden)), name="sm_w")
    sm_b = dtf.Variable(dtf.zeros([classes]), name="sm_b")
    if extra_var:
        dtf.Variable(dtf.zeros([3]), name="unused")
    hid = dtf.nn.relu(dtf.nn.xw_plus_b(x, hid_w, hid_b))
    y = dtf.nn.softmax(dtf.nn.xw_plus_b(hid, sm_w, sm_b))
    red = dtf.reduce_mean if mean else dtf.reduce_sum
    return -red(y_ * dtf.log(dtf.clip_by_value(y, clip, 1.0))), (x, y_, hid_w, hid_b, sm_w, sm_b)


def test_matcher_recognises_the_reference_network():
    loss, (x, y_, hid_w, hid_b, sm_w, sm_b) = _reference_model(hidden=64, classes=7)
    m = match_reference_mlp(dtf.convert_to_tensor(loss), dtf.trainable_variables())
    assert m is not None
    assert m["x"] is dtf.convert_to_tensor(x) and m["y_"] is dtf.convert_to_tensor(y_)
    assert (m["hid_w"], m["hid_b"], m["sm_w"], m["sm_b"]) == (hid_w, hid_b, sm_w, sm_b)
    assert (m["in_dim"], m["hidden"], m["classes"], m["clip_min"]) == (784, 64, 7, 1e-10)


def test_matcher_recognises_the_fused_loss_node_too():
    from distributed_tensorflow_b200.models import build_mnist_mlp
    for fused in (False, True):
        dtf.reset_default_graph()
        net = build_mnist_mlp(hidden=32, fused=fused)
        m = match_reference_mlp(dtf.convert_to_tensor(net["loss"]), dtf.trainable_variables())
        assert m is not None and (m["hidden"], m["classes"]) == (32, 10) and m["hid_w"] is net["vars"][0]


@pytest.mark.parametrize("kw", [dict(extra_var=True), dict(mean=True), dict(clip=0.1)])
def test_matcher_declines_anything_else(kw):
    loss, _ = _reference_model(**kw)
    assert match_reference_mlp(dtf.convert_to_tensor(loss), dtf.trainable_variables()) is None


def test_minimize_stays_on_the_graph_tier_without_a_gpu_worker_task(monkeypatch):
    """No worker Server bound to a GPU in this process: minimize builds the ordinary update (auto mode never raises)."""
    loss, _ = _reference_model()
    gs = dtf.train.get_or_create_global_step()
    opt = dtf.train.AdamOptimizer(0.01)
    assert maybe_route_minimize(opt, loss, gs) is None
    train_op = opt.minimize(loss, global_step=gs)
    assert getattr(opt, "_fabric_strategy", None) is None and train_op is not None
    monkeypatch.setenv("DTF_FABRIC", "0")
    assert maybe_route_minimize(opt, loss, gs) is None


def test_session_fetch_override_answers_and_declines():
    """The Session mechanism the routed loss fetch rides on: an override answers a fetch, or declines (NotImplemented) and the
    tensor is computed through the graph as usual."""
    a = dtf.constant(3.0)
    b = a * 2.0
    c = b + 1.0
    g = dtf.get_default_graph()
    calls = []

    def over(sess, feeds, fetch_ids):
        calls.append(sorted(fetch_ids))
        return NotImplemented if len(calls) == 1 else dtf.convert_to_tensor(0) is None or __import__("torch").tensor(42.0)
    g.__dict__.setdefault("_fetch_overrides", {})[dtf.convert_to_tensor(b).id] = over
    with dtf.Session() as sess:
        assert sess.run([b, c]) == [6.0, 7.0]            # declined: computed normally
        assert sess.run([b, c]) == [42.0, 7.0]           # answered: c still comes from the graph
    assert len(calls) == 2
