"""Property-based tests (hypothesis) of the pure-logic pieces every distributed program leans on: device strings,
round-robin placement, the wire codec, the accumulator's mean / stale-drop rule and the tensor-bundle table."""
import pickle

import numpy as np
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.framework.device import DeviceSpec
from distributed_tensorflow_b200.parallel.rpc import from_wire, to_wire
from distributed_tensorflow_b200.train import tensor_bundle as tb

jobs = st.sampled_from([None, "ps", "worker"])
idx = st.one_of(st.none(), st.integers(0, 7))
dev_types = st.sampled_from([None, "CPU", "GPU"])


def _spec(job, task, dtype, dindex):
    return DeviceSpec(job=job, task=task, device_type=dtype, device_index=(dindex if dtype else None))


@settings(max_examples=200, deadline=None)
@given(jobs, idx, dev_types, idx)
def test_device_string_roundtrip(job, task, dtype, dindex):
    s = _spec(job, task, dtype, dindex)
    again = DeviceSpec.from_string(s.to_string())
    assert (again.job, again.task, again.device_type, again.device_index) == (s.job, s.task, s.device_type, s.device_index)


@settings(max_examples=200, deadline=None)
@given(jobs, idx, dev_types, idx, jobs, idx, dev_types, idx)
def test_device_merge_inner_wins_per_field(j1, t1, d1, i1, j2, t2, d2, i2):
    outer, inner = _spec(j1, t1, d1, i1), _spec(j2, t2, d2, i2)
    m = outer.merge_from(inner)
    assert m.job == (inner.job if inner.job is not None else outer.job)
    assert m.task == (inner.task if inner.task is not None else outer.task)
    assert m.device_type == (inner.device_type if inner.device_type is not None else outer.device_type)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 5), st.integers(1, 12))
def test_replica_device_setter_is_round_robin_over_ps(num_ps, num_vars):
    dtf.reset_default_graph()
    cluster = dtf.train.ClusterSpec({"ps": ["h:%d" % (2000 + i) for i in range(num_ps)], "worker": ["h:3000"]})
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0")):
        vs = [dtf.Variable(dtf.zeros([2]), name="v%d" % i) for i in range(num_vars)]
        op = dtf.add(vs[0], vs[-1])
    for i, v in enumerate(vs):
        assert DeviceSpec.from_string(v.device).job == "ps" and DeviceSpec.from_string(v.device).task == i % num_ps
    assert DeviceSpec.from_string(op.device).job == "worker"


tensor_values = st.one_of(
    st.lists(st.floats(-1e6, 1e6, allow_nan=False, width=32), min_size=0, max_size=20).map(lambda l: torch.tensor(l, dtype=torch.float32)),
    st.lists(st.integers(-2 ** 40, 2 ** 40), min_size=1, max_size=8).map(lambda l: torch.tensor(l, dtype=torch.int64)),
    st.floats(-100, 100, allow_nan=False, width=16).map(lambda f: torch.tensor(f, dtype=torch.bfloat16)),
    st.lists(st.booleans(), min_size=1, max_size=9).map(lambda l: torch.tensor(l)))
payloads = st.recursive(st.one_of(tensor_values, st.integers(-5, 5), st.text(max_size=5), st.none()),
                        lambda c: st.one_of(st.lists(c, max_size=3), st.tuples(c, c), st.dictionaries(st.text(max_size=3), c, max_size=3)),
                        max_leaves=8)


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return isinstance(b, torch.Tensor) and a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b)
    if isinstance(a, dict):
        return isinstance(b, dict) and a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return type(a) is type(b) and len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    return a == b


@settings(max_examples=150, deadline=None)
@given(payloads)
def test_rpc_wire_codec_roundtrips_nested_payloads(value):
    assert _same(value, from_wire(pickle.loads(pickle.dumps(to_wire(value), protocol=pickle.HIGHEST_PROTOCOL))))


@settings(max_examples=60, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 3), st.floats(-10, 10, allow_nan=False, width=32)), min_size=1, max_size=12),
       st.integers(0, 3))
def test_accumulator_mean_of_fresh_gradients_only(pushes, global_step):
    """gradients stamped before the accumulator's global step are dropped; take_grad = MEAN of the rest."""
    from distributed_tensorflow_b200.utils import native_runtime
    acc = native_runtime.make_accumulator("p")
    acc.set_global_step(global_step)
    fresh = []
    for stamp, g in pushes:
        accepted = acc.apply_grad(torch.full((3,), g), stamp)
        assert accepted == (stamp >= global_step)
        if accepted:
            fresh.append(g)
    assert acc.num_accumulated() == len(fresh)
    if fresh:
        got = acc.take_grad(len(fresh), timeout=5.0)
        np.testing.assert_allclose(got.numpy(), np.full(3, np.mean(np.float32(fresh), dtype=np.float64)), rtol=1e-5, atol=1e-5)


@settings(max_examples=40, deadline=None)
@given(st.dictionaries(st.text(alphabet="abcdefghij/_0123456789", min_size=1, max_size=24),
                       st.tuples(st.sampled_from(["float32", "int64", "bfloat16", "bool"]),
                                 st.lists(st.integers(0, 300), max_size=4), st.integers(0, 2 ** 40), st.integers(0, 2 ** 31),
                                 st.integers(0, 2 ** 32 - 1)), max_size=40))
def test_tensor_bundle_table_roundtrips_any_entry_set(tmp_path_factory, entries):
    path = str(tmp_path_factory.mktemp("tb") / "x.index")
    tb.write_index(path, {k: tb.entry_proto(dt, shape, off, size, crc) for k, (dt, shape, off, size, crc) in entries.items()})
    header, got = tb.read_index(path)
    assert header["num_shards"] == 1 and set(got) == set(entries)
    for k, (dt, shape, off, size, crc) in entries.items():
        e = got[k]
        assert (e["dtype"], e["shape"], e["offset"], e["size"], e["crc32c"]) == (dt, list(shape), off, size, crc)
