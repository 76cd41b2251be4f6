"""The event files are real TensorBoard ``tfevents`` files: TFRecord framing + protobuf wire format.  TensorFlow's
.proto files are not available offline, so the test declares the same messages (field numbers from
tensorflow/core/util/event.proto, framework/{graph,node_def,attr_value,tensor_shape,summary}.proto) with the protobuf
runtime and parses what ``FileWriter`` wrote with it."""
import struct

import pytest

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.utils import summary as S


def _messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="dtf_tb_min.proto", package="tensorflow", syntax="proto3")

    def msg(parent, name):
        m = parent.message_type.add() if isinstance(parent, descriptor_pb2.FileDescriptorProto) else parent.nested_type.add()
        m.name = name
        return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None):
        f = m.field.add(name=name, number=num, type=typ, label=label)
        if type_name:
            f.type_name = type_name
        return f
    shape = msg(fd, "TensorShapeProto")
    dim = msg(shape, "Dim")
    field(dim, "size", 1, F.TYPE_INT64)
    field(dim, "name", 2, F.TYPE_STRING)
    field(shape, "dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.TensorShapeProto.Dim")
    attr = msg(fd, "AttrValue")
    lst = msg(attr, "ListValue")
    field(lst, "shape", 7, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.TensorShapeProto")
    field(attr, "list", 1, F.TYPE_MESSAGE, type_name=".tensorflow.AttrValue.ListValue")
    field(attr, "type", 6, F.TYPE_INT32)
    node = msg(fd, "NodeDef")
    field(node, "name", 1, F.TYPE_STRING)
    field(node, "op", 2, F.TYPE_STRING)
    field(node, "input", 3, F.TYPE_STRING, F.LABEL_REPEATED)
    field(node, "device", 4, F.TYPE_STRING)
    entry = msg(node, "AttrEntry")
    entry.options.map_entry = True
    field(entry, "key", 1, F.TYPE_STRING)
    field(entry, "value", 2, F.TYPE_MESSAGE, type_name=".tensorflow.AttrValue")
    field(node, "attr", 5, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.NodeDef.AttrEntry")
    ver = msg(fd, "VersionDef")
    field(ver, "producer", 1, F.TYPE_INT32)
    graph = msg(fd, "GraphDef")
    field(graph, "node", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.NodeDef")
    field(graph, "versions", 4, F.TYPE_MESSAGE, type_name=".tensorflow.VersionDef")
    summ = msg(fd, "Summary")
    val = msg(summ, "Value")
    field(val, "tag", 1, F.TYPE_STRING)
    field(val, "simple_value", 2, F.TYPE_FLOAT)
    field(summ, "value", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.Summary.Value")
    trm = msg(fd, "TaggedRunMetadata")
    field(trm, "tag", 1, F.TYPE_STRING)
    field(trm, "run_metadata", 2, F.TYPE_BYTES)
    ev = msg(fd, "Event")
    field(ev, "wall_time", 1, F.TYPE_DOUBLE)
    field(ev, "step", 2, F.TYPE_INT64)
    field(ev, "file_version", 3, F.TYPE_STRING)
    field(ev, "graph_def", 4, F.TYPE_BYTES)
    field(ev, "summary", 5, F.TYPE_MESSAGE, type_name=".tensorflow.Summary")
    field(ev, "tagged_run_metadata", 8, F.TYPE_MESSAGE, type_name=".tensorflow.TaggedRunMetadata")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + n))
    return get("Event"), get("GraphDef")


def test_crc32c_and_mask_known_vectors():
    assert S.crc32c(b"123456789") == 0xE3069283                       # the CRC-32C check value
    assert S.crc32c(b"") == 0
    c = S.crc32c(b"123456789")
    assert S.masked_crc32c(b"123456789") == ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def test_event_file_is_tfrecord_of_event_protos(tmp_path):
    Event, GraphDef = _messages()
    with dtf.device("/job:ps/task:0/cpu:0"):
        w = dtf.Variable(dtf.zeros([3, 2]), name="w")
    x = dtf.placeholder(dtf.float32, [None, 3], name="x")
    with dtf.control_dependencies([w.initializer]):
        y = dtf.matmul(x, w, name="y")
    g = dtf.get_default_graph()
    writer = dtf.summary.FileWriter(str(tmp_path), g)
    writer.add_scalar("loss", 0.25, global_step=7)
    writer.add_summary({"a": 1.5, "b": -2.0}, global_step=8)
    writer.close()
    assert "events.out.tfevents." in writer.path
    raw = open(writer.path, "rb").read()
    # TFRecord framing: u64 length, masked crc of the length, payload, masked crc of the payload
    (n0,) = struct.unpack("<Q", raw[:8])
    assert struct.unpack("<I", raw[8:12])[0] == S.masked_crc32c(raw[:8])
    assert struct.unpack("<I", raw[12 + n0:16 + n0])[0] == S.masked_crc32c(raw[12:12 + n0])
    events = []
    for payload in S._read_tfrecords(writer.path):
        e = Event()
        e.ParseFromString(payload)                                      # the real protobuf runtime accepts every record
        events.append(e)
    assert events[0].file_version == "brain.Event:2" and events[0].wall_time > 1e9
    gd = GraphDef()
    gd.ParseFromString(events[1].graph_def)
    by_name = {n.name: n for n in gd.node}
    assert by_name["y"].op == "MatMul" and list(by_name["y"].input)[:2] == ["x", by_name["y"].input[1]]
    assert any(i.startswith("^") for i in by_name["y"].input)           # the control dependency is an input "^name"
    assert by_name["w"].device.endswith("/job:ps/task:0/device:CPU:0") or "ps" in by_name["w"].device
    assert by_name["x"].attr["T"].type == 1                             # DT_FLOAT
    dims = [d.size for d in by_name["x"].attr["_output_shapes"].list.shape[0].dim]
    assert dims == [-1, 3]
    assert gd.versions.producer == 27
    assert events[2].step == 7 and events[2].summary.value[0].tag == "loss"
    assert events[2].summary.value[0].simple_value == pytest.approx(0.25)
    assert [(v.tag, v.simple_value) for v in events[3].summary.value] == [("a", 1.5), ("b", -2.0)]
    # and our own reader sees the same things
    mine = dtf.summary.read_events(writer.path)
    assert mine[0]["file_version"] == "brain.Event:2"
    assert any(n["name"] == "y" and n["op"] == "MatMul" for n in mine[1]["graph_def"]["node"])
    assert mine[2]["scalar"] == {"tag": "loss", "value": 0.25} and mine[2]["step"] == 7


def test_histogram_summaries_are_real_histogram_protos(tmp_path):
    import numpy as np
    import distributed_tensorflow_b200 as tf
    from distributed_tensorflow_b200.utils import summary as S
    w = tf.get_variable("hw", [200], initializer=tf.truncated_normal_initializer(stddev=0.5, seed=3))
    tf.summary.histogram("weights", w)
    tf.summary.scalar("mean_w", tf.reduce_mean(w))
    merged = tf.summary.merge_all()
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        wv, out = sess.run([w, merged])
        with tf.summary.FileWriter(str(tmp_path)) as fw:
            fw.add_summary(out, 7)
            path = fw.path
    ev = [e for e in S.read_events(path) if "histograms" in e or "scalars" in e][-1]
    assert ev["step"] == 7 and ev["scalars"][0]["tag"] == "mean_w"
    h = ev["histograms"][0]["histo"]
    assert ev["histograms"][0]["tag"] == "weights" and h["num"] == 200.0
    assert h["min"] == float(wv.min()) and h["max"] == float(wv.max())
    np.testing.assert_allclose(h["sum"], wv.astype(np.float64).sum(), rtol=1e-12)
    np.testing.assert_allclose(h["sum_squares"], (wv.astype(np.float64) ** 2).sum(), rtol=1e-12)
    assert sum(h["bucket"]) == 200.0 and len(h["bucket"]) == len(h["bucket_limit"])
    lim = np.asarray(h["bucket_limit"])
    assert np.all(np.diff(lim) > 0) and lim[0] >= wv.min() and lim[-1] >= wv.max()      # trimmed to the occupied range
    # every value sits in the bucket whose limit is the first one >= the value
    idx = np.searchsorted(lim, wv.astype(np.float64), side="left")
    assert np.array_equal(np.bincount(idx, minlength=len(lim)).astype(float), np.asarray(h["bucket"]))
    assert set(tf.summary.merge([tf.summary.scalar("a", tf.constant(1.0))])) == {"a"}
