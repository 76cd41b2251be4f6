"""TFRecord files + ``tf.train.Example`` (utils/tfrecord.py): framing and message bytes checked against the protobuf runtime's own
encoder (a hand-built descriptor of ``tensorflow.Example``), parsing, and a ``TFRecordDataset`` pipeline feeding a training loop."""
import os
import struct

import numpy as np
import pytest

import distributed_tensorflow_b200 as tf
from distributed_tensorflow_b200.utils import tfrecord
from distributed_tensorflow_b200.utils.summary import masked_crc32c


@pytest.fixture(autouse=True)
def _fresh_graph():
    tf.reset_default_graph()
    yield
    tf.reset_default_graph()


def _example(i):
    img = (np.arange(6, dtype=np.float32) + i).reshape(2, 3)
    return tf.train.Example(features=tf.train.Features(feature={
        "image_raw": tf.train.Feature(bytes_list=tf.train.BytesList(value=[img.tobytes()])),
        "label": tf.train.Feature(int64_list=tf.train.Int64List(value=[i % 3])),
        "weights": tf.train.Feature(float_list=tf.train.FloatList(value=[0.5 * i, -1.0])),
        "tags": tf.train.Feature(bytes_list=tf.train.BytesList(value=[b"a"] * (i % 2 + 1))),
        "big": tf.train.Feature(int64_list=tf.train.Int64List(value=[-1, 2 ** 40 + i])),
    })), img


def test_record_framing_and_checksums(tmp_path):
    path = str(tmp_path / "d.tfrecord")
    with tf.python_io.TFRecordWriter(path) as w:
        for rec in (b"", b"hello", bytes(range(256)) * 5):
            w.write(rec)
    raw = open(path, "rb").read()
    (n,) = struct.unpack("<Q", raw[:8])
    assert n == 0 and struct.unpack("<I", raw[8:12])[0] == masked_crc32c(raw[:8])          # TensorFlow's layout
    assert list(tf.python_io.tf_record_iterator(path)) == [b"", b"hello", bytes(range(256)) * 5]
    bad = bytearray(raw)
    bad[-10] ^= 0xFF                                                   # inside the last payload
    open(path, "wb").write(bytes(bad))
    it = tf.python_io.tf_record_iterator(path)
    assert next(it) == b"" and next(it) == b"hello"
    with pytest.raises(tf.errors.DataLossError):
        next(it)


def test_example_bytes_are_what_the_protobuf_runtime_writes_and_reads():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="ex_test.proto", package="dtftest", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, ftype, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=ftype, label=label)
            if tname:
                f.type_name = tname
        return m
    T = descriptor_pb2.FieldDescriptorProto
    msg("BytesList", [("value", 1, T.TYPE_BYTES, T.LABEL_REPEATED, None)])
    msg("FloatList", [("value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, None)])
    msg("Int64List", [("value", 1, T.TYPE_INT64, T.LABEL_REPEATED, None)])
    feat = msg("Feature", [("bytes_list", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".dtftest.BytesList"),
                           ("float_list", 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".dtftest.FloatList"),
                           ("int64_list", 3, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".dtftest.Int64List")])
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    entry = msg("Entry", [("key", 1, T.TYPE_STRING, T.LABEL_OPTIONAL, None), ("value", 2, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".dtftest.Feature")])
    msg("Features", [("feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".dtftest.Entry")])       # a map IS a repeated entry on the wire
    msg("Example", [("features", 1, T.TYPE_MESSAGE, T.LABEL_OPTIONAL, ".dtftest.Features")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Ex = message_factory.GetMessageClass(pool.FindMessageTypeByName("dtftest.Example"))
    ours, img = _example(4)
    theirs = Ex()
    theirs.ParseFromString(ours.SerializeToString())                       # the runtime reads our bytes
    got = {e.key: e.value for e in theirs.features.feature}
    assert got["image_raw"].bytes_list.value[0] == img.tobytes() and list(got["label"].int64_list.value) == [1]
    assert list(got["weights"].float_list.value) == [2.0, -1.0] and list(got["big"].int64_list.value) == [-1, 2 ** 40 + 4]
    back = tf.train.Example.FromString(theirs.SerializeToString())         # and we read the runtime's bytes
    assert back.features.feature["big"].int64_list.value == [-1, 2 ** 40 + 4] and back.features.feature["tags"].bytes_list.value == [b"a"]
    assert back.features.feature["weights"].float_list.value == [2.0, -1.0]
    assert back.SerializeToString() == ours.SerializeToString()            # deterministic (sorted keys), round-trips exactly


def test_parse_single_example_fixed_var_default_and_errors():
    ex, img = _example(5)
    s = ex.SerializeToString()
    out = tf.parse_single_example(s, {"image_raw": tf.FixedLenFeature([], tf.string), "label": tf.FixedLenFeature([], tf.int64),
                                      "weights": tf.FixedLenFeature([2], tf.float32), "tags": tf.VarLenFeature(tf.string),
                                      "missing": tf.FixedLenFeature([2], tf.float32, default_value=7.0),
                                      "absent_var": tf.VarLenFeature(tf.int64)})
    assert np.array_equal(tf.decode_raw(out["image_raw"], tf.float32).reshape(2, 3), img) and out["label"] == 2 and out["label"].dtype == np.int64
    assert out["weights"].tolist() == [2.5, -1.0] and out["tags"].tolist() == [b"a", b"a"] and out["missing"].tolist() == [7.0, 7.0]
    assert out["absent_var"].shape == (0,)
    with pytest.raises(ValueError, match="no feature"):
        tf.parse_single_example(s, {"nope": tf.FixedLenFeature([], tf.int64)})
    with pytest.raises(ValueError, match="needs 3"):
        tf.parse_single_example(s, {"weights": tf.FixedLenFeature([3], tf.float32)})
    with pytest.raises(ValueError, match="float_list"):
        tf.parse_single_example(s, {"weights": tf.FixedLenFeature([2], tf.int64)})
    batch = tf.parse_example([_example(i)[0].SerializeToString() for i in range(4)], {"label": tf.FixedLenFeature([], tf.int64),
                                                                                       "weights": tf.FixedLenFeature([2], tf.float32)})
    assert batch["label"].tolist() == [0, 1, 2, 0] and batch["weights"].shape == (4, 2)


def test_tfrecord_dataset_feeds_a_training_loop(tmp_path):
    """y = 3x - 2 stored as Example records in two files, read back through TFRecordDataset -> map(parse) -> shuffle / repeat / batch ->
    get_next, and fitted by gradient descent: the record path end to end."""
    rng = np.random.default_rng(0)
    files = []
    for k in range(2):
        path = str(tmp_path / ("part-%d.tfrecord" % k))
        files.append(path)
        with tf.python_io.TFRecordWriter(path) as w:
            for _ in range(64):
                x = float(rng.uniform(-1, 1))
                w.write(tf.train.Example(features=tf.train.Features(feature={
                    "x": tf.train.Feature(float_list=tf.train.FloatList(value=[x])),
                    "y": tf.train.Feature(float_list=tf.train.FloatList(value=[3 * x - 2]))})).SerializeToString())
    spec = {"x": tf.FixedLenFeature([1], tf.float32), "y": tf.FixedLenFeature([1], tf.float32)}

    def parse(rec):
        d = tf.parse_single_example(rec, spec)
        return d["x"], d["y"]
    ds = tf.data.TFRecordDataset(files).map(parse).shuffle(128, seed=1).repeat().batch(16)
    assert sum(1 for _ in tf.data.TFRecordDataset(files)._make()) == 128
    xb, yb = ds.make_one_shot_iterator().get_next()
    w, b = tf.Variable(0.0, name="w"), tf.Variable(0.0, name="b")
    loss = tf.reduce_mean(tf.square(w * xb + b - yb))
    step = tf.train.GradientDescentOptimizer(0.3).minimize(loss)
    with tf.Session() as sess:
        sess.run(tf.global_variables_initializer())
        for _ in range(300):
            sess.run(step)
        assert sess.run(w) == pytest.approx(3.0, abs=0.02) and sess.run(b) == pytest.approx(-2.0, abs=0.02)


def test_native_scanner_agrees_with_the_python_reader_on_good_truncated_and_corrupt_files(tmp_path):
    """``dtf_tfrecord_scan`` (csrc/runtime/bundle_io.cpp) verifies both checksums of every record outside the interpreter; the
    iterator built on it yields what the record-by-record Python reader yields -- including the records before a corrupt one
    and the complete records of a file whose writer is still appending."""
    from distributed_tensorflow_b200.utils.summary import _read_tfrecords
    if tfrecord._scan_native(b"") is None:
        pytest.skip("native runtime library not built")
    rng = np.random.default_rng(3)
    recs = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in (0, 1, 7, 4096, 100_000, 3)]
    path = str(tmp_path / "r.tfrecord")
    with tfrecord.TFRecordWriter(path) as w:
        for r in recs:
            w.write(r)
    raw = open(path, "rb").read()
    off, ln, bad = tfrecord._scan_native(raw)
    assert bad == -1 and ln.tolist() == [len(r) for r in recs] and [raw[o:o + n] for o, n in zip(off.tolist(), ln.tolist())] == recs
    assert list(tfrecord.tf_record_iterator(path)) == recs == list(_read_tfrecords(path))
    # a writer that is still appending: header / payload / trailing checksum cut at every interesting place
    for cut in (len(raw) - 1, len(raw) - 4, len(raw) - 6, len(raw) - 12, len(raw) - 18):
        open(path, "wb").write(raw[:cut])
        assert list(tfrecord.tf_record_iterator(path)) == recs[:-1]
    # corruption in record 3's payload / in record 1's length field
    for pos, first_bad in ((int(off[3]) + 100, 3), (int(off[1]) - 12, 1)):
        dmg = bytearray(raw)
        dmg[pos] ^= 0x40
        open(path, "wb").write(bytes(dmg))
        it, got = tfrecord.tf_record_iterator(path), []
        with pytest.raises(tf.errors.DataLossError):
            for r in it:
                got.append(r)
        assert got == recs[:first_bad]
    open(path, "wb").write(b"")
    assert list(tfrecord.tf_record_iterator(path)) == []


def test_native_batch_parser_agrees_with_the_per_record_python_path():
    """``dtf_parse_examples`` (csrc/runtime/example_parser.cpp) against ``parse_single_example`` on the same records: packed and
    unpacked encodings, negative / 40-bit integers, multi-element features, bytes, unknown keys, defaults for absent keys; its
    error reports are left to the Python path."""
    from distributed_tensorflow_b200.utils.summary import _f_bytes, _key, _varint
    if tfrecord._parse_examples_native([b""], {}) is None and tfrecord._parse_examples_native([_example(0)[0].SerializeToString()],
                                                                                             {"label": tf.FixedLenFeature([], tf.int64)}) is None:
        pytest.skip("native runtime library not built")
    recs = [_example(i)[0].SerializeToString() for i in range(17)]
    spec = {"image_raw": tf.FixedLenFeature([], tf.string), "label": tf.FixedLenFeature([], tf.int64),
            "weights": tf.FixedLenFeature([2], tf.float32), "big": tf.FixedLenFeature([2, 1], tf.int64),
            "absent": tf.FixedLenFeature([3], tf.float32, default_value=[1.0, 2.0, 3.0])}
    nat = tfrecord._parse_examples_native(recs, spec)
    assert nat is not None
    for i, r in enumerate(recs):
        one = tf.parse_single_example(r, spec)
        for k in spec:
            assert np.array_equal(nat[k][i], one[k]) and np.shape(nat[k][i]) == one[k].shape, (i, k)
    assert nat["big"].shape == (17, 2, 1) and nat["big"][4, :, 0].tolist() == [-1, 2 ** 40 + 4] and nat["label"].dtype == np.int64
    imgs = tf.decode_raw(nat["image_raw"], tf.float32)                                  # batched: [17, 6]
    assert imgs.shape == (17, 6) and np.array_equal(imgs[5].reshape(2, 3), _example(5)[1])
    # unpacked repeated scalars (old writers): FloatList {1: fixed32}{1: fixed32}, Int64List {1: varint}{1: varint}
    fl = b"".join(_key(1, 5) + struct.pack("<f", v) for v in (0.25, -4.0))
    il = b"".join(_key(1, 0) + _varint(v) for v in (7, -3))
    entry = lambda k, feat: _f_bytes(1, _f_bytes(1, k.encode()) + _f_bytes(2, feat))          # noqa: E731
    rec = _f_bytes(1, entry("weights", _f_bytes(2, fl)) + entry("big", _f_bytes(3, il)) + entry("extra", _f_bytes(1, _f_bytes(1, b"zz"))))
    got = tfrecord._parse_examples_native([rec], {"weights": tf.FixedLenFeature([2], tf.float32), "big": tf.FixedLenFeature([2], tf.int64)})
    assert got["weights"].tolist() == [[0.25, -4.0]] and got["big"].tolist() == [[7, -3]]
    assert tf.parse_single_example(rec, {"big": tf.FixedLenFeature([2], tf.int64)})["big"].tolist() == [7, -3]
    # problems: the native parser steps aside (None) and parse_example raises what the Python path raises
    for bad_spec, msg in (({"weights": tf.FixedLenFeature([3], tf.float32)}, "needs 3"), ({"weights": tf.FixedLenFeature([2], tf.int64)}, "float_list"),
                          ({"nope": tf.FixedLenFeature([], tf.int64)}, "no feature")):
        assert tfrecord._parse_examples_native(recs, bad_spec) is None
        with pytest.raises(ValueError, match=msg):
            tf.parse_example(recs, bad_spec)
    assert tfrecord._parse_examples_native([recs[0][:-3]], {"label": tf.FixedLenFeature([], tf.int64)}) is None          # truncated record
    assert tfrecord._parse_examples_native(recs, {"tags": tf.VarLenFeature(tf.string)}) is None                          # not a fixed-length spec


def test_batch_then_parse_pipeline_matches_parse_then_batch(tmp_path):
    path = str(tmp_path / "d.tfrecord")
    with tf.python_io.TFRecordWriter(path) as w:
        for i in range(40):
            w.write(_example(i)[0].SerializeToString())
    spec = {"image_raw": tf.FixedLenFeature([], tf.string), "label": tf.FixedLenFeature([], tf.int64)}

    def parse_batch(recs):
        d = tf.parse_example(recs, spec)
        return tf.decode_raw(d["image_raw"], tf.float32), d["label"]

    def parse_one(rec):
        d = tf.parse_single_example(rec, spec)
        return tf.decode_raw(d["image_raw"], tf.float32), d["label"]
    a = list(tf.data.TFRecordDataset(path).batch(16).map(parse_batch)._make())
    b = list(tf.data.TFRecordDataset(path).map(parse_one).batch(16)._make())
    assert len(a) == len(b) == 3 and a[2][0].shape == (8, 6)
    for (xa, la), (xb, lb) in zip(a, b):
        assert np.array_equal(xa, xb) and np.array_equal(la, lb)


def test_read_all_bulk_loads_record_files_for_the_epoch_batcher(tmp_path):
    """Record files -> in-memory arrays (``tfrecord.read_all``) -> ``EpochBatcher`` (shuffled epochs in one buffer, native row
    gather): the bulk path in front of the fabric engine's host-fed loop."""
    from distributed_tensorflow_b200.utils.input_pipeline import EpochBatcher
    files = []
    for k in range(3):
        path = str(tmp_path / ("p%d.tfrecord" % k))
        files.append(path)
        with tf.python_io.TFRecordWriter(path) as w:
            for i in range(k * 10, k * 10 + 10):
                w.write(tf.train.Example(features=tf.train.Features(feature={
                    "x": tf.train.Feature(float_list=tf.train.FloatList(value=[i, i + 0.5, -i])),
                    "y": tf.train.Feature(int64_list=tf.train.Int64List(value=[i % 4]))})).SerializeToString())
    spec = {"x": tf.FixedLenFeature([3], tf.float32), "y": tf.FixedLenFeature([], tf.int64)}
    got = tfrecord.read_all(files, spec, chunk=7)                      # several parser calls per file boundary
    assert got["x"].shape == (30, 3) and got["x"][:, 0].tolist() == list(map(float, range(30))) and got["y"].tolist() == [i % 4 for i in range(30)]
    assert tfrecord.read_all([], spec)["x"].shape == (0, 3)
    onehot = np.eye(4, dtype=np.float32)[got["y"]]
    eb = EpochBatcher(got["x"], onehot, batch=5, seed=0, pin=False)
    xb, yb = eb.next_epoch()
    assert tuple(xb.shape) == (6, 5, 3) and sorted(np.asarray(xb)[:, :, 0].reshape(-1).tolist()) == list(map(float, range(30)))
    rows = np.asarray(xb).reshape(-1, 3)
    assert np.array_equal(np.asarray(yb).reshape(-1, 4).argmax(1), rows[:, 0].astype(int) % 4)          # labels travel with their rows


def test_native_parsers_survive_mutated_records_under_address_and_ub_sanitizers(tmp_path):
    """Record files are external input: the C++ scanner and Example parser are fuzzed with byte-flipped / truncated / extended
    records in exact-size heap buffers under ASAN + UBSAN (tests/emu/fuzz_record_parsers.cpp)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rt = os.path.join(root, "distributed_tensorflow_b200", "csrc", "runtime")
    exe = str(tmp_path / "fuzz")
    cc = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                         "-o", exe, os.path.join(root, "tests", "emu", "fuzz_record_parsers.cpp"), os.path.join(rt, "example_parser.cpp"),
                         os.path.join(rt, "bundle_io.cpp"), "-pthread"], capture_output=True, text=True)
    if cc.returncode != 0:
        pytest.skip("sanitizer build not available here: %s" % cc.stderr[-300:])
    seeds = str(tmp_path / "seeds.bin")
    with open(seeds, "wb") as f:
        recs = []
        for i in range(8):
            recs.append(tf.train.Example(features=tf.train.Features(feature={
                "image_raw": tf.train.Feature(bytes_list=tf.train.BytesList(value=[(np.arange(6, dtype=np.float32) + i).tobytes()])),
                "label": tf.train.Feature(int64_list=tf.train.Int64List(value=[i % 3, -i])),
                "weights": tf.train.Feature(float_list=tf.train.FloatList(value=[0.5 * i, -1.0, 2.0]))})).SerializeToString())
        f.write(struct.pack("<I", len(recs)))
        for r in recs:
            f.write(struct.pack("<I", len(r)) + r)
    r = subprocess.run([exe, seeds, "60000"], capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "parsed ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])
    ok, rejected = (int(x) for x in r.stdout.replace(",", "").split() if x.isdigit())
    assert ok > 1000 and rejected > 1000                       # both outcomes exercised
