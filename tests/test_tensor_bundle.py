"""Checkpoints are TensorFlow checkpoint-V2 tensor bundles: ``.index`` = LevelDB-format table of BundleEntryProto,
``.data-00000-of-00001`` = raw bytes.  TensorFlow is not available offline; the test re-parses the table with an
independent reader written from the format description, parses the protos with the real protobuf runtime and
re-derives every checksum."""
import struct

import numpy as np
import pytest
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.train import saver as saver_mod
from distributed_tensorflow_b200.train import tensor_bundle as tb
from distributed_tensorflow_b200.utils.summary import crc32c


def _proto_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="dtf_bundle_min.proto", package="tensorflow", syntax="proto3")

    def msg(parent, name):
        m = parent.message_type.add() if isinstance(parent, descriptor_pb2.FileDescriptorProto) else parent.nested_type.add()
        m.name = name
        return m

    def field(m, name, num, typ, label=F.LABEL_OPTIONAL, type_name=None):
        f = m.field.add(name=name, number=num, type=typ, label=label)
        if type_name:
            f.type_name = type_name
    shape = msg(fd, "TensorShapeProto")
    dim = msg(shape, "Dim")
    field(dim, "size", 1, F.TYPE_INT64)
    field(shape, "dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".tensorflow.TensorShapeProto.Dim")
    ver = msg(fd, "VersionDef")
    field(ver, "producer", 1, F.TYPE_INT32)
    hdr = msg(fd, "BundleHeaderProto")
    field(hdr, "num_shards", 1, F.TYPE_INT32)
    field(hdr, "endianness", 2, F.TYPE_INT32)
    field(hdr, "version", 3, F.TYPE_MESSAGE, type_name=".tensorflow.VersionDef")
    ent = msg(fd, "BundleEntryProto")
    field(ent, "dtype", 1, F.TYPE_INT32)
    field(ent, "shape", 2, F.TYPE_MESSAGE, type_name=".tensorflow.TensorShapeProto")
    field(ent, "shard_id", 3, F.TYPE_INT32)
    field(ent, "offset", 4, F.TYPE_INT64)
    field(ent, "size", 5, F.TYPE_INT64)
    field(ent, "crc32c", 6, F.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


def _mask(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(buf, i):
    v = s = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return v, i


def _independent_table_read(buf):
    """LevelDB table reader written separately from train/tensor_bundle.py (footer -> index block -> data blocks)."""
    assert struct.unpack("<Q", buf[-8:])[0] == 0xDB4775248B80FB57
    foot = buf[-48:-8]
    _, i = _varint(foot, 0)
    _, i = _varint(foot, i)
    ioff, i = _varint(foot, i)
    isz, i = _varint(foot, i)

    def block(off, size):
        body, typ, crc = buf[off:off + size], buf[off + size], struct.unpack("<I", buf[off + size + 1:off + size + 5])[0]
        assert typ == 0 and crc == _mask(crc32c(body + bytes([typ])))           # per-block masked CRC32C
        n = struct.unpack("<I", body[-4:])[0]
        end, out, i, key = len(body) - 4 - 4 * n, [], 0, b""
        restarts = struct.unpack("<%dI" % n, body[end:end + 4 * n])
        while i < end:
            if i in restarts:
                pass
            shared, i = _varint(body, i)
            non_shared, i = _varint(body, i)
            vlen, i = _varint(body, i)
            key = key[:shared] + body[i:i + non_shared]
            i += non_shared
            out.append((key, body[i:i + vlen]))
            i += vlen
        return out
    items = []
    for sep, handle in block(ioff, isz):
        off, j = _varint(handle, 0)
        size, _ = _varint(handle, j)
        blk = block(off, size)
        assert blk[-1][0] <= sep                                                # index key >= last key of its block
        items += blk
    return items


def test_bundle_index_is_a_leveldb_table_of_bundle_entries(tmp_path):
    Header, Entry = _proto_classes()
    tensors = {"hid_w": torch.randn(784, 100), "hid_b": torch.zeros(100), "global_step": torch.tensor(1200, dtype=torch.int64),
               "flags/ok": torch.tensor([True, False, True]), "half/w": torch.randn(7, 3).to(torch.bfloat16),
               "beta1_power": torch.tensor(0.9 ** 5)}
    prefix = str(tmp_path / "model.ckpt-1200")
    saver_mod.write_bundle(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    items = _independent_table_read(raw)
    keys = [k for k, _ in items]
    assert keys == sorted(keys) and keys[0] == b""                               # sorted, header first
    h = Header()
    h.ParseFromString(items[0][1])
    assert h.num_shards == 1 and h.endianness == 0 and h.version.producer == 1
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    dt = {"float32": 1, "int64": 9, "bool": 10, "bfloat16": 14}
    for key, val in items[1:]:
        e = Entry()
        e.ParseFromString(val)
        t = tensors[key.decode()]
        assert e.dtype == dt[str(t.dtype).replace("torch.", "")]
        assert [d.size for d in e.shape.dim] == list(t.shape)
        assert e.shard_id == 0 and e.size == t.numel() * t.element_size()
        blob = data[e.offset:e.offset + e.size]
        assert e.crc32c == _mask(crc32c(blob))                                   # masked CRC32C of the tensor bytes
        if t.dtype == torch.float32:
            np.testing.assert_array_equal(np.frombuffer(blob, np.float32).reshape(t.shape), t.numpy())
    # and the framework's own reader round-trips every tensor bit-exactly
    r = saver_mod.CheckpointReader(prefix)
    assert sorted(r.get_variable_to_shape_map()) == sorted(tensors)
    for k, t in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == t.dtype and torch.equal(got, t)


def test_corruption_is_detected_and_legacy_json_index_still_reads(tmp_path):
    prefix = str(tmp_path / "m.ckpt-1")
    saver_mod.write_bundle(prefix, {"w": torch.arange(10, dtype=torch.float32)})
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[4] ^= 0xFF
    open(prefix + ".data-00000-of-00001", "wb").write(raw)
    with pytest.raises(dtf.errors.OpError):
        saver_mod.CheckpointReader(prefix).get_tensor("w")
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 0x01
    open(prefix + ".index", "wb").write(idx)
    with pytest.raises(ValueError):
        tb.read_index(prefix + ".index")
    # an index written by an earlier version of this framework (JSON + zlib crc32)
    import json
    import zlib
    p2 = str(tmp_path / "old.ckpt-3")
    b = np.arange(6, dtype=np.float32).tobytes()
    open(p2 + ".data-00000-of-00001", "wb").write(b)
    json.dump({"format": "dtf-bundle-v1", "total_bytes": len(b), "tensors": {
        "w": {"dtype": "float32", "shape": [2, 3], "offset": 0, "nbytes": len(b), "crc32": zlib.crc32(b) & 0xFFFFFFFF}}},
        open(p2 + ".index", "w"))
    assert torch.equal(saver_mod.CheckpointReader(p2).get_tensor("w"), torch.arange(6, dtype=torch.float32).reshape(2, 3))


def test_meta_file_is_a_meta_graph_def(tmp_path):
    """``model.ckpt-N.meta`` parses as MetaGraphDef {meta_info_def, graph_def, saver_def} with the protobuf runtime."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto(name="dtf_meta_min.proto", package="tensorflow", syntax="proto3")

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
    O, R = F.LABEL_OPTIONAL, F.LABEL_REPEATED
    msg("NodeDef", [("name", 1, F.TYPE_STRING, O, None), ("op", 2, F.TYPE_STRING, O, None), ("input", 3, F.TYPE_STRING, R, None),
                    ("device", 4, F.TYPE_STRING, O, None)])
    msg("GraphDef", [("node", 1, F.TYPE_MESSAGE, R, ".tensorflow.NodeDef")])
    msg("MetaInfoDef", [("tags", 4, F.TYPE_STRING, R, None), ("tensorflow_version", 5, F.TYPE_STRING, O, None)])
    msg("SaverDef", [("filename_tensor_name", 1, F.TYPE_STRING, O, None), ("save_tensor_name", 2, F.TYPE_STRING, O, None),
                     ("restore_op_name", 3, F.TYPE_STRING, O, None), ("max_to_keep", 4, F.TYPE_INT32, O, None),
                     ("keep_checkpoint_every_n_hours", 6, F.TYPE_FLOAT, O, None), ("version", 7, F.TYPE_INT32, O, None)])
    msg("MetaGraphDef", [("meta_info_def", 1, F.TYPE_MESSAGE, O, ".tensorflow.MetaInfoDef"), ("graph_def", 2, F.TYPE_MESSAGE, O, ".tensorflow.GraphDef"),
                         ("saver_def", 3, F.TYPE_MESSAGE, O, ".tensorflow.SaverDef")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    Meta = message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow.MetaGraphDef"))
    w = dtf.Variable(dtf.zeros([2, 2]), name="w")
    gs = dtf.train.get_or_create_global_step()
    with dtf.Session() as sess:
        sess.run(dtf.global_variables_initializer())
        path = dtf.train.Saver(max_to_keep=3).save(sess, str(tmp_path / "model.ckpt"), global_step=7)
    m = Meta()
    m.ParseFromString(open(path + ".meta", "rb").read())
    assert m.meta_info_def.tensorflow_version.startswith("1.12") and list(m.meta_info_def.tags) == ["train"]
    assert {"w", "global_step"} <= {n.name for n in m.graph_def.node}
    assert m.saver_def.max_to_keep == 3 and m.saver_def.version == 2 and m.saver_def.restore_op_name == "save/restore_all"
