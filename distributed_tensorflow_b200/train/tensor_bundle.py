"""TensorFlow "tensor bundle" (checkpoint V2) index: an immutable sorted table (LevelDB table format) mapping tensor name
-> ``BundleEntryProto``, next to a ``<prefix>.data-00000-of-00001`` file of raw little-endian tensor bytes.

This is the on-disk format behind ``tf.train.Saver`` / ``SaveV2`` / ``RestoreV2`` in TF >= 1.0 (the reference's
checkpoints: ``distributed_mnist.py:144-147``, ``distributed_mnist_predict.py:36-40``).  Layout written here:

* table = ``[data block][metaindex block][index block][footer]``; every block is followed by a 5-byte trailer (compression
  type 0 + masked CRC32C of contents+type); the 48-byte footer holds the metaindex and index block handles (varint64 offset,
  size) padded to 40 bytes and the magic ``0xdb4775248b80fb57``;
* a block = prefix-compressed entries ``varint(shared) varint(non_shared) varint(value_len) key_delta value`` followed by
  the restart offsets (fixed32 each) and their count; one restart point per entry is written (always ``shared = 0``),
  the reader handles any restart interval;
* key ``""`` -> ``BundleHeaderProto{num_shards=1, endianness=LITTLE, version{producer=1}}``; every other key is a tensor
  name -> ``BundleEntryProto{dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6 (fixed32, masked CRC32C of the bytes)}``.

TensorFlow itself is not available offline, so compatibility was checked against the format description only:
``tests/test_tensor_bundle.py`` parses the protos with the real protobuf runtime and re-derives every checksum.
"""
from __future__ import annotations

import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

from ..utils.summary import _decode, _f_bytes, _f_varint, _key, _varint, masked_crc32c as _py_masked_crc32c

__all__ = ["write_index", "read_index", "masked_crc32c", "TABLE_MAGIC", "DT_OF", "NAME_OF_DT"]

TABLE_MAGIC = 0xDB4775248B80FB57
# tensorflow.DataType values
DT_OF = {"float32": 1, "float64": 2, "int32": 3, "uint8": 4, "int16": 5, "int8": 6, "int64": 9, "bool": 10,
         "bfloat16": 14, "float16": 19}
NAME_OF_DT = {v: k for k, v in DT_OF.items()}


def masked_crc32c(data) -> int:
    """Masked CRC32C (``((crc >> 15) | (crc << 17)) + 0xa282ead8``), native (SSE4.2) when the runtime library is built."""
    from ..utils import native_runtime
    lib = native_runtime.load()
    if lib is not None and hasattr(lib, "dtf_crc32c"):
        mv = memoryview(data)
        c = _crc_native(lib, mv)
        return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF
    return _py_masked_crc32c(bytes(data))


def _crc_native(lib, mv: memoryview) -> int:
    import ctypes
    if not getattr(lib, "_crc32c_declared", False):
        lib.dtf_crc32c.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint32]
        lib.dtf_crc32c.restype = ctypes.c_uint32
        lib._crc32c_declared = True
    if mv.nbytes == 0:
        return 0
    buf = (ctypes.c_char * mv.nbytes).from_buffer_copy(mv) if mv.readonly else (ctypes.c_char * mv.nbytes).from_buffer(mv)
    return int(lib.dtf_crc32c(ctypes.addressof(buf), mv.nbytes, 0))


# ------------------------------------------------------------------------------------------------------------------
# protos
# ------------------------------------------------------------------------------------------------------------------
def _shape(shape: Sequence[int]) -> bytes:
    return b"".join(_f_bytes(2, _f_varint(1, int(d))) for d in shape)


def entry_proto(dtype: str, shape: Sequence[int], offset: int, size: int, crc_masked: int, shard_id: int = 0) -> bytes:
    out = _f_varint(1, DT_OF[dtype]) + _f_bytes(2, _shape(shape))
    if shard_id:
        out += _f_varint(3, shard_id)
    if offset:
        out += _f_varint(4, offset)
    out += _f_varint(5, size)
    out += _key(6, 5) + struct.pack("<I", crc_masked)
    return out


def header_proto(num_shards: int = 1) -> bytes:
    return _f_varint(1, num_shards) + _f_bytes(3, _f_varint(1, 1))        # endianness LITTLE (0) is the default: omitted


def parse_entry(b: bytes) -> Dict[str, object]:
    e: Dict[str, object] = {"dtype": None, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0}
    for f, wire, v in _decode(b):
        if f == 1:
            e["dtype"] = NAME_OF_DT.get(v, "dt%d" % v)
        elif f == 2:
            dims = []
            for ff, _, dim in _decode(v):
                if ff == 2:
                    size = 0
                    for f3, _, x in _decode(dim):
                        if f3 == 1:
                            size = x if x < (1 << 63) else x - (1 << 64)
                    dims.append(size)
            e["shape"] = dims
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = struct.unpack("<I", v)[0]
    return e


# ------------------------------------------------------------------------------------------------------------------
# table writer
# ------------------------------------------------------------------------------------------------------------------
def _block(entries: Sequence[Tuple[bytes, bytes]]) -> bytes:
    body, restarts = bytearray(), []
    for key, value in entries:
        restarts.append(len(body))
        body += _varint(0) + _varint(len(key)) + _varint(len(value)) + key + value
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


def _with_trailer(block: bytes) -> bytes:
    return block + b"\x00" + struct.pack("<I", _py_masked_crc32c(block + b"\x00"))


def _handle(offset: int, size: int) -> bytes:
    return _varint(offset) + _varint(size)


def write_index(path: str, entries: Dict[str, bytes], num_shards: int = 1) -> None:
    """``entries``: tensor name -> serialized ``BundleEntryProto``.  Writes the whole ``.index`` table atomically."""
    import os
    items: List[Tuple[bytes, bytes]] = [(b"", header_proto(num_shards))]
    items += sorted((k.encode("utf-8"), v) for k, v in entries.items())
    data = _block(items)
    out = bytearray(_with_trailer(data))
    meta_off = len(out)
    meta = _block([])
    out += _with_trailer(meta)
    index_off = len(out)
    index = _block([(items[-1][0], _handle(0, len(data)))])          # one data block: separator = its last key
    out += _with_trailer(index)
    footer = _handle(meta_off, len(meta)) + _handle(index_off, len(index))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out += footer
    tmp = "%s.tmp%d" % (path, os.getpid())
    with open(tmp, "wb") as f:
        f.write(out)
    os.replace(tmp, path)


# ------------------------------------------------------------------------------------------------------------------
# table reader
# ------------------------------------------------------------------------------------------------------------------
def _read_varint(buf: bytes, i: int) -> Tuple[int, int]:
    v = shift = 0
    while True:
        b = buf[i]
        i += 1
        v |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return v, i


def _read_block(buf: bytes, offset: int, size: int, what: str) -> bytes:
    block, trailer = buf[offset:offset + size], buf[offset + size:offset + size + 5]
    if len(block) != size or len(trailer) != 5:
        raise ValueError("%s: truncated table block" % what)
    if trailer[0] != 0:
        raise ValueError("%s: compressed table blocks (type %d) are not supported" % (what, trailer[0]))
    if struct.unpack("<I", trailer[1:])[0] != _py_masked_crc32c(block + trailer[:1]):
        raise ValueError("%s: table block checksum mismatch" % what)
    return block


def _iter_block(block: bytes) -> Iterator[Tuple[bytes, bytes]]:
    (num_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * num_restarts
    i, key = 0, b""
    while i < end:
        shared, i = _read_varint(block, i)
        non_shared, i = _read_varint(block, i)
        vlen, i = _read_varint(block, i)
        key = key[:shared] + block[i:i + non_shared]
        i += non_shared
        yield key, block[i:i + vlen]
        i += vlen


def read_index(path: str) -> Tuple[Dict[str, object], Dict[str, Dict[str, object]]]:
    """-> (header fields, {tensor name: entry dict}) of a ``.index`` table (uncompressed blocks)."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != TABLE_MAGIC:
        raise ValueError("%s is not a tensor-bundle index (bad table magic)" % path)
    footer = buf[-48:]
    _, i = _read_varint(footer, 0)
    _, i = _read_varint(footer, i)
    index_off, i = _read_varint(footer, i)
    index_size, i = _read_varint(footer, i)
    header: Dict[str, object] = {"num_shards": 1, "endianness": 0}
    entries: Dict[str, Dict[str, object]] = {}
    for _, handle in _iter_block(_read_block(buf, index_off, index_size, path)):
        off, j = _read_varint(handle, 0)
        size, _ = _read_varint(handle, j)
        for key, value in _iter_block(_read_block(buf, off, size, path)):
            if key == b"":
                for f, _, v in _decode(value):
                    if f == 1:
                        header["num_shards"] = v
                    elif f == 2:
                        header["endianness"] = v
            else:
                entries[key.decode("utf-8")] = parse_entry(value)
    return header, entries
