"""TFRecord files and ``tf.train.Example`` records: the on-disk input format of TF-1.x programs that do not feed numpy arrays
(the reference feeds ``mnist.train.next_batch`` through placeholders, ``/root/reference/distributed_mnist.py:149-152``; this is
the same role for data sets converted to records).

* ``python_io.TFRecordWriter(path)`` / ``python_io.tf_record_iterator(path)``: TensorFlow's record framing -- ``uint64 length,
  masked crc32c(length), payload, masked crc32c(payload)`` -- the framing the event files of ``utils/summary.py`` use (files
  written here are read by TensorFlow and vice versa).
* ``train.Example`` / ``Features`` / ``Feature`` / ``BytesList`` / ``FloatList`` / ``Int64List``: the ``tensorflow.Example`` message in
  protobuf wire format (``SerializeToString`` / ``FromString``; packed and unpacked repeated scalars are both read).
* ``parse_single_example`` / ``parse_example`` with ``FixedLenFeature`` / ``VarLenFeature``: host-side parsing to numpy (the
  function a ``Dataset.map`` takes), ``decode_raw`` for byte strings holding raw tensors.
* ``data.TFRecordDataset(filenames)``: a :class:`~.dataset.Dataset` of serialized records (one ``bytes`` object per element).
"""
from __future__ import annotations

import struct
from typing import Any, Dict, Iterator, List, Optional, Sequence, Union

import numpy as np

from .summary import _decode, _f_bytes, _read_tfrecords, _varint

__all__ = ["TFRecordWriter", "tf_record_iterator", "Example", "Features", "Feature", "BytesList", "FloatList", "Int64List",
           "FixedLenFeature", "VarLenFeature", "parse_single_example", "parse_example", "decode_raw", "TFRecordDataset", "read_all"]


class TFRecordWriter:
    def __init__(self, path: str, options=None):
        if options not in (None, "", 0):
            raise NotImplementedError("compressed TFRecord files are not provided")
        self._f = open(path, "wb")

    def write(self, record: Union[bytes, bytearray, memoryview]) -> None:
        from ..train.tensor_bundle import masked_crc32c            # SSE4.2 when the runtime library is built
        payload = bytes(record)
        head = struct.pack("<Q", len(payload))
        self._f.write(head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload)))

    def flush(self) -> None:
        self._f.flush()

    def close(self) -> None:
        if not self._f.closed:
            self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _scan_native(buf) -> Optional[tuple]:
    """(offsets, lengths, bad) of the records in ``buf`` through ``csrc/runtime/bundle_io.cpp: dtf_tfrecord_scan`` (both checksums
    of every record verified by the SSE4.2 CRC32C, outside the interpreter); None without the native runtime library.  ``bad``: index
    of the first corrupt record, or -1."""
    import ctypes
    from . import native_runtime
    lib = native_runtime.load()
    if lib is None or not hasattr(lib, "dtf_tfrecord_scan"):
        return None
    fn = lib.dtf_tfrecord_scan
    if not getattr(fn, "_declared", False):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        fn.restype = ctypes.c_int64
        fn._declared = True
    arr = np.frombuffer(buf, np.uint8)
    cap = max(1, arr.size // 16)
    off, ln = np.empty(cap, np.int64), np.empty(cap, np.int64)
    n = int(fn(arr.ctypes.data, arr.size, off.ctypes.data, ln.ctypes.data, cap, 1))
    bad = -1
    if n < 0:                        # records before the corrupt one are still good: rescan up to it
        bad = -n - 1
        n = int(fn(arr.ctypes.data, arr.size, off.ctypes.data, ln.ctypes.data, bad, 1)) if bad > 0 else 0
    return off[:n], ln[:n], bad


def tf_record_iterator(path: str, options=None) -> Iterator[bytes]:
    """Every record of the file, checksums verified (a truncated tail ends the iteration; a corrupt record raises
    ``DataLossError`` when the iteration reaches it).  With the native runtime library the file is mapped and scanned by
    ``dtf_tfrecord_scan``; otherwise record by record in Python."""
    import mmap
    import os
    from ..framework import errors
    if os.path.getsize(path) > 0:
        with open(path, "rb") as f:
            mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        try:
            got = _scan_native(mm)
            if got is not None:
                off, ln, bad = got
                for o, n in zip(off.tolist(), ln.tolist()):
                    yield mm[o:o + n]
                if bad >= 0:
                    raise errors.DataLossError("%s: corrupt record %d (checksum mismatch)" % (path, bad))
                return
        finally:
            mm.close()
    try:
        yield from _read_tfrecords(path)
    except ValueError as e:
        raise errors.DataLossError(str(e))


# ---- tensorflow.Example --------------------------------------------------------------------------------------------------------
class BytesList:
    def __init__(self, value: Sequence[bytes] = ()):
        self.value = [v.encode() if isinstance(v, str) else bytes(v) for v in value]

    def _encode(self) -> bytes:
        return b"".join(_f_bytes(1, v) for v in self.value)


class FloatList:
    def __init__(self, value: Sequence[float] = ()):
        self.value = [float(v) for v in np.asarray(value, np.float32).reshape(-1)]

    def _encode(self) -> bytes:          # packed, like every modern writer
        return _f_bytes(1, np.asarray(self.value, "<f4").tobytes()) if self.value else b""


class Int64List:
    def __init__(self, value: Sequence[int] = ()):
        self.value = [int(v) for v in np.asarray(value).reshape(-1)]

    def _encode(self) -> bytes:
        return _f_bytes(1, b"".join(_varint(v) for v in self.value)) if self.value else b""


class Feature:
    def __init__(self, bytes_list: Optional[BytesList] = None, float_list: Optional[FloatList] = None, int64_list: Optional[Int64List] = None):
        if sum(x is not None for x in (bytes_list, float_list, int64_list)) > 1:
            raise ValueError("a Feature holds ONE of bytes_list / float_list / int64_list")
        self.bytes_list, self.float_list, self.int64_list = bytes_list, float_list, int64_list

    @property
    def kind(self) -> Optional[str]:
        return "bytes_list" if self.bytes_list is not None else "float_list" if self.float_list is not None else \
            "int64_list" if self.int64_list is not None else None

    def _encode(self) -> bytes:
        if self.bytes_list is not None:
            return _f_bytes(1, self.bytes_list._encode())
        if self.float_list is not None:
            return _f_bytes(2, self.float_list._encode())
        if self.int64_list is not None:
            return _f_bytes(3, self.int64_list._encode())
        return b""

    @staticmethod
    def _parse(buf: bytes) -> "Feature":
        for f, _, v in _decode(buf):
            if f == 1:
                return Feature(bytes_list=BytesList([x for ff, _, x in _decode(v) if ff == 1]))
            if f == 2:
                vals: List[float] = []
                for ff, wire, x in _decode(v):
                    if ff == 1:
                        vals += list(np.frombuffer(x, "<f4")) if wire == 2 else [struct.unpack("<f", x)[0]]
                return Feature(float_list=FloatList(vals))
            if f == 3:
                ints: List[int] = []
                for ff, wire, x in _decode(v):
                    if ff != 1:
                        continue
                    if wire == 2:                # packed varints
                        ints += [y if y < (1 << 63) else y - (1 << 64) for _, _, y in _decode_varints(x)]
                    else:
                        ints.append(x if x < (1 << 63) else x - (1 << 64))
                return Feature(int64_list=Int64List(ints))
        return Feature()


def _decode_varints(buf: bytes):
    i, n = 0, len(buf)
    while i < n:
        v = shift = 0
        while True:
            b = buf[i]
            i += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        yield (1, 0, v)


class Features:
    def __init__(self, feature: Optional[Dict[str, Feature]] = None):
        self.feature: Dict[str, Feature] = dict(feature or {})

    def _encode(self) -> bytes:          # map<string, Feature> = repeated entry {1: key, 2: value}; keys sorted: deterministic bytes
        return b"".join(_f_bytes(1, _f_bytes(1, k.encode()) + _f_bytes(2, self.feature[k]._encode())) for k in sorted(self.feature))


class Example:
    def __init__(self, features: Optional[Features] = None):
        self.features = features if features is not None else Features()

    def SerializeToString(self) -> bytes:      # noqa: N802 - protobuf's name
        return _f_bytes(1, self.features._encode())

    @staticmethod
    def FromString(buf: bytes) -> "Example":   # noqa: N802
        ex = Example()
        for f, _, v in _decode(bytes(buf)):
            if f != 1:
                continue
            for ff, _, entry in _decode(v):
                if ff != 1:
                    continue
                key, val = None, Feature()
                for fff, _, x in _decode(entry):
                    if fff == 1:
                        key = x.decode()
                    elif fff == 2:
                        val = Feature._parse(x)
                if key is not None:
                    ex.features.feature[key] = val
        return ex

    def ParseFromString(self, buf: bytes) -> None:   # noqa: N802
        self.features = Example.FromString(buf).features


# ---- parsing -------------------------------------------------------------------------------------------------------------------
class FixedLenFeature:
    def __init__(self, shape, dtype, default_value=None):
        self.shape, self.dtype, self.default_value = tuple(int(d) for d in shape), dtype, default_value


class VarLenFeature:
    def __init__(self, dtype):
        self.dtype = dtype


def _np_dtype(dtype):
    import torch
    if dtype in (str, bytes, "string") or getattr(dtype, "__name__", "") == "string":
        return object
    if isinstance(dtype, torch.dtype):
        return {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32,
                torch.uint8: np.uint8, torch.bool: np.bool_}[dtype]
    return np.dtype(dtype)


def _values(feat: Feature, want) -> list:
    if feat.kind is None:
        return []
    vals = getattr(feat, feat.kind).value
    kind_ok = {"bytes_list": want is object, "float_list": want is not object and np.issubdtype(want, np.floating),
               "int64_list": want is not object and (np.issubdtype(want, np.integer) or want == np.bool_)}[feat.kind]
    if not kind_ok:
        raise ValueError("feature holds a %s but %s was asked for" % (feat.kind, want))
    return vals


def parse_single_example(serialized, features: Dict[str, Any], name=None) -> Dict[str, np.ndarray]:
    """One serialized ``Example`` -> ``{key: numpy array}``: ``FixedLenFeature`` values reshaped to their shape (the default when
    the key is absent, an error without one), ``VarLenFeature`` values as a 1-D array of whatever length the record holds."""
    if isinstance(serialized, np.ndarray):
        serialized = serialized.item() if serialized.shape == () else serialized.tobytes()
    ex = Example.FromString(serialized)
    out: Dict[str, np.ndarray] = {}
    for key, spec in features.items():
        want = _np_dtype(spec.dtype)
        feat = ex.features.feature.get(key)
        if isinstance(spec, VarLenFeature):
            out[key] = np.asarray(_values(feat, want) if feat is not None else [], dtype=want)
            continue
        n = int(np.prod(spec.shape)) if spec.shape else 1
        if feat is None or feat.kind is None:
            if spec.default_value is None:
                raise ValueError("Example has no feature %r and the FixedLenFeature has no default_value" % key)
            out[key] = np.broadcast_to(np.asarray(spec.default_value, dtype=want), spec.shape).copy()
            continue
        vals = _values(feat, want)
        if len(vals) != n:
            raise ValueError("feature %r holds %d values, FixedLenFeature%s needs %d" % (key, len(vals), spec.shape, n))
        out[key] = np.asarray(vals, dtype=want).reshape(spec.shape)
    return out


_KIND = {"bytes": 0, "float": 1, "int64": 2}


def _parse_examples_native(records: List[bytes], features: Dict[str, Any]) -> Optional[Dict[str, np.ndarray]]:
    """The batch through ``csrc/runtime/example_parser.cpp: dtf_parse_examples`` (one call, outside the interpreter): float / int64
    features land directly in their ``[batch, *shape]`` arrays, a bytes feature comes back as (offset, length) pairs that are
    sliced here.  None when the native library is missing, a feature is not a float32 / int64 / string ``FixedLenFeature``, or the
    parser reports a problem -- the Python path then produces the value or the precise error."""
    import ctypes
    from . import native_runtime
    lib = native_runtime.load()
    if lib is None or not hasattr(lib, "dtf_parse_examples") or not records:
        return None
    names = list(features)
    kinds, counts = [], []
    for k in names:
        spec = features[k]
        if not isinstance(spec, FixedLenFeature):
            return None
        want = _np_dtype(spec.dtype)
        kind = 0 if want is object else 1 if want == np.float32 else 2 if want == np.int64 else None
        if kind is None:
            return None
        kinds.append(kind)
        counts.append(int(np.prod(spec.shape)) if spec.shape else 1)
    fn = lib.dtf_parse_examples
    if not getattr(fn, "_declared", False):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        fn.restype = ctypes.c_int64
        fn._declared = True
    n, nf = len(records), len(names)
    buf = b"".join(records)
    lens = np.fromiter((len(r) for r in records), np.int64, n)
    offs = np.cumsum(lens) - lens
    outs = [np.zeros((n, c * 2), np.int64) if kd == 0 else np.zeros((n, c), np.float32 if kd == 1 else np.int64) for kd, c in zip(kinds, counts)]
    keyb = [k.encode() for k in names]
    keys = (ctypes.c_char_p * nf)(*keyb)
    klen = (ctypes.c_int * nf)(*[len(k) for k in keyb])
    kind_a = (ctypes.c_int * nf)(*kinds)
    cnt_a = (ctypes.c_int64 * nf)(*counts)
    out_a = (ctypes.c_void_p * nf)(*[o.ctypes.data for o in outs])
    present = np.zeros((n, nf), np.uint8)
    ef, ec = ctypes.c_int(0), ctypes.c_int(0)
    data = np.frombuffer(buf, np.uint8)
    rc = fn(data.ctypes.data if data.size else None, offs.ctypes.data, lens.ctypes.data, n, nf, keys, klen, kind_a, cnt_a, out_a,
            present.ctypes.data, ctypes.byref(ef), ctypes.byref(ec))
    if rc != 0:
        return None
    result: Dict[str, np.ndarray] = {}
    for j, k in enumerate(names):
        spec, kd, c = features[k], kinds[j], counts[j]
        missing = present[:, j] == 0
        if missing.any() and spec.default_value is None:
            return None                              # the Python path names the record / feature
        if kd == 0:
            pairs = outs[j].reshape(n, c, 2)
            arr = np.empty((n, c), dtype=object)
            for i in range(n):
                if missing[i]:
                    arr[i, :] = spec.default_value if isinstance(spec.default_value, bytes) else str(spec.default_value).encode()
                else:
                    for q in range(c):
                        o, m = int(pairs[i, q, 0]), int(pairs[i, q, 1])
                        arr[i, q] = buf[o:o + m]
            result[k] = arr.reshape((n,) + spec.shape)
        else:
            arr = outs[j]
            if missing.any():
                arr[missing] = np.broadcast_to(np.asarray(spec.default_value, arr.dtype), spec.shape).reshape(-1)
            result[k] = arr.reshape((n,) + spec.shape)
    return result


def parse_example(serialized: Sequence[bytes], features: Dict[str, Any], name=None) -> Dict[str, np.ndarray]:
    """A batch of serialized Examples -> ``[batch, *shape]`` arrays of its ``FixedLenFeature``s (``VarLenFeature`` is per-record: use
    ``parse_single_example`` before batching).  float32 / int64 / string features take the native batch parser
    (``dtf_parse_examples``); anything else, and every error report, the per-record Python path."""
    if any(isinstance(s, VarLenFeature) for s in features.values()):
        raise NotImplementedError("parse_example with VarLenFeature (sparse batches): parse single examples before batching")
    recs = serialized.reshape(-1).tolist() if isinstance(serialized, np.ndarray) else list(serialized)
    recs = [r.item() if isinstance(r, np.ndarray) else bytes(r) for r in recs]
    got = _parse_examples_native(recs, features)
    if got is not None:
        return got
    rows = [parse_single_example(s, features) for s in recs]
    return {k: np.stack([r[k] for r in rows]) for k in features}


def read_all(filenames, features: Dict[str, Any], chunk: int = 4096) -> Dict[str, np.ndarray]:
    """Every record of the files parsed into ``{key: [n, *shape]}`` arrays (native scanner + native batch parser, ``chunk`` records
    per call): the bulk loader in front of ``utils/input_pipeline.EpochBatcher`` -- a data set that lives in record files becomes
    the in-memory arrays the fabric engine's host-fed loop (``PSTrainEngine.train_loop``) cycles through pinned memory."""
    files = [filenames] if isinstance(filenames, (str, bytes)) else list(filenames)
    parts: List[Dict[str, np.ndarray]] = []
    pending: List[bytes] = []
    for path in files:
        for rec in tf_record_iterator(path.decode() if isinstance(path, bytes) else str(path)):
            pending.append(rec)
            if len(pending) >= chunk:
                parts.append(parse_example(pending, features))
                pending = []
    if pending:
        parts.append(parse_example(pending, features))
    if not parts:
        return {k: np.zeros((0,) + tuple(getattr(s, "shape", ())), dtype=_np_dtype(s.dtype)) for k, s in features.items()}
    return {k: np.concatenate([p[k] for p in parts], 0) for k in features}


def decode_raw(data, out_type, little_endian: bool = True, name=None) -> np.ndarray:
    """The bytes of a string feature reinterpreted as a 1-D array of ``out_type`` (images stored with ``tobytes()``); a batch of
    equally long strings (a list, or the object array ``parse_example`` returns) gives ``[batch, n]``."""
    dt = np.dtype(_np_dtype(out_type))
    if isinstance(data, (list, tuple)) or (isinstance(data, np.ndarray) and data.dtype == object and data.ndim >= 1):
        rows = [bytes(r) for r in (data.reshape(-1).tolist() if isinstance(data, np.ndarray) else data)]
        if len({len(r) for r in rows}) > 1:
            raise ValueError("decode_raw(): the strings of a batch must have the same length")
        flat = np.frombuffer(b"".join(rows), dt.newbyteorder("<" if little_endian else ">")).astype(dt)
        lead = data.shape if isinstance(data, np.ndarray) else (len(rows),)
        return flat.reshape(tuple(lead) + (-1,))
    if isinstance(data, np.ndarray):
        data = data.item() if data.shape == () else data.tobytes()
    return np.frombuffer(bytes(data), dt.newbyteorder("<" if little_endian else ">")).astype(dt)


def TFRecordDataset(filenames, compression_type=None, buffer_size=None, num_parallel_reads=None):      # noqa: N802 - TF's name
    """``tf.data.TFRecordDataset``: the records of the files, in order, one ``bytes`` object (0-d object array) per element."""
    from .dataset import Dataset, _Spec
    if compression_type not in (None, ""):
        raise NotImplementedError("compressed TFRecord files are not provided")
    files = [filenames] if isinstance(filenames, (str, bytes)) else [f for f in filenames]
    files = [f.decode() if isinstance(f, bytes) else str(f) for f in files]

    def make():
        for path in files:
            for rec in tf_record_iterator(path):
                a = np.empty((), dtype=object)
                a[()] = rec
                yield a
    return Dataset(make, _Spec(object, ()))
