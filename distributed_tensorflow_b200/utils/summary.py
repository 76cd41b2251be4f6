"""TensorBoard event files: graph dump + scalar summaries.

``summary.FileWriter("logs/", sess.graph)`` (reference ``example_in_graph.py:62``, ``example_distributed_client.py:41``)
writes a real ``events.out.tfevents.<time>.<host>`` file: TFRecord framing (length + masked CRC32C + payload + masked
CRC32C) around ``tensorflow.Event`` protocol buffers, so ``tensorboard --logdir logs/`` shows the graph (one
``NodeDef`` per node with its op, inputs incl. ``^control`` edges, device and ``_output_shapes`` / ``T`` attributes)
and the scalars written by ``add_scalar`` / ``SummarySaverHook`` / ``StepCounterHook``.  TensorFlow's ``.proto``
files are not available here, so the handful of messages needed -- Event, Summary, GraphDef, NodeDef, AttrValue,
TensorShapeProto, TaggedRunMetadata -- are encoded (and, for :func:`read_events`, decoded) directly at the protobuf
wire level; ``tests/test_summary_events.py`` cross-checks the bytes against the real protobuf runtime.
"""
from __future__ import annotations

import json
import os
import socket
import struct

import numpy as np
import threading
import time
from typing import Any, Dict, Iterator, List, Optional, Tuple

from ..framework.graph import GraphKeys, get_default_graph

__all__ = ["FileWriter", "scalar", "merge_all", "read_events", "crc32c", "masked_crc32c"]

# ------------------------------------------------------------------------------------------------------------------
# CRC32C (Castagnoli) + TFRecord framing
# ------------------------------------------------------------------------------------------------------------------
_CRC_TABLE: List[int] = []


def _crc_table() -> List[int]:
    if not _CRC_TABLE:
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            _CRC_TABLE.append(c)
    return _CRC_TABLE


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli).  Small inputs: the byte loop.  Large inputs (checkpoint tensors when the native runtime library
    is missing): the CRC register update is linear over GF(2), so the buffer is cut into equal chunks whose registers are
    advanced TOGETHER, one numpy table lookup per byte position for all chunks at once, and then folded left to right
    with the "append L zero bytes" operator -- ~0.3 s for a 45 MB tensor instead of tens of seconds."""
    t = _crc_table()
    n = len(data)
    if n < (1 << 15):
        c = 0xFFFFFFFF
        for b in data:
            c = t[(c ^ b) & 0xFF] ^ (c >> 8)
        return c ^ 0xFFFFFFFF
    import numpy as np
    L = 4096
    K = n // L
    tab = np.asarray(t, dtype=np.uint32)
    body = np.frombuffer(data, dtype=np.uint8, count=K * L).reshape(K, L).T.copy()       # [L, K]: one row per byte position
    state = np.zeros(K, dtype=np.uint32)
    state[0] = 0xFFFFFFFF
    # the zero-byte operator rides along: 4 x 256 basis registers (one set byte each) advanced by the same L steps
    basis = (np.arange(256, dtype=np.uint32)[None, :] << (8 * np.arange(4, dtype=np.uint32))[:, None]).reshape(-1)
    for j in range(L):
        state = tab[(state ^ body[j]) & 0xFF] ^ (state >> 8)
        basis = tab[basis & 0xFF] ^ (basis >> 8)
    T0, T1, T2, T3 = (basis[i * 256:(i + 1) * 256].tolist() for i in range(4))
    regs = state.tolist()
    r = regs[0]
    for i in range(1, K):
        r = T0[r & 0xFF] ^ T1[(r >> 8) & 0xFF] ^ T2[(r >> 16) & 0xFF] ^ T3[r >> 24] ^ regs[i]
    for b in bytes(data[K * L:]):
        r = t[(r ^ b) & 0xFF] ^ (r >> 8)
    return r ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _tfrecord(payload: bytes) -> bytes:
    head = struct.pack("<Q", len(payload))
    return head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload))


def _read_tfrecords(path: str) -> Iterator[bytes]:
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if len(head) < 8:
                return
            (n,) = struct.unpack("<Q", head)
            (hc,) = struct.unpack("<I", f.read(4))
            if hc != masked_crc32c(head):
                raise ValueError("%s: corrupt record header" % path)
            payload = f.read(n)
            (pc,) = struct.unpack("<I", f.read(4))
            if len(payload) != n or pc != masked_crc32c(payload):
                raise ValueError("%s: corrupt record payload" % path)
            yield payload


# ------------------------------------------------------------------------------------------------------------------
# protobuf wire encoding (only what the Event family needs)
# ------------------------------------------------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode("utf-8"))


def _f_double(field: int, v: float) -> bytes:
    return _key(field, 1) + struct.pack("<d", float(v))


def _f_float(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", float(v))


_DT = {"torch.float32": 1, "torch.float64": 2, "torch.int32": 3, "torch.uint8": 4, "torch.int16": 5, "torch.int8": 6,
       "torch.int64": 9, "torch.bool": 10, "torch.bfloat16": 14, "torch.float16": 19}       # tensorflow.DataType values


def _shape_proto(shape) -> bytes:
    out = b""
    for d in shape:
        out += _f_bytes(2, _f_varint(1, -1 if d is None else int(d)))           # TensorShapeProto.dim{size}
    return out


def _attr_entry(key: str, attr_value: bytes) -> bytes:
    return _f_bytes(5, _f_str(1, key) + _f_bytes(2, attr_value))                 # NodeDef.attr map entry


def _node_def(n: Dict[str, Any]) -> bytes:
    out = _f_str(1, n["name"]) + _f_str(2, n["op"])
    for i in n.get("input", []):
        out += _f_str(3, i)
    if n.get("device"):
        out += _f_str(4, n["device"])
    dt = _DT.get(n.get("dtype") or "")
    if dt is not None:
        out += _attr_entry("T", _f_varint(6, dt))                                # AttrValue.type
    if n.get("shape") is not None:
        out += _attr_entry("_output_shapes", _f_bytes(1, _f_bytes(7, _shape_proto(n["shape"]))))   # AttrValue.list{shape}
    return out


def _graph_def(gd: Dict[str, Any]) -> bytes:
    out = b"".join(_f_bytes(1, _node_def(n)) for n in gd["node"])
    return out + _f_bytes(4, _f_varint(1, 27))                                   # versions { producer: 27 } (TF 1.12)


def _event(wall_time: float, step: Optional[int] = None, **what: bytes) -> bytes:
    out = _f_double(1, wall_time)
    if step is not None:
        out += _f_varint(2, step)
    if "file_version" in what:
        out += _f_bytes(3, what["file_version"])
    if "graph_def" in what:
        out += _f_bytes(4, what["graph_def"])
    if "summary" in what:
        out += _f_bytes(5, what["summary"])
    if "tagged_run_metadata" in what:
        out += _f_bytes(8, what["tagged_run_metadata"])
    return out


# ------------------------------------------------------------------------------------------------------------------
# generic wire decoder (read_events)
# ------------------------------------------------------------------------------------------------------------------
def _decode(buf: bytes) -> List[Tuple[int, int, Any]]:
    """[(field, wire_type, value)]: varint -> int, fixed64 / fixed32 -> raw bytes, length-delimited -> bytes."""
    out, i, n = [], 0, len(buf)
    while i < n:
        key = shift = 0
        while True:
            b = buf[i]
            i += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        field, wire = key >> 3, key & 7
        if wire == 0:
            v = shift = 0
            while True:
                b = buf[i]
                i += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            out.append((field, wire, v))
        elif wire == 1:
            out.append((field, wire, buf[i:i + 8]))
            i += 8
        elif wire == 5:
            out.append((field, wire, buf[i:i + 4]))
            i += 4
        elif wire == 2:
            ln = shift = 0
            while True:
                b = buf[i]
                i += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            out.append((field, wire, buf[i:i + ln]))
            i += ln
        else:
            raise ValueError("unsupported wire type %d" % wire)
    return out


def _decode_node(b: bytes) -> Dict[str, Any]:
    n: Dict[str, Any] = {"input": [], "device": ""}
    for f, _, v in _decode(b):
        if f == 1:
            n["name"] = v.decode()
        elif f == 2:
            n["op"] = v.decode()
        elif f == 3:
            n["input"].append(v.decode())
        elif f == 4:
            n["device"] = v.decode()
    return n


def _decode_event(b: bytes) -> Dict[str, Any]:
    rec: Dict[str, Any] = {}
    for f, _, v in _decode(b):
        if f == 1:
            rec["wall_time"] = struct.unpack("<d", v)[0]
        elif f == 2:
            rec["step"] = v if v < (1 << 63) else v - (1 << 64)
        elif f == 3:
            rec["file_version"] = v.decode()
        elif f == 4:
            rec["graph_def"] = {"node": [_decode_node(x) for ff, _, x in _decode(v) if ff == 1]}
        elif f == 5:
            for ff, _, val in _decode(v):
                if ff == 1:
                    tag, simple = "", None
                    for f3, _, x in _decode(val):
                        if f3 == 1:
                            tag = x.decode()
                        elif f3 == 2:
                            simple = struct.unpack("<f", x)[0]
                        elif f3 == 5:
                            h: Dict[str, Any] = {}
                            for f4, _, y in _decode(x):
                                if f4 in (1, 2, 3, 4, 5):
                                    h[{1: "min", 2: "max", 3: "num", 4: "sum", 5: "sum_squares"}[f4]] = struct.unpack("<d", y)[0]
                                elif f4 in (6, 7):
                                    h["bucket_limit" if f4 == 6 else "bucket"] = list(struct.unpack("<%dd" % (len(y) // 8), y))
                            rec.setdefault("histograms", []).append({"tag": tag, "histo": h})
                            simple = "histogram"
                    if simple == "histogram":
                        continue
                    rec.setdefault("scalars", []).append({"tag": tag, "value": simple})
                    rec["scalar"] = {"tag": tag, "value": simple}
        elif f == 8:
            tag, payload = "", b""
            for ff, _, x in _decode(v):
                if ff == 1:
                    tag = x.decode()
                elif ff == 2:
                    payload = x
            rec["run_metadata"] = {"tag": tag, "step_stats": json.loads(payload.decode()) if payload else []}
    return rec


# ------------------------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------------------------
class FileWriter:
    def __init__(self, logdir: str, graph=None, max_queue: int = 10, flush_secs: float = 120, filename_suffix: str = ""):
        os.makedirs(logdir, exist_ok=True)
        self._path = os.path.join(logdir, "events.out.tfevents.%010d.%s%s" % (time.time(), socket.gethostname(),
                                                                            filename_suffix))
        self._f = open(self._path, "ab")
        self._lock = threading.Lock()
        self._write(_event(time.time(), file_version=b"brain.Event:2"))
        if graph is not None:
            self.add_graph(graph)

    @property
    def path(self) -> str:
        return self._path

    def get_logdir(self) -> str:
        return os.path.dirname(self._path)

    def _write(self, event: bytes) -> None:
        with self._lock:
            self._f.write(_tfrecord(event))

    def add_graph(self, graph, global_step: Optional[int] = None) -> None:
        self._write(_event(time.time(), global_step, graph_def=_graph_def(graph.as_graph_def())))
        self.flush()

    def add_scalar(self, tag: str, value: float, global_step: Optional[int] = None) -> None:
        self._write(_event(time.time(), global_step, summary=_f_bytes(1, _f_str(1, tag) + _f_float(2, value))))

    def add_summary(self, summary: Dict[str, Any], global_step: Optional[int] = None) -> None:
        """``summary``: what ``sess.run(tf.summary.merge_all())`` returned -- tag -> scalar (``simple_value``) or tag -> array
        (a ``tf.summary.histogram``: written as a ``HistogramProto`` over TensorFlow's default bucket limits)."""
        values = b""
        for k, v in summary.items():
            arr = np.asarray(v)
            if arr.ndim == 0:
                values += _f_bytes(1, _f_str(1, k) + _f_float(2, float(arr)))
            else:
                values += _f_bytes(1, _f_str(1, k) + _f_bytes(5, _histogram_proto(arr)))
        self._write(_event(time.time(), global_step, summary=values))

    def add_histogram(self, tag: str, values, global_step: Optional[int] = None) -> None:
        self.add_summary({tag: np.asarray(values).reshape(-1)}, global_step)

    def add_run_metadata(self, run_metadata, tag: str, global_step: Optional[int] = None) -> None:
        """The step trace of one ``Session.run`` (``RunOptions.FULL_TRACE``).  TF stores a serialized ``RunMetadata``
        proto here; ours carries the timeline events as JSON bytes in the same ``TaggedRunMetadata`` envelope."""
        payload = json.dumps(list(run_metadata.step_stats)).encode()
        self._write(_event(time.time(), global_step, tagged_run_metadata=_f_str(1, tag) + _f_bytes(2, payload)))

    def flush(self) -> None:
        with self._lock:
            self._f.flush()

    def close(self) -> None:
        with self._lock:
            if not self._f.closed:
                self._f.flush()
                self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
        return False


class _ScalarSummary:
    def __init__(self, tag: str, tensor):
        self.tag, self.tensor, self.name = tag, tensor, tag


def scalar(name: str, tensor, collections=None) -> _ScalarSummary:
    s = _ScalarSummary(name, tensor)
    get_default_graph().add_to_collection(GraphKeys.SUMMARIES, s)
    return s


def histogram(name: str, values, collections=None) -> _ScalarSummary:
    """``tf.summary.histogram``: the fetched tensor is written as a HistogramProto by ``FileWriter.add_summary``."""
    from ..framework.ops import reshape
    s = _ScalarSummary(name, reshape(values, [-1], name=name + "/values"))
    get_default_graph().add_to_collection(GraphKeys.SUMMARIES, s)
    return s


def merge(inputs, collections=None, name=None) -> Dict[str, Any]:
    out: Dict[str, Any] = {}
    for s in inputs:
        out.update(s if isinstance(s, dict) else {s.tag: s.tensor})
    return out


def merge_all() -> Dict[str, Any]:
    return {s.tag: s.tensor for s in get_default_graph().get_collection(GraphKeys.SUMMARIES)}


def _default_bucket_limits() -> List[float]:
    """TensorFlow's histogram buckets: +-1e-12 * 1.1^k up to 1e20, mirrored around zero, closed by DBL_MAX."""
    pos, v = [], 1e-12
    while v < 1e20:
        pos.append(v)
        v *= 1.1
    return [-x for x in reversed(pos)] + [0.0] + pos + [1.7976931348623157e308]


_BUCKET_LIMITS: List[float] = []


def _histogram_proto(values) -> bytes:
    global _BUCKET_LIMITS
    if not _BUCKET_LIMITS:
        _BUCKET_LIMITS = _default_bucket_limits()
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    limits = np.asarray(_BUCKET_LIMITS)
    counts = np.zeros(len(limits))
    if v.size:
        idx = np.searchsorted(limits, v, side="left")            # bucket i holds (limit[i-1], limit[i]]
        np.add.at(counts, np.minimum(idx, len(limits) - 1), 1.0)
    nz = np.nonzero(counts)[0]
    lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)  # TF drops the empty buckets at both ends
    out = _f_double(1, float(v.min()) if v.size else 0.0) + _f_double(2, float(v.max()) if v.size else 0.0)
    out += _f_double(3, float(v.size)) + _f_double(4, float(v.sum())) + _f_double(5, float((v * v).sum()))
    out += _f_bytes(6, struct.pack("<%dd" % (hi - lo), *limits[lo:hi])) + _f_bytes(7, struct.pack("<%dd" % (hi - lo), *counts[lo:hi]))
    return out


def read_events(path: str) -> List[Dict[str, Any]]:
    """Decode an event file written by :class:`FileWriter` (or by TensorFlow: the fields decoded are the standard ones)."""
    return [_decode_event(p) for p in _read_tfrecords(path)]
