"""Name-keyed checkpoints: ``Saver``, ``get_checkpoint_state``, ``latest_checkpoint``.

On-disk layout = TensorFlow's checkpoint V2 ("tensor bundle"):

* ``<dir>/checkpoint`` -- text state file: ``model_checkpoint_path: "model.ckpt-1200"`` plus one
  ``all_model_checkpoint_paths`` line per retained checkpoint (``max_to_keep=5``);
* ``<prefix>.index`` -- a LevelDB-format table: tensor name -> ``BundleEntryProto`` {dtype, shape, shard, offset,
  size, masked CRC32C}, key ``""`` -> ``BundleHeaderProto`` (``train/tensor_bundle.py``);
* ``<prefix>.data-00000-of-00001`` -- the raw little-endian tensor bytes, 64-byte aligned
  (written by ``csrc/runtime/bundle_io.cpp`` with positional writes when the native runtime
  is built, by Python otherwise; CRC32C by the SSE4.2 instruction);
* ``<prefix>.meta`` -- a ``MetaGraphDef`` proto: the graph (``NodeDef`` per node, as in the event files) + ``saver_def``.

Checkpoints written by earlier versions of this framework (JSON ``.index``) are still readable.

Tensors are keyed **by variable name**, and restore is **partial**: a Saver only
looks up the names of *its* variables, so the predict program can pull
``hid_w, hid_b, sm_w, sm_b`` out of a training checkpoint that also holds
``global_step``, Adam slots and beta powers (reference
``distributed_mnist_predict.py:17-26,36-40``).  The directory must be on a
filesystem every task can reach when ps and workers are on different hosts
(reference ``example_between_graph.py:90``); HDFS URLs of the reference map to
a local path through ``resolve_path``.
"""
from __future__ import annotations

import glob
import json
import os
import re
import time
import zlib
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np
import torch

from ..framework import errors
from ..framework import ops as _ops
from ..framework.graph import get_default_graph
from ..framework.variables import Variable, assign, global_variables
from . import tensor_bundle

__all__ = ["Saver", "CheckpointState", "get_checkpoint_state", "latest_checkpoint", "update_checkpoint_state",
           "checkpoint_exists", "load_checkpoint", "list_variables", "resolve_path", "NewCheckpointReader"]

_DTYPES = {
    "float32": torch.float32, "float64": torch.float64, "float16": torch.float16, "bfloat16": torch.bfloat16,
    "int64": torch.int64, "int32": torch.int32, "int16": torch.int16, "int8": torch.int8, "uint8": torch.uint8,
    "bool": torch.bool,
}
_ALIGN = 64


def resolve_path(path: str) -> str:
    """Map ``hdfs://host:port/a/b`` (reference checkpoint dirs) onto a local directory.

    ``DTF_HDFS_ROOT`` (default ``/tmp/dtf_hdfs``) plays the role of the shared filesystem.
    """
    m = re.match(r"^[a-zA-Z][a-zA-Z0-9+.-]*://[^/]*(/.*)?$", path)
    if m:
        root = os.environ.get("DTF_HDFS_ROOT", "/tmp/dtf_hdfs")
        return os.path.join(root, (m.group(1) or "/").lstrip("/"))
    return path


class CheckpointState:
    def __init__(self, model_checkpoint_path: str, all_model_checkpoint_paths: Optional[List[str]] = None):
        self.model_checkpoint_path = model_checkpoint_path
        self.all_model_checkpoint_paths = list(all_model_checkpoint_paths or [model_checkpoint_path])

    def __repr__(self):
        return "CheckpointState(model_checkpoint_path=%r)" % self.model_checkpoint_path


def _state_file(checkpoint_dir: str, latest_filename: Optional[str] = None) -> str:
    return os.path.join(resolve_path(checkpoint_dir), latest_filename or "checkpoint")


def update_checkpoint_state(save_dir: str, model_checkpoint_path: str,
                            all_model_checkpoint_paths: Optional[Sequence[str]] = None,
                            latest_filename: Optional[str] = None) -> None:
    save_dir = resolve_path(save_dir)
    os.makedirs(save_dir, exist_ok=True)
    rel = lambda p: os.path.relpath(p, save_dir) if os.path.isabs(p) else p
    lines = ['model_checkpoint_path: "%s"' % rel(model_checkpoint_path)]
    for p in (all_model_checkpoint_paths or [model_checkpoint_path]):
        lines.append('all_model_checkpoint_paths: "%s"' % rel(p))
    tmp = _state_file(save_dir, latest_filename) + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        f.write("\n".join(lines) + "\n")
    os.replace(tmp, _state_file(save_dir, latest_filename))


def get_checkpoint_state(checkpoint_dir: str, latest_filename: Optional[str] = None) -> Optional[CheckpointState]:
    path = _state_file(checkpoint_dir, latest_filename)
    if not os.path.exists(path):
        return None
    base = resolve_path(checkpoint_dir)
    model, all_paths = None, []
    with open(path) as f:
        for line in f:
            m = re.match(r'^\s*(model_checkpoint_path|all_model_checkpoint_paths)\s*:\s*"(.*)"\s*$', line)
            if not m:
                continue
            p = m.group(2)
            if not os.path.isabs(p):
                p = os.path.join(base, p)
            if m.group(1) == "model_checkpoint_path":
                model = p
            else:
                all_paths.append(p)
    if model is None:
        return None
    return CheckpointState(model, all_paths or [model])


def checkpoint_exists(prefix: str) -> bool:
    return os.path.exists(resolve_path(prefix) + ".index")


def latest_checkpoint(checkpoint_dir: str, latest_filename: Optional[str] = None) -> Optional[str]:
    st = get_checkpoint_state(checkpoint_dir, latest_filename)
    if st and checkpoint_exists(st.model_checkpoint_path):
        return st.model_checkpoint_path
    return None


# ---------------------------------------------------------------------------
# bundle read / write
# ---------------------------------------------------------------------------
def _dtype_name(dt: torch.dtype) -> str:
    return str(dt).replace("torch.", "")


def _tensor_bytes(t: torch.Tensor) -> bytes:
    t = t.detach().contiguous().cpu()
    if t.dtype == torch.bfloat16:
        return t.view(torch.int16).numpy().tobytes()
    if t.dtype == torch.bool:
        return t.to(torch.uint8).numpy().tobytes()
    return t.numpy().tobytes()


def write_bundle(prefix: str, tensors: Dict[str, torch.Tensor], meta: Optional[Dict[str, Any]] = None) -> None:
    prefix = resolve_path(prefix)
    os.makedirs(os.path.dirname(prefix) or ".", exist_ok=True)
    index: Dict[str, Any] = {}
    offset = 0
    blobs = []
    for name in sorted(tensors):
        t = tensors[name]
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(t)
        b = _tensor_bytes(t)
        index[name] = tensor_bundle.entry_proto(_dtype_name(t.dtype), list(t.shape), offset, len(b),
                                                tensor_bundle.masked_crc32c(b))
        blobs.append((offset, b))
        offset += (len(b) + _ALIGN - 1) // _ALIGN * _ALIGN
    data_path = prefix + ".data-00000-of-00001"
    from ..utils import native_runtime
    lib = native_runtime.load()
    tmp = data_path + ".tmp%d" % os.getpid()
    if lib is not None:
        from ..utils._native_bindings import bundle_write
        bundle_write(lib, tmp, blobs, offset)
    else:
        with open(tmp, "wb") as f:
            for off, b in blobs:
                f.seek(off)
                f.write(b)
            f.truncate(offset)
    os.replace(tmp, data_path)
    tensor_bundle.write_index(prefix + ".index", index)
    if meta is not None:
        with open(prefix + ".meta", "wb") as f:
            f.write(_meta_graph_def(meta))


def _meta_graph_def(meta: Dict[str, Any]) -> bytes:
    """``<prefix>.meta`` as a ``tensorflow.MetaGraphDef``: ``meta_info_def`` {tensorflow_version, tags}, ``graph_def`` (the same
    GraphDef encoding the event files use) and a ``saver_def`` {restore_op_name, max_to_keep, version = V2}."""
    from ..utils.summary import _f_bytes, _f_float, _f_str, _f_varint, _graph_def
    info = _f_str(4, "train") + _f_str(5, "1.12.0-dtf_b200")                 # MetaInfoDef.tags / tensorflow_version
    saver = (_f_str(1, "save/Const:0") + _f_str(2, "save/control_dependency:0") + _f_str(3, "save/restore_all")
             + _f_varint(4, int(meta.get("saver", {}).get("max_to_keep", 5))) + _f_float(6, 10000.0) + _f_varint(7, 2))
    return _f_bytes(1, info) + _f_bytes(2, _graph_def(meta["graph_def"])) + _f_bytes(3, saver)


class CheckpointReader:
    def __init__(self, prefix: str):
        self.prefix = resolve_path(prefix)
        idx = self.prefix + ".index"
        if not os.path.exists(idx):
            raise errors.NotFoundError("checkpoint %r not found (no %s)" % (prefix, idx))
        with open(idx, "rb") as f:
            legacy = f.read(1) == b"{"
        if legacy:                                   # JSON index of earlier versions of this framework
            with open(idx) as f:
                self._index = json.load(f)["tensors"]
        else:
            header, entries = tensor_bundle.read_index(idx)
            if header.get("endianness", 0) != 0:
                raise errors.OpError("checkpoint %s was written big-endian" % self.prefix)
            self._index = {k: {"dtype": e["dtype"], "shape": e["shape"], "offset": e["offset"], "nbytes": e["size"],
                               "crc32c": e["crc32c"], "shard": e["shard_id"]} for k, e in entries.items()}
            self._num_shards = int(header.get("num_shards", 1))
        self._data = self.prefix + ".data-00000-of-00001"

    def has_tensor(self, name: str) -> bool:
        return name in self._index

    def get_variable_to_shape_map(self) -> Dict[str, List[int]]:
        return {k: list(v["shape"]) for k, v in self._index.items()}

    def get_variable_to_dtype_map(self) -> Dict[str, torch.dtype]:
        return {k: _DTYPES[v["dtype"]] for k, v in self._index.items()}

    def get_tensor(self, name: str, verify: bool = True) -> torch.Tensor:
        try:
            e = self._index[name]
        except KeyError:
            raise errors.NotFoundError("Key %s not found in checkpoint %s" % (name, self.prefix)) from None
        data = self._data
        if e.get("shard"):                           # multi-shard bundles (written by TensorFlow with sharded savers)
            data = "%s.data-%05d-of-%05d" % (self.prefix, e["shard"], getattr(self, "_num_shards", 1))
        with open(data, "rb") as f:
            f.seek(e["offset"])
            b = f.read(e["nbytes"])
        if len(b) != e["nbytes"]:
            raise errors.OpError("checkpoint %s is truncated at tensor %s" % (self.prefix, name))
        if verify:
            bad = (tensor_bundle.masked_crc32c(b) != e["crc32c"]) if "crc32c" in e \
                else ((zlib.crc32(b) & 0xFFFFFFFF) != e["crc32"])
            if bad:
                raise errors.OpError("checksum mismatch for tensor %s in %s" % (name, self.prefix))
        dt = _DTYPES[e["dtype"]]
        if dt == torch.bfloat16:
            t = torch.from_numpy(np.frombuffer(b, dtype=np.int16).copy()).view(torch.bfloat16)
        elif dt == torch.bool:
            t = torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()).to(torch.bool)
        else:
            npdt = {"float32": np.float32, "float64": np.float64, "float16": np.float16, "int64": np.int64,
                    "int32": np.int32, "int16": np.int16, "int8": np.int8, "uint8": np.uint8}[e["dtype"]]
            t = torch.from_numpy(np.frombuffer(b, dtype=npdt).copy())
        return t.reshape(e["shape"])


NewCheckpointReader = CheckpointReader
load_checkpoint = CheckpointReader


def list_variables(ckpt_dir_or_file: str):
    path = ckpt_dir_or_file
    if os.path.isdir(resolve_path(path)):
        path = latest_checkpoint(path)
        if path is None:
            raise errors.NotFoundError("no checkpoint in %r" % ckpt_dir_or_file)
    r = CheckpointReader(path)
    return sorted((k, v) for k, v in r.get_variable_to_shape_map().items())


# ---------------------------------------------------------------------------
# Saver
# ---------------------------------------------------------------------------
class Saver:
    def __init__(self, var_list: Union[None, Sequence[Variable], Dict[str, Variable]] = None, max_to_keep: int = 5,
                 keep_checkpoint_every_n_hours: float = 10000.0, allow_empty: bool = False, name: str = "save",
                 sharded: bool = False, write_version=None, save_relative_paths: bool = True, defer_build=False):
        if var_list is None:
            var_list = global_variables()
        if isinstance(var_list, dict):
            self._vars: Dict[str, Variable] = dict(var_list)
        else:
            self._vars = {v.var_name: v for v in var_list}
        if not self._vars and not allow_empty:
            raise ValueError("No variables to save")
        self._max_to_keep = max_to_keep
        self._name = name
        self._last_checkpoints: List[str] = []
        # restore ops: one feedable Assign per variable, placed on the variable's task
        g = get_default_graph()
        self._restore: Dict[str, Any] = {}
        with g.name_scope(name):
            for key, v in self._vars.items():
                ph = _ops.placeholder(v.dtype, v.shape, name="restore_in")
                self._restore[key] = (ph, assign(v, ph, name="restore"))

    @property
    def last_checkpoints(self) -> List[str]:
        return list(self._last_checkpoints)

    def set_last_checkpoints(self, paths: Sequence[str]) -> None:
        self._last_checkpoints = list(paths)

    def recover_last_checkpoints(self, checkpoint_paths: Sequence[str]) -> None:
        self._last_checkpoints = [p for p in checkpoint_paths if checkpoint_exists(p)]

    # -- save --------------------------------------------------------------------------------------------
    def save(self, sess, save_path: str, global_step=None, latest_filename: Optional[str] = None,
             meta_graph_suffix: str = "meta", write_meta_graph: bool = True, write_state: bool = True) -> str:
        if global_step is not None:
            if not isinstance(global_step, (int, np.integer)):
                global_step = int(sess.run(global_step))
            prefix = "%s-%d" % (save_path, int(global_step))
        else:
            prefix = save_path
        raw = getattr(sess, "raw_session", lambda: sess)()
        keys = list(self._vars)
        values = raw.run([self._vars[k] for k in keys]) if keys else []
        tensors = {k: torch.as_tensor(np.asarray(v)) for k, v in zip(keys, values)}
        for k in keys:                               # keep the declared dtype (numpy has no bfloat16)
            want = self._vars[k].dtype
            if want is not None and tensors[k].dtype != want:
                tensors[k] = tensors[k].to(want)
        meta = None
        if write_meta_graph:
            meta = {"graph_def": raw.graph.as_graph_def(), "saver": {"variables": keys, "max_to_keep": self._max_to_keep},
                    "written": time.time()}
        write_bundle(prefix, tensors, meta)
        full = resolve_path(prefix)
        if write_state:
            self._last_checkpoints = [p for p in self._last_checkpoints if p != full] + [full]
            while self._max_to_keep and len(self._last_checkpoints) > self._max_to_keep:
                old = self._last_checkpoints.pop(0)
                for f in glob.glob(old + ".*"):
                    if re.match(re.escape(old) + r"\.(index|meta|data-\d+-of-\d+)$", f):
                        try:
                            os.remove(f)
                        except OSError:
                            pass
            update_checkpoint_state(os.path.dirname(full), full, self._last_checkpoints, latest_filename)
        return full

    # -- restore -------------------------------------------------------------------------------------------
    def restore(self, sess, save_path: Optional[str]) -> None:
        if save_path is None:
            raise ValueError("Can't load save_path when it is None.")
        reader = CheckpointReader(save_path)
        raw = getattr(sess, "raw_session", lambda: sess)()
        feed, ops = {}, []
        for key, (ph, op) in self._restore.items():
            t = reader.get_tensor(key)              # NotFoundError when the checkpoint lacks one of OUR names
            want = self._vars[key].shape
            if want is not None and tuple(t.shape) != tuple(want) and all(d is not None for d in want):
                raise errors.InvalidArgumentError(
                    "shape mismatch restoring %s: checkpoint %s vs variable %s" % (key, tuple(t.shape), want))
            feed[ph] = t
            ops.append(op)
        if ops:
            raw.run(_ops_group_cached(self, ops), feed_dict=feed)

    def export_meta_graph(self, filename: Optional[str] = None):
        meta = {"graph_def": get_default_graph().as_graph_def(), "saver": {"variables": list(self._vars)}}
        if filename:
            with open(resolve_path(filename), "wb") as f:
                f.write(_meta_graph_def(meta))
        return meta


def _ops_group_cached(saver: Saver, ops):
    g = getattr(saver, "_restore_all", None)
    if g is None:
        g = saver._restore_all = _ops.group(*ops, name="restore_all")
    return g
