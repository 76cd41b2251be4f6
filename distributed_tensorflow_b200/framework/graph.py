"""Deferred-execution graph: nodes, name scopes, collections, default-graph stack.

The reference programs are written against TF-1.x graph mode: build symbolic
tensors (``placeholder``, ``Variable``, ``matmul`` ...), then evaluate fetches
with ``Session.run(fetches, feed_dict)`` (reference ``distributed_mnist.py:96-152``,
``example_in_graph.py:32-59``).  This module provides that programming model
with a deliberately small core: every node has exactly one output, carries its
resolved device string, and is evaluated by the executor in
``framework/executor.py`` with PyTorch tensors (sm_100a kernels on CUDA
devices, see ``ops/``).  Nothing here is a translation of TF's C++ graph code.
"""
from __future__ import annotations

import contextlib
import threading
from collections import defaultdict
from typing import Any, Dict, Iterable, List, Optional, Sequence, Tuple

from . import device as _device

__all__ = ["Graph", "Tensor", "get_default_graph", "reset_default_graph", "name_scope", "GraphKeys",
           "add_to_collection", "get_collection", "convert_to_tensor", "control_dependencies"]


class GraphKeys:
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"
    LOCAL_VARIABLES = "local_variables"
    GLOBAL_STEP = "global_step"
    LOSSES = "losses"
    SUMMARIES = "summaries"
    INIT_OP = "init_op"
    QUEUE_RUNNERS = "queue_runners"
    UPDATE_OPS = "update_ops"
    REGULARIZATION_LOSSES = "regularization_losses"


class Tensor:
    """A graph node with a single output.  ``op`` returns self (TF exposes both)."""

    __slots__ = ("graph", "id", "name", "op_type", "inputs", "control_inputs", "attrs", "device",
                 "dtype", "shape", "__weakref__")

    def __init__(self, graph: "Graph", op_type: str, inputs: Sequence["Tensor"], attrs: Dict[str, Any],
                 name: str, dtype=None, shape=None):
        self.graph = graph
        self.op_type = op_type
        self.inputs = list(inputs)
        self.control_inputs: List[Tensor] = list(graph._control_stack_flat())
        self.attrs = attrs
        self.name = name
        self.dtype = dtype
        self.shape = tuple(shape) if shape is not None else None
        self.device = ""
        self.id = -1

    # -- TF-style accessors -------------------------------------------------
    @property
    def op(self) -> "Tensor":
        return self

    @property
    def type(self) -> str:
        return self.op_type

    def get_shape(self):
        from .shapes import TensorShape
        return TensorShape(self.shape)

    def set_shape(self, shape) -> None:
        self.shape = tuple(shape)

    def eval(self, feed_dict=None, session=None):
        from ..client.session import get_default_session
        sess = session or get_default_session()
        if sess is None:
            raise ValueError("no default session; use `with Session():` or pass session=")
        return sess.run(self, feed_dict=feed_dict)

    def run(self, feed_dict=None, session=None):
        return self.eval(feed_dict, session)

    def __repr__(self) -> str:
        return "<dtf.Tensor %r op=%s shape=%s device=%r>" % (self.name, self.op_type, self.shape, self.device)

    __hash__ = object.__hash__

    def __bool__(self):
        raise TypeError("a symbolic dtf.Tensor has no truth value; evaluate it with Session.run")

    def __iter__(self):
        raise TypeError("a symbolic dtf.Tensor is not iterable")

    # arithmetic operators are attached in framework/ops.py (avoids an import cycle)


class Graph:
    def __init__(self) -> None:
        self.nodes: List[Tensor] = []
        self._names: Dict[str, int] = {}
        self._by_name: Dict[str, Tensor] = {}
        self.collections: Dict[str, List[Any]] = defaultdict(list)
        self._name_stack: List[str] = []
        self._control_stack: List[List[Tensor]] = []
        self._lock = threading.RLock()
        self.variables: Dict[str, Any] = {}          # full name -> Variable
        self._var_scope = None                        # managed by framework/variables.py
        self.seed: Optional[int] = None
        self.version = 0

    # -- naming -------------------------------------------------------------
    def unique_name(self, name: str, mark_as_used: bool = True) -> str:
        scope = "/".join(self._name_stack)
        full = "%s/%s" % (scope, name) if scope else name
        n = self._names.get(full)
        if n is None:
            if mark_as_used:
                self._names[full] = 1
            return full
        while True:
            cand = "%s_%d" % (full, n)
            n += 1
            if cand not in self._names:
                if mark_as_used:
                    self._names[full] = n
                    self._names[cand] = 1
                return cand

    @contextlib.contextmanager
    def name_scope(self, name: Optional[str]):
        if not name:
            saved = self._name_stack
            self._name_stack = []
            try:
                yield ""
            finally:
                self._name_stack = saved
            return
        if name.endswith("/"):
            # re-enter an absolute scope
            saved = self._name_stack
            self._name_stack = [p for p in name.split("/") if p]
            try:
                yield name
            finally:
                self._name_stack = saved
            return
        scope = "/".join(self._name_stack)
        full = "%s/%s" % (scope, name) if scope else name
        n = self._names.get(full)
        if n is None:
            self._names[full] = 1
            leaf = name
        else:
            while True:
                leaf = "%s_%d" % (name, n)
                cand = "%s/%s" % (scope, leaf) if scope else leaf
                n += 1
                if cand not in self._names:
                    self._names[full] = n
                    self._names[cand] = 1
                    break
        self._name_stack.append(leaf)
        try:
            yield "/".join(self._name_stack) + "/"
        finally:
            self._name_stack.pop()

    # -- node creation --------------------------------------------------------
    def create_node(self, op_type: str, inputs: Sequence[Tensor] = (), attrs: Optional[Dict[str, Any]] = None,
                    name: Optional[str] = None, dtype=None, shape=None, device: Optional[str] = None,
                    exact_name: bool = False) -> Tensor:
        with self._lock:
            base = name or op_type
            full = base if exact_name else self.unique_name(base)
            if exact_name:
                self._names.setdefault(full, 1)
            node = Tensor(self, op_type, inputs, attrs or {}, full, dtype, shape)
            node.device = device if device is not None else _device.apply_device_stack(node)
            node.id = len(self.nodes)
            self.nodes.append(node)
            self._by_name[full] = node
            self.version += 1
            return node

    def get_tensor_by_name(self, name: str) -> Tensor:
        key = name[:-2] if name.endswith(":0") else name
        try:
            return self._by_name[key]
        except KeyError:
            raise KeyError("no tensor named %r in the graph" % name) from None

    get_operation_by_name = get_tensor_by_name

    def get_operations(self) -> List[Tensor]:
        return list(self.nodes)

    # -- control dependencies ---------------------------------------------------
    def _control_stack_flat(self) -> List[Tensor]:
        out: List[Tensor] = []
        for level in self._control_stack:
            out.extend(level)
        return out

    @contextlib.contextmanager
    def control_dependencies(self, control_inputs: Optional[Iterable[Tensor]]):
        if control_inputs is None:
            saved = self._control_stack
            self._control_stack = []
            try:
                yield
            finally:
                self._control_stack = saved
            return
        deps = []
        for c in control_inputs:
            deps.append(c._node if hasattr(c, "_node") else c)
        self._control_stack.append(deps)
        try:
            yield
        finally:
            self._control_stack.pop()

    # -- collections ------------------------------------------------------------
    def add_to_collection(self, name: str, value: Any) -> None:
        self.collections[name].append(value)

    def get_collection(self, name: str, scope: Optional[str] = None) -> List[Any]:
        items = list(self.collections.get(name, ()))
        if scope:
            items = [v for v in items if getattr(v, "name", "").startswith(scope)]
        return items

    def get_collection_ref(self, name: str) -> List[Any]:
        return self.collections[name]

    # -- default-graph handling ---------------------------------------------------
    @contextlib.contextmanager
    def as_default(self):
        _graph_stack().append(self)
        try:
            yield self
        finally:
            _graph_stack().pop()

    def device(self, spec):
        return _device.device(spec)

    # -- export (summary.FileWriter / timeline use this) ----------------------------
    def as_graph_def(self) -> Dict[str, Any]:
        nodes = []
        for n in self.nodes:
            nodes.append({
                "name": n.name, "op": n.op_type, "device": n.device,
                "input": [i.name for i in n.inputs] + ["^" + c.name for c in n.control_inputs],
                "shape": list(n.shape) if n.shape is not None else None,
                "dtype": str(n.dtype) if n.dtype is not None else None,
            })
        return {"node": nodes, "version": self.version}


_tls = threading.local()
_global_default = Graph()


def _graph_stack() -> List[Graph]:
    st = getattr(_tls, "stack", None)
    if st is None:
        st = _tls.stack = []
    return st


def get_default_graph() -> Graph:
    st = _graph_stack()
    return st[-1] if st else _global_default


def reset_default_graph() -> None:
    global _global_default
    if _graph_stack():
        raise AssertionError("reset_default_graph() inside a `with graph.as_default()` block")
    _global_default = Graph()


def name_scope(name: Optional[str], default_name: Optional[str] = None, values=None):
    return get_default_graph().name_scope(name or default_name)


def control_dependencies(control_inputs):
    return get_default_graph().control_dependencies(control_inputs)


def add_to_collection(name: str, value: Any) -> None:
    get_default_graph().add_to_collection(name, value)


def get_collection(name: str, scope: Optional[str] = None) -> List[Any]:
    return get_default_graph().get_collection(name, scope)


def convert_to_tensor(value, dtype=None, name: Optional[str] = None) -> Tensor:
    """Symbolic tensors pass through; variables read; python/numpy/torch values become Const nodes."""
    if isinstance(value, Tensor):
        return value
    node = getattr(value, "_node", None)        # Variable
    if node is not None:
        return node
    from . import ops as _ops
    return _ops.constant(value, dtype=dtype, name=name or "Const")
