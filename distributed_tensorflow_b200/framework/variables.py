"""Variables, variable scopes, initializers and the global step.

Capability parity:
* ``Variable(initial_value, name=..., trainable=...)`` -- reference
  ``distributed_mnist.py:98-104``, ``example_in_graph.py:33-36``;
* ``get_variable(name, shape, initializer=...)`` with ``variable_scope`` /
  ``reuse_variables()`` for weight sharing across towers -- reference
  ``example_between_graph.py:53-56``, ``standalone.py:42,49-58,110,121``;
* ``train.get_or_create_global_step()`` -- reference ``distributed_mnist.py:96``.

A variable is a *named resource* owned by the task its node is placed on.
Two client graphs that create a variable with the same name on the same ps
task share storage; that is what makes between-graph replication work.  On
the B200 fabric the storage is a slice of the ps shard's flat parameter
buffer (``parallel/ps_engine.py``); on the control-plane path it is a tensor
in the task's :class:`ResourceStore`.
"""
from __future__ import annotations

import contextlib
import math
from typing import Any, Callable, List, Optional, Sequence

import torch

from . import device as _device
from . import ops as _ops
from .graph import GraphKeys, Tensor, convert_to_tensor, get_default_graph
from .ops import register_kernel

__all__ = ["Variable", "get_variable", "variable_scope", "get_variable_scope", "global_variables",
           "trainable_variables", "global_variables_initializer", "variables_initializer",
           "get_or_create_global_step", "get_global_step", "create_global_step", "assign", "assign_add",
           "assign_sub", "is_variable_initialized", "report_uninitialized_variables",
           "zeros_initializer", "ones_initializer", "constant_initializer", "random_normal_initializer",
           "truncated_normal_initializer", "random_uniform_initializer", "glorot_uniform_initializer",
           "variance_scaling_initializer", "local_variables_initializer", "initialize_all_variables"]


# ---------------------------------------------------------------------------
# initializers: callables (shape, dtype) -> symbolic tensor
# ---------------------------------------------------------------------------
class _Init:
    def __call__(self, shape, dtype=None):
        raise NotImplementedError


class zeros_initializer(_Init):
    def __init__(self, dtype=torch.float32):
        self.dtype = dtype

    def __call__(self, shape, dtype=None):
        return _ops.zeros(shape, dtype or self.dtype, name="Initializer/zeros")


class ones_initializer(_Init):
    def __init__(self, dtype=torch.float32):
        self.dtype = dtype

    def __call__(self, shape, dtype=None):
        return _ops.ones(shape, dtype or self.dtype, name="Initializer/ones")


class constant_initializer(_Init):
    def __init__(self, value=0, dtype=torch.float32):
        self.value, self.dtype = value, dtype

    def __call__(self, shape, dtype=None):
        return _ops.constant(self.value, dtype=dtype or self.dtype, shape=tuple(shape), name="Initializer/Const")


class random_normal_initializer(_Init):
    def __init__(self, mean=0.0, stddev=1.0, seed=None, dtype=torch.float32):
        self.mean, self.stddev, self.seed, self.dtype = mean, stddev, seed, dtype

    def __call__(self, shape, dtype=None):
        return _ops.random_normal(shape, self.mean, self.stddev, dtype or self.dtype, self.seed,
                                  name="Initializer/random_normal")


class truncated_normal_initializer(_Init):
    def __init__(self, mean=0.0, stddev=1.0, seed=None, dtype=torch.float32):
        self.mean, self.stddev, self.seed, self.dtype = mean, stddev, seed, dtype

    def __call__(self, shape, dtype=None):
        return _ops.truncated_normal(shape, self.mean, self.stddev, dtype or self.dtype, self.seed,
                                     name="Initializer/truncated_normal")


class random_uniform_initializer(_Init):
    def __init__(self, minval=0.0, maxval=1.0, seed=None, dtype=torch.float32):
        self.minval, self.maxval, self.seed, self.dtype = minval, maxval, seed, dtype

    def __call__(self, shape, dtype=None):
        return _ops.random_uniform(shape, self.minval, self.maxval, dtype or self.dtype, self.seed,
                                   name="Initializer/random_uniform")


def _fans(shape):
    shape = list(shape)
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    rf = 1
    for d in shape[:-2]:
        rf *= d
    return shape[-2] * rf, shape[-1] * rf


class variance_scaling_initializer(_Init):
    def __init__(self, scale=2.0, mode="fan_in", distribution="truncated_normal", seed=None, dtype=torch.float32):
        self.scale, self.mode, self.distribution, self.seed, self.dtype = scale, mode, distribution, seed, dtype

    def __call__(self, shape, dtype=None):
        fi, fo = _fans(shape)
        n = {"fan_in": fi, "fan_out": fo, "fan_avg": (fi + fo) / 2.0}[self.mode]
        if self.distribution == "uniform":
            lim = math.sqrt(3.0 * self.scale / max(1.0, n))
            return _ops.random_uniform(shape, -lim, lim, dtype or self.dtype, self.seed, name="Initializer/vs")
        std = math.sqrt(self.scale / max(1.0, n))
        return _ops.truncated_normal(shape, 0.0, std, dtype or self.dtype, self.seed, name="Initializer/vs")


def glorot_uniform_initializer(seed=None, dtype=torch.float32):
    return variance_scaling_initializer(1.0, "fan_avg", "uniform", seed, dtype)


# ---------------------------------------------------------------------------
# Variable
# ---------------------------------------------------------------------------
class Variable:
    """Named, task-resident mutable tensor.

    ``var._node`` is the graph node of op type ``VariableV2``; evaluating it
    reads the current value from the owning task.  Arithmetic on a Variable
    builds ops on that read.
    """

    def __init__(self, initial_value=None, trainable: bool = True, collections=None, name: Optional[str] = None,
                 dtype=None, validate_shape: bool = True, _exact_name: bool = False, shape=None):
        if initial_value is None:
            raise ValueError("initial_value must be specified")
        g = get_default_graph()
        base = name or "Variable"
        # Resolve the variable's device FIRST (device functions see op_type VariableV2), then build the
        # initial value under that device so initialisation runs where the storage lives.
        if callable(initial_value) and not isinstance(initial_value, (Tensor, Variable)):
            init_fn = initial_value
        else:
            init_fn = None
        full_name = base if _exact_name else g.unique_name(base)
        if _exact_name:
            g._names.setdefault(full_name, 1)
        self._node = g.create_node("VariableV2", [], {"var_name": full_name, "trainable": bool(trainable)},
                                   full_name, exact_name=True)
        with _device.device(None), _device.device(self._node.device or None), g.name_scope(None), \
                g.control_dependencies(None):
            with g.name_scope(full_name + "/"):
                init = init_fn() if init_fn is not None else initial_value
                if isinstance(init, Variable):
                    init = init.initialized_value()
                init = convert_to_tensor(init, dtype=_ops.as_dtype(dtype))
                if dtype is not None and init.dtype != _ops.as_dtype(dtype):
                    init = _ops.cast(init, dtype)
                self._initial_value = init
                self._initializer = g.create_node("Assign", [init], {"var_name": full_name, "init": True},
                                                  "Assign", init.dtype, init.shape, device=self._node.device)
        self._node.dtype = self._initial_value.dtype
        self._node.shape = self._initial_value.shape if shape is None else tuple(shape)
        self._node.attrs["dtype"] = self._node.dtype
        self._node.attrs["shape"] = self._node.shape
        self._trainable = bool(trainable)
        self.graph = g
        if collections is None:
            collections = [GraphKeys.GLOBAL_VARIABLES]
        collections = list(collections)
        if trainable and GraphKeys.TRAINABLE_VARIABLES not in collections:
            collections.append(GraphKeys.TRAINABLE_VARIABLES)
        for c in collections:
            g.add_to_collection(c, self)
        g.variables[full_name] = self

    # -- identity ---------------------------------------------------------------
    @property
    def name(self) -> str:
        return self._node.name + ":0"

    @property
    def op(self) -> Tensor:
        return self._node

    @property
    def var_name(self) -> str:
        return self._node.name

    @property
    def device(self) -> str:
        return self._node.device

    @property
    def dtype(self):
        return self._node.dtype

    @property
    def shape(self):
        return self._node.shape

    def get_shape(self):
        return self._node.get_shape()

    @property
    def trainable(self) -> bool:
        return self._trainable

    @property
    def initializer(self) -> Tensor:
        return self._initializer

    @property
    def initial_value(self) -> Tensor:
        return self._initial_value

    def initialized_value(self) -> Tensor:
        return _ops.with_dependencies([self._initializer], self._node, name="initialized_value")

    def value(self) -> Tensor:
        return self._node

    def read_value(self) -> Tensor:
        return _ops.identity(self._node, name="read")

    def eval(self, session=None):
        return self._node.eval(session=session)

    def assign(self, value, use_locking=False, name="Assign"):
        return assign(self, value, name=name)

    def assign_add(self, delta, use_locking=False, name="AssignAdd"):
        return assign_add(self, delta, name=name)

    def assign_sub(self, delta, use_locking=False, name="AssignSub"):
        return assign_sub(self, delta, name=name)

    def load(self, value, session=None):
        from ..client.session import get_default_session
        sess = session or get_default_session()
        sess.run(assign(self, _ops.constant(value, dtype=self.dtype)))

    def __repr__(self) -> str:
        return "<dtf.Variable %r shape=%s dtype=%s device=%r>" % (self.name, self.shape, self.dtype, self.device)

    __hash__ = object.__hash__

    # arithmetic: operate on the read value
    def __add__(self, o): return _ops.add(self._node, o)
    def __radd__(self, o): return _ops.add(o, self._node)
    def __sub__(self, o): return _ops.subtract(self._node, o)
    def __rsub__(self, o): return _ops.subtract(o, self._node)
    def __mul__(self, o): return _ops.multiply(self._node, o)
    def __rmul__(self, o): return _ops.multiply(o, self._node)
    def __truediv__(self, o): return _ops.divide(self._node, o)
    def __rtruediv__(self, o): return _ops.divide(o, self._node)
    def __neg__(self): return _ops.negative(self._node)
    def __matmul__(self, o): return _ops.matmul(self._node, o)
    def __getitem__(self, k): return self._node[k]


# -- variable kernels ----------------------------------------------------------
@register_kernel("VariableV2", stateful=True)
def _k_var(ctx, node):
    return ctx.store.read(node.attrs["var_name"])


def _var_name(ref) -> str:
    if isinstance(ref, Variable):
        return ref.var_name
    if isinstance(ref, Tensor) and ref.op_type == "VariableV2":
        return ref.attrs["var_name"]
    raise TypeError("expected a Variable, got %r" % (ref,))


def _var_device(ref) -> str:
    return ref.device if isinstance(ref, Variable) else ref.device


def assign(ref, value, validate_shape=None, use_locking=None, name="Assign") -> Tensor:
    v = convert_to_tensor(value)
    return get_default_graph().create_node("Assign", [v], {"var_name": _var_name(ref)}, name, v.dtype, v.shape,
                                           device=_var_device(ref))


def assign_add(ref, value, use_locking=None, name="AssignAdd") -> Tensor:
    v = convert_to_tensor(value)
    return get_default_graph().create_node("AssignAdd", [v], {"var_name": _var_name(ref)}, name, v.dtype, None,
                                           device=_var_device(ref))


def assign_sub(ref, value, use_locking=None, name="AssignSub") -> Tensor:
    v = convert_to_tensor(value)
    return get_default_graph().create_node("AssignSub", [v], {"var_name": _var_name(ref)}, name, v.dtype, None,
                                           device=_var_device(ref))


@register_kernel("Assign", stateful=True)
def _k_assign(ctx, node, value):
    return ctx.store.assign(node.attrs["var_name"], value.detach(), device=ctx.torch_device(node))


@register_kernel("AssignAdd", stateful=True)
def _k_assign_add(ctx, node, value):
    return ctx.store.assign_add(node.attrs["var_name"], value.detach())


@register_kernel("AssignSub", stateful=True)
def _k_assign_sub(ctx, node, value):
    return ctx.store.assign_add(node.attrs["var_name"], -value.detach())


def is_variable_initialized(variable, name="IsVariableInitialized") -> Tensor:
    return get_default_graph().create_node("IsVariableInitialized", [], {"var_name": _var_name(variable)}, name,
                                           torch.bool, (), device=_var_device(variable))


@register_kernel("IsVariableInitialized", stateful=True)
def _k_is_init(ctx, node):
    return torch.tensor(ctx.store.is_initialized(node.attrs["var_name"]))


def report_uninitialized_variables(var_list=None, name="report_uninitialized_variables") -> Tensor:
    """Evaluates to a python list of names of uninitialised variables (A14: non-chief wait)."""
    vs = global_variables() if var_list is None else list(var_list)
    g = get_default_graph()
    checks = [is_variable_initialized(v) for v in vs]
    return g.create_node("ReportUninitialized", checks, {"names": [v.var_name for v in vs]}, name, device="")


@register_kernel("ReportUninitialized")
def _k_report(ctx, node, *flags):
    return [n for n, f in zip(node.attrs["names"], flags) if not bool(f)]


# ---------------------------------------------------------------------------
# collections helpers
# ---------------------------------------------------------------------------
def global_variables(scope=None) -> List[Variable]:
    return get_default_graph().get_collection(GraphKeys.GLOBAL_VARIABLES, scope)


all_variables = global_variables


def trainable_variables(scope=None) -> List[Variable]:
    return get_default_graph().get_collection(GraphKeys.TRAINABLE_VARIABLES, scope)


def local_variables(scope=None) -> List[Variable]:
    return get_default_graph().get_collection(GraphKeys.LOCAL_VARIABLES, scope)


def variables_initializer(var_list: Sequence[Variable], name="init") -> Tensor:
    return _ops.group(*[v.initializer for v in var_list], name=name)


def global_variables_initializer() -> Tensor:
    return variables_initializer(global_variables(), name="init")


initialize_all_variables = global_variables_initializer


def local_variables_initializer() -> Tensor:
    return variables_initializer(local_variables(), name="init_local")


# ---------------------------------------------------------------------------
# variable scopes / get_variable
# ---------------------------------------------------------------------------
class VariableScope:
    def __init__(self, name: str, reuse: Optional[bool] = False, initializer=None, parent=None):
        self.name, self.reuse, self.initializer, self.parent = name, reuse, initializer, parent

    def reuse_variables(self) -> None:
        self.reuse = True

    @property
    def original_name_scope(self) -> str:
        return self.name + "/" if self.name else ""


def get_variable_scope() -> VariableScope:
    g = get_default_graph()
    if g._var_scope is None:
        g._var_scope = VariableScope("", reuse=False)
    return g._var_scope


AUTO_REUSE = "auto_reuse"


@contextlib.contextmanager
def variable_scope(name_or_scope, default_name=None, reuse=None, initializer=None):
    g = get_default_graph()
    parent = get_variable_scope()
    if isinstance(name_or_scope, VariableScope):
        full, leaf = name_or_scope.name, name_or_scope.name.split("/")[-1]
        inherit = name_or_scope
    else:
        leaf = name_or_scope or default_name or ""
        full = "%s/%s" % (parent.name, leaf) if parent.name and leaf else (leaf or parent.name)
        inherit = parent
    scope = VariableScope(full, reuse if reuse is not None else (True if parent.reuse is True else inherit.reuse),
                          initializer if initializer is not None else inherit.initializer, parent)
    g._var_scope = scope
    try:
        with g.name_scope(leaf if leaf else None) if leaf else contextlib.nullcontext():
            yield scope
    finally:
        # `reuse_variables()` on the *root* scope persists (standalone.py:110 relies on it)
        g._var_scope = parent


def get_variable(name: str, shape=None, dtype=torch.float32, initializer=None, trainable: bool = True,
                 collections=None, regularizer=None) -> Variable:
    g = get_default_graph()
    scope = get_variable_scope()
    full = "%s/%s" % (scope.name, name) if scope.name else name
    existing = g.variables.get(full)
    if scope.reuse is True:
        if existing is None:
            raise ValueError("Variable %s does not exist, or was not created with get_variable()" % full)
        return existing
    if existing is not None:
        if scope.reuse == AUTO_REUSE:
            return existing
        raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True?" % full)
    init = initializer if initializer is not None else scope.initializer
    dt = _ops.as_dtype(dtype) or torch.float32
    if init is None:
        init = glorot_uniform_initializer() if dt.is_floating_point else zeros_initializer(dt)
    if isinstance(init, type):
        init = init()          # standalone.py:51 passes the class `tf.zeros_initializer`
    if isinstance(init, (Tensor, Variable)) or not callable(init):
        init_value = init
    else:
        if shape is None:
            raise ValueError("shape of a new variable (%s) must be known" % full)
        shp = tuple(int(d) for d in shape)
        init_value = (lambda: init(shp, dt))
    return Variable(init_value, trainable=trainable, collections=collections, name=full, dtype=dt,
                    _exact_name=True)


# ---------------------------------------------------------------------------
# global step
# ---------------------------------------------------------------------------
def get_global_step(graph=None) -> Optional[Variable]:
    g = graph or get_default_graph()
    coll = g.get_collection(GraphKeys.GLOBAL_STEP)
    if coll:
        return coll[0]
    return g.variables.get("global_step")


def create_global_step(graph=None) -> Variable:
    g = graph or get_default_graph()
    if get_global_step(g) is not None:
        raise ValueError('"global_step" already exists.')
    with g.as_default(), g.name_scope(None):
        v = Variable(_ops.constant(0, dtype=torch.int64, name="global_step/Initializer"), trainable=False,
                     name="global_step", collections=[GraphKeys.GLOBAL_VARIABLES, GraphKeys.GLOBAL_STEP],
                     _exact_name=True)
    return v


def get_or_create_global_step(graph=None) -> Variable:
    g = graph or get_default_graph()
    v = get_global_step(g)
    return v if v is not None else create_global_step(g)
