"""``tf.train`` pieces next to the reference's own surface: ``ExponentialMovingAverage``, ``AdadeltaOptimizer``,
``global_step`` / ``load_variable`` / ``init_from_checkpoint`` / ``write_graph`` / ``start_queue_runners``.  Graph tier only
(the fabric engines fuse SGD / Momentum / Adam)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from ..framework import device as _device
from ..framework import ops as _ops
from ..framework.ops import register_kernel
from ..framework.graph import Tensor, convert_to_tensor, get_default_graph
from ..framework.variables import Variable, assign, global_variables, trainable_variables
from .optimizer import Optimizer, _lr_at_run_time, _lr_inputs
from .saver import CheckpointReader, latest_checkpoint, resolve_path

__all__ = ["ExponentialMovingAverage", "AdadeltaOptimizer", "global_step", "load_variable", "init_from_checkpoint", "write_graph",
           "start_queue_runners", "add_queue_runner"]


class AdadeltaOptimizer(Optimizer):
    """TF's Adadelta: ``accum = rho*accum + (1-rho)*g^2 ; update = sqrt(accum_update + eps) / sqrt(accum + eps) * g ;
    accum_update = rho*accum_update + (1-rho)*update^2 ; var -= lr * update`` (both slots start at zero)."""

    def __init__(self, learning_rate=0.001, rho: float = 0.95, epsilon: float = 1e-8, use_locking: bool = False, name: str = "Adadelta"):
        super().__init__(use_locking, name)
        self._lr, self._rho, self._eps = learning_rate, float(rho), float(epsilon)

    def _create_slots(self, var_list):
        for v in var_list:
            self._zeros_slot(v, "accum", self._name)
            self._zeros_slot(v, "accum_update", self._name + "_1")

    def _apply_dense(self, grad, var, prep):
        extra, lr = _lr_inputs(self._lr)
        return get_default_graph().create_node(
            "ApplyAdadelta", [grad] + extra,
            {"var_name": var.var_name, "accum_name": self.get_slot(var, "accum").var_name,
             "accum_update_name": self.get_slot(var, "accum_update").var_name, "lr": lr, "rho": self._rho, "eps": self._eps},
            "update_%s/ApplyAdadelta" % var.var_name.replace("/", "_"), device=var.device)


@register_kernel("ApplyAdadelta", stateful=True)
def _k_apply_adadelta(ctx, node, grad, *extra):
    a = dict(node.attrs, lr=_lr_at_run_time(node, extra))
    var, acc, upd = ctx.store.read(a["var_name"]), ctx.store.read(a["accum_name"]), ctx.store.read(a["accum_update_name"])
    g = grad.to(device=var.device, dtype=var.dtype)
    rho, eps = a["rho"], a["eps"]
    acc.mul_(rho).addcmul_(g, g, value=1.0 - rho)
    step = (upd + eps).sqrt_().div_((acc + eps).sqrt_()).mul_(g)
    upd.mul_(rho).addcmul_(step, step, value=1.0 - rho)
    var.add_(step, alpha=-a["lr"])
    return None


@register_kernel("InitializedValue", stateful=True)
def _k_initialized_value(ctx, node):
    return ctx.store.read(node.attrs["var_name"]).clone()


class ExponentialMovingAverage:
    """``tf.train.ExponentialMovingAverage``: ``apply(var_list)`` creates one shadow variable per variable (same device,
    initialised to the variable's initial value, named ``<var>/ExponentialMovingAverage``) and returns the op that moves every
    shadow towards its variable, ``shadow -= (1 - decay) * (shadow - var)``; with ``num_updates`` the decay is
    ``min(decay, (1 + n) / (10 + n))``.  ``average(var)`` / ``average_name(var)`` / ``variables_to_restore()`` as in TF (evaluate
    with the averages: ``Saver(ema.variables_to_restore())``)."""

    def __init__(self, decay, num_updates=None, zero_debias: bool = False, name: str = "ExponentialMovingAverage"):
        if zero_debias:
            raise NotImplementedError("ExponentialMovingAverage(zero_debias=True) is not provided")
        self._decay, self._num_updates, self._name = decay, num_updates, name
        self._averages: Dict[str, Variable] = {}

    @property
    def name(self) -> str:
        return self._name

    def apply(self, var_list: Optional[Sequence[Variable]] = None) -> Tensor:
        g = get_default_graph()
        var_list = list(trainable_variables() if var_list is None else var_list)
        decay = convert_to_tensor(self._decay, dtype=_ops.float32)
        if self._num_updates is not None:
            n = _ops.cast(convert_to_tensor(self._num_updates), _ops.float32)
            decay = _ops.minimum(decay, _ops.divide(_ops.add(n, 1.0), _ops.add(n, 10.0)))
        updates: List[Tensor] = []
        for v in var_list:
            if not isinstance(v, Variable):
                raise TypeError("ExponentialMovingAverage.apply() takes variables, got %r" % (v,))
            if v.var_name in self._averages:
                raise ValueError("Moving average already computed for: %s" % v.var_name)
            with _device.device(None), _device.device(v.device or None), g.name_scope(None), g.control_dependencies(None):
                # starts at the variable's INITIALISED value (not a second draw of a random initialiser): read from the
                # variable's task after its initialiser ran
                init = g.create_node("InitializedValue", [], {"var_name": v.var_name}, "%s/%s/initialized_value" % (v.var_name, self._name),
                                     v.dtype, v.shape, device=v.device)
                init.control_inputs = list(init.control_inputs) + [v.initializer]
                shadow = Variable(init, trainable=False, name="%s/%s" % (v.var_name, self._name), _exact_name=True)
            self._averages[v.var_name] = shadow
            with _device.device(None), _device.device(v.device or None):
                delta = _ops.multiply(_ops.subtract(shadow, v), _ops.subtract(1.0, decay))
                updates.append(assign(shadow, _ops.subtract(shadow, delta), name="%s/%s_update" % (v.var_name.replace("/", "_"), self._name)))
        return _ops.group(*updates, name=self._name)

    def average(self, var: Variable) -> Optional[Variable]:
        return self._averages.get(var.var_name)

    def average_name(self, var: Variable) -> str:
        return "%s/%s" % (var.var_name, self._name)

    def variables_to_restore(self, moving_avg_variables: Optional[Sequence[Variable]] = None) -> Dict[str, Variable]:
        """Map checkpoint name -> variable to restore INTO: averaged variables are read from their shadow's name."""
        out: Dict[str, Variable] = {}
        avg = {v.var_name for v in (moving_avg_variables or [])} | set(self._averages)
        shadows = {s.var_name for s in self._averages.values()}
        for v in global_variables():
            if v.var_name in shadows:
                continue
            out[self.average_name(v) if v.var_name in avg else v.var_name] = v
        return out


def global_step(sess, global_step_tensor) -> int:
    """``tf.train.global_step(sess, global_step_tensor)``: the step as a python int."""
    return int(sess.run(global_step_tensor))


def _reader(ckpt_dir_or_file: str) -> CheckpointReader:
    path = ckpt_dir_or_file
    if os.path.isdir(resolve_path(path)):
        latest = latest_checkpoint(path)
        if latest is None:
            from ..framework import errors
            raise errors.NotFoundError("no checkpoint in %r" % ckpt_dir_or_file)
        path = latest
    return CheckpointReader(path)


def load_variable(ckpt_dir_or_file: str, name: str):
    """The value of one checkpointed tensor as a numpy array."""
    if name.endswith(":0"):
        name = name[:-2]
    return _reader(ckpt_dir_or_file).get_tensor(name).numpy()


def init_from_checkpoint(ckpt_dir_or_file: str, assignment_map: Dict[str, object]) -> None:
    """Warm start: replace the INITIALISERS of the mapped variables by the checkpoint's values, so the ordinary init op
    (``global_variables_initializer`` / the chief's Scaffold) loads them.  ``assignment_map``: checkpoint tensor name ->
    variable (or its name); a key ending in ``/`` maps a whole scope prefix onto another (``{"old_scope/": "new_scope/"}``)."""
    r = _reader(ckpt_dir_or_file)
    byname = {v.var_name: v for v in global_variables()}
    pairs = []
    for ck, target in assignment_map.items():
        if isinstance(target, str) and (ck.endswith("/") or ck == "") and (target.endswith("/") or target == ""):
            for vn, v in byname.items():
                if vn.startswith(target) and r.has_tensor(ck + vn[len(target):]):
                    pairs.append((ck + vn[len(target):], v))
            continue
        v = target if isinstance(target, Variable) else byname.get(str(target)[:-2] if str(target).endswith(":0") else str(target))
        if v is None:
            raise ValueError("init_from_checkpoint(): no variable %r in the graph" % (target,))
        pairs.append((ck, v))
    for ck, v in pairs:
        t = r.get_tensor(ck)
        if v.shape is not None and tuple(t.shape) != tuple(v.shape):
            raise ValueError("init_from_checkpoint(): shape of %s in the checkpoint %s != %s of variable %s" % (ck, tuple(t.shape), tuple(v.shape), v.var_name))
        with _device.device(None), _device.device(v.device or None):
            v._initial_value = _ops.constant(t.numpy(), dtype=v.dtype)
            v._initializer = get_default_graph().create_node("Assign", [v._initial_value], {"var_name": v.var_name, "init": True},
                                                             "%s/warm_start" % v.var_name, v.dtype, v._initial_value.shape, device=v.device)


def write_graph(graph_or_graph_def, logdir: str, name: str, as_text: bool = True) -> str:
    """Dump the graph's node list (name, op, inputs, device) -- text form by default, like ``tf.train.write_graph``."""
    gd = graph_or_graph_def.as_graph_def() if hasattr(graph_or_graph_def, "as_graph_def") else graph_or_graph_def
    os.makedirs(logdir, exist_ok=True)
    path = os.path.join(logdir, name)
    if as_text:
        with open(path, "w") as f:
            for n in gd["node"]:
                f.write("node {\n  name: %r\n  op: %r\n" % (n["name"], n["op"]))
                for i in n.get("input", []):
                    f.write("  input: %r\n" % (i,))
                if n.get("device"):
                    f.write("  device: %r\n" % (n["device"],))
                f.write("}\n")
    else:
        from ..utils.summary import _graph_def
        with open(path, "wb") as f:
            f.write(_graph_def(gd))
    return path


def add_queue_runner(qr, collection: str = "queue_runners") -> None:
    get_default_graph().add_to_collection(collection, qr)


def start_queue_runners(sess=None, coord=None, daemon: bool = True, start: bool = True, collection: str = "queue_runners"):
    """Start the threads of every QueueRunner in the collection (the sync-replica token queue runner registers itself there
    when a Supervisor / MonitoredSession is not driving it)."""
    threads = []
    for qr in get_default_graph().get_collection(collection):
        threads.extend(qr.create_threads(sess, coord=coord, daemon=daemon, start=start))
    return threads
