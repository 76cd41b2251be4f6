"""Optimizers: GradientDescent, Momentum, Adam.

API parity: ``minimize`` = ``compute_gradients`` + ``apply_gradients`` (+
``global_step += 1``); the split form is what tower averaging uses (reference
``standalone.py:112,123,126``).  Apply ops execute **on the variable's task**:
the gradient is pushed to the ps, the update runs ps-side
(reference ``distributed_mnist.py:115,126``, ``example_between_graph.py:61,73``).

Adam follows the TF formulation the reference trains with, not torch.optim's:
``lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v EMA; var -= lr_t*m/(sqrt(v)+eps)``
(epsilon outside the bias correction), with ``beta1_power``/``beta2_power``
kept as variables colocated with the first parameter.

On a CUDA ps the apply kernels are the fused sm_100a ``optimizer_apply``
kernel (``csrc/optimizer_apply.cu``); the fabric engine additionally fuses
the N-worker mean into the same pass (``parallel/ps_engine.py``).
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..framework import device as _device
from ..framework import ops as _ops
from ..framework.graph import GraphKeys, Tensor, convert_to_tensor, get_default_graph
from ..framework.ops import register_kernel
from ..framework.variables import Variable, assign, assign_add, trainable_variables
from ..utils import native_runtime

__all__ = ["Optimizer", "GradientDescentOptimizer", "MomentumOptimizer", "AdamOptimizer", "adam_reference_step",
           "exponential_decay", "piecewise_constant"]


class Optimizer:
    GATE_NONE, GATE_OP, GATE_GRAPH = 0, 1, 2

    def __init__(self, use_locking: bool = False, name: str = "Optimizer"):
        self._use_locking = use_locking
        self._name = name
        self._slots: Dict[str, Dict[str, Variable]] = {}

    def get_name(self) -> str:
        return self._name

    # -- gradients -----------------------------------------------------------------------------
    def compute_gradients(self, loss, var_list: Optional[Sequence[Variable]] = None, gate_gradients=None,
                          aggregation_method=None, colocate_gradients_with_ops=False, grad_loss=None
                          ) -> List[Tuple[Optional[Tensor], Variable]]:
        if var_list is None:
            var_list = trainable_variables()
        var_list = list(var_list)
        if not var_list:
            raise ValueError("No variables to optimize.")
        loss_t = convert_to_tensor(loss)
        # place the backward pass with the loss (the worker), not under the caller's variable scope device fn
        with _device.device(None), _device.device(loss_t.device or None):
            grads = _ops.gradients(loss_t, [v._node for v in var_list], name="gradients")
        return list(zip(grads, var_list))

    # -- apply ------------------------------------------------------------------------------------
    def apply_gradients(self, grads_and_vars, global_step: Optional[Variable] = None, name: Optional[str] = None) -> Tensor:
        gv = [(g, v) for g, v in grads_and_vars if g is not None]
        if not gv:
            raise ValueError("No gradients provided for any variable.")
        g = get_default_graph()
        var_list = [v for _, v in gv]
        with g.name_scope(name or self._name):
            self._create_slots(var_list)
            prep = self._prepare()
            updates = []
            for grad, var in gv:
                with _device.device(None), _device.device(var.device or None):
                    updates.append(self._apply_dense(convert_to_tensor(grad), var, prep))
            finish = self._finish(updates)
            if global_step is None:
                return _ops.group(*finish, name="update")
            with g.control_dependencies(finish), _device.device(None), _device.device(global_step.device or None):
                inc = assign_add(global_step, _ops.constant(1, dtype=global_step.dtype), name="update_global_step")
            return inc

    def minimize(self, loss, global_step: Optional[Variable] = None, var_list=None, gate_gradients=None,
                 aggregation_method=None, colocate_gradients_with_ops=False, name=None, grad_loss=None) -> Tensor:
        # a ps/worker program whose worker task is bound to a B200 runs its step on the NVLink fabric (fused kernels, device-side
        # accumulate / apply / tokens) without being told to: parallel/auto_fabric.py
        from ..parallel.auto_fabric import maybe_route_minimize
        routed = maybe_route_minimize(self, loss, global_step, var_list)
        if routed is not None:
            self._gradients_applied = True
            return routed
        gv = self.compute_gradients(loss, var_list)
        if all(g is None for g, _ in gv):
            raise ValueError("No gradients provided for any variable, check your graph.")
        return self.apply_gradients(gv, global_step=global_step, name=name)

    # -- slots --------------------------------------------------------------------------------------
    def _zeros_slot(self, var: Variable, slot_name: str, op_name: str) -> Variable:
        named = self._slots.setdefault(slot_name, {})
        s = named.get(var.var_name)
        if s is None:
            g = get_default_graph()
            name = "%s/%s" % (var.var_name, op_name)
            if name in g.variables:
                s = g.variables[name]
            else:
                shape, dt = var.shape, var.dtype
                with _device.device(None), _device.device(var.device or None), g.name_scope(None):
                    s = Variable(lambda: _ops.zeros(shape, dt), trainable=False, name=name, _exact_name=True)
            named[var.var_name] = s
        return s

    def _const_slot(self, var: Variable, slot_name: str, op_name: str, value: float) -> Variable:
        """Like :meth:`_zeros_slot` with every element starting at ``value`` (Adagrad's accumulator, RMSProp's mean square)."""
        named = self._slots.setdefault(slot_name, {})
        s = named.get(var.var_name)
        if s is None:
            g = get_default_graph()
            name = "%s/%s" % (var.var_name, op_name)
            if name in g.variables:
                s = g.variables[name]
            else:
                shape, dt = var.shape, var.dtype
                with _device.device(None), _device.device(var.device or None), g.name_scope(None):
                    s = Variable(lambda: _ops.constant(float(value), dtype=dt, shape=shape), trainable=False, name=name, _exact_name=True)
            named[var.var_name] = s
        return s

    def _scalar_slot(self, colocate: Variable, value: float, name: str) -> Variable:
        g = get_default_graph()
        if name in g.variables:
            return g.variables[name]
        with _device.device(None), _device.device(colocate.device or None), g.name_scope(None):
            return Variable(lambda: _ops.constant(value, dtype=torch.float32), trainable=False, name=name,
                            _exact_name=True)

    def get_slot(self, var: Variable, name: str) -> Optional[Variable]:
        return self._slots.get(name, {}).get(var.var_name)

    def get_slot_names(self) -> List[str]:
        return sorted(self._slots)

    def variables(self) -> List[Variable]:
        out = []
        for d in self._slots.values():
            out.extend(d.values())
        return out

    # -- subclass hooks ---------------------------------------------------------------------------------
    def _create_slots(self, var_list: Sequence[Variable]) -> None:
        pass

    def _prepare(self) -> Any:
        return None

    def _apply_dense(self, grad: Tensor, var: Variable, prep: Any) -> Tensor:
        raise NotImplementedError

    def _finish(self, update_ops: List[Tensor]) -> List[Tensor]:
        return update_ops

    # -- description used by the fabric engine (fused reduce+apply kernel) ---------------------------------
    def fused_spec(self) -> Dict[str, Any]:
        raise NotImplementedError


def _is_dynamic(lr) -> bool:
    return isinstance(lr, (Tensor, Variable))


def _lr_value(lr) -> float:
    if _is_dynamic(lr):
        raise ValueError("this path needs a constant learning rate; got the tensor %r (a schedule such as "
                         "tf.train.exponential_decay works with the graph optimizers, not with the fused fabric engines)"
                         % getattr(lr, "name", lr))
    return float(lr)


def _lr_inputs(lr):
    """(extra input nodes, attr value): a tensor-valued learning rate (``exponential_decay`` ...) is a run-time INPUT
    of the apply op -- evaluated wherever it was placed, shipped to the variable's task like any other tensor."""
    if _is_dynamic(lr):
        return [convert_to_tensor(lr)], None
    return [], float(lr)


def _lr_at_run_time(node, extra) -> float:
    return float(extra[-1]) if node.attrs.get("lr") is None else node.attrs["lr"]


class GradientDescentOptimizer(Optimizer):
    def __init__(self, learning_rate, use_locking: bool = False, name: str = "GradientDescent"):
        super().__init__(use_locking, name)
        self._lr = learning_rate

    def _apply_dense(self, grad, var, prep):
        extra, lr = _lr_inputs(self._lr)
        return get_default_graph().create_node(
            "ApplyGradientDescent", [grad] + extra, {"var_name": var.var_name, "lr": lr},
            "update_%s/ApplyGradientDescent" % var.var_name.replace("/", "_"), device=var.device)

    def fused_spec(self):
        return {"kind": "sgd", "lr": _lr_value(self._lr)}


def _our_kernels_apply(var: torch.Tensor) -> bool:
    """CUDA variables are updated by ``optimizer_apply_kernel``; host variables too under the kernel emulation (tests)."""
    if var.is_cuda:
        return True
    import sys
    mod = sys.modules.get("distributed_tensorflow_b200.ops.cuda_lib")
    return bool(mod is not None and mod.EMULATION and var.dtype == torch.float32 and var.is_contiguous())


@register_kernel("ApplyGradientDescent", stateful=True)
def _k_apply_sgd(ctx, node, grad, *extra):
    var = ctx.store.read(node.attrs["var_name"])
    g = grad.to(device=var.device)
    lr = _lr_at_run_time(node, extra)
    if _our_kernels_apply(var):
        from ..ops import cuda_lib
        cuda_lib.apply_sgd_(var, g, lr)
    else:
        var.sub_(g.to(var.dtype), alpha=lr)
    return None


class MomentumOptimizer(Optimizer):
    def __init__(self, learning_rate, momentum, use_locking: bool = False, name: str = "Momentum",
                 use_nesterov: bool = False):
        super().__init__(use_locking, name)
        self._lr, self._momentum, self._nesterov = learning_rate, momentum, use_nesterov

    def _create_slots(self, var_list):
        for v in var_list:
            self._zeros_slot(v, "momentum", self._name)

    def _apply_dense(self, grad, var, prep):
        slot = self.get_slot(var, "momentum")
        extra, lr = _lr_inputs(self._lr)
        return get_default_graph().create_node(
            "ApplyMomentum", [grad] + extra, {"var_name": var.var_name, "accum_name": slot.var_name,
                                              "lr": lr, "momentum": float(self._momentum),
                                              "nesterov": bool(self._nesterov)},
            "update_%s/ApplyMomentum" % var.var_name.replace("/", "_"), device=var.device)

    def fused_spec(self):
        return {"kind": "momentum", "lr": _lr_value(self._lr), "momentum": float(self._momentum),
                "nesterov": bool(self._nesterov)}


@register_kernel("ApplyMomentum", stateful=True)
def _k_apply_momentum(ctx, node, grad, *extra):
    a = dict(node.attrs, lr=_lr_at_run_time(node, extra))
    var, acc = ctx.store.read(a["var_name"]), ctx.store.read(a["accum_name"])
    g = grad.to(device=var.device)
    if _our_kernels_apply(var):
        from ..ops import cuda_lib
        cuda_lib.apply_momentum_(var, acc, g, a["lr"], a["momentum"], a["nesterov"])
    else:
        g = g.to(var.dtype)
        if not native_runtime.cpu_optimizer_apply(1, var, acc, None, g, a["lr"], momentum=a["momentum"], nesterov=a["nesterov"]):
            acc.mul_(a["momentum"]).add_(g)                       # accum = momentum*accum + grad (TF)
            if a["nesterov"]:
                var.sub_(g * a["lr"] + acc * (a["momentum"] * a["lr"]))
            else:
                var.sub_(acc, alpha=a["lr"])
    return None



class AdagradOptimizer(Optimizer):
    """TF's Adagrad: ``accum += g^2 ; var -= lr * g / sqrt(accum)`` with ``accum`` starting at ``initial_accumulator_value``
    (graph tier; the fabric engines fuse SGD / Momentum / Adam only, so ``minimize`` keeps such a program on the control plane /
    the generic path)."""

    def __init__(self, learning_rate, initial_accumulator_value: float = 0.1, use_locking: bool = False, name: str = "Adagrad"):
        super().__init__(use_locking, name)
        if initial_accumulator_value <= 0.0:
            raise ValueError("initial_accumulator_value must be positive: %s" % initial_accumulator_value)
        self._lr, self._init_acc = learning_rate, float(initial_accumulator_value)

    def _create_slots(self, var_list):
        for v in var_list:
            self._const_slot(v, "accumulator", self._name, self._init_acc)

    def _apply_dense(self, grad, var, prep):
        slot = self.get_slot(var, "accumulator")
        extra, lr = _lr_inputs(self._lr)
        return get_default_graph().create_node(
            "ApplyAdagrad", [grad] + extra, {"var_name": var.var_name, "accum_name": slot.var_name, "lr": lr},
            "update_%s/ApplyAdagrad" % var.var_name.replace("/", "_"), device=var.device)


@register_kernel("ApplyAdagrad", stateful=True)
def _k_apply_adagrad(ctx, node, grad, *extra):
    a = dict(node.attrs, lr=_lr_at_run_time(node, extra))
    var, acc = ctx.store.read(a["var_name"]), ctx.store.read(a["accum_name"])
    g = grad.to(device=var.device, dtype=var.dtype)
    acc.addcmul_(g, g)
    var.addcdiv_(g, acc.sqrt(), value=-a["lr"])
    return None


class RMSPropOptimizer(Optimizer):
    """TF's RMSProp: ``ms = decay * ms + (1 - decay) * g^2 ; mom = momentum * mom + lr * g / sqrt(ms + eps) ; var -= mom``
    (``ms`` starts at one, ``mom`` at zero; ``centered=True`` also tracks the mean gradient and subtracts its square)."""

    def __init__(self, learning_rate, decay: float = 0.9, momentum: float = 0.0, epsilon: float = 1e-10, use_locking: bool = False,
                 centered: bool = False, name: str = "RMSProp"):
        super().__init__(use_locking, name)
        self._lr, self._decay, self._momentum, self._eps, self._centered = learning_rate, float(decay), float(momentum), float(epsilon), bool(centered)

    def _create_slots(self, var_list):
        for v in var_list:
            self._const_slot(v, "rms", self._name, 1.0)
            self._zeros_slot(v, "momentum", self._name + "_1")
            if self._centered:
                self._zeros_slot(v, "mg", self._name + "_2")

    def _apply_dense(self, grad, var, prep):
        extra, lr = _lr_inputs(self._lr)
        attrs = {"var_name": var.var_name, "ms_name": self.get_slot(var, "rms").var_name,
                 "mom_name": self.get_slot(var, "momentum").var_name, "lr": lr, "decay": self._decay, "momentum": self._momentum,
                 "eps": self._eps, "mg_name": self.get_slot(var, "mg").var_name if self._centered else None}
        return get_default_graph().create_node("ApplyRMSProp", [grad] + extra, attrs,
                                               "update_%s/ApplyRMSProp" % var.var_name.replace("/", "_"), device=var.device)


@register_kernel("ApplyRMSProp", stateful=True)
def _k_apply_rmsprop(ctx, node, grad, *extra):
    a = dict(node.attrs, lr=_lr_at_run_time(node, extra))
    var, ms, mom = ctx.store.read(a["var_name"]), ctx.store.read(a["ms_name"]), ctx.store.read(a["mom_name"])
    g = grad.to(device=var.device, dtype=var.dtype)
    ms.mul_(a["decay"]).addcmul_(g, g, value=1.0 - a["decay"])
    denom = ms
    if a["mg_name"] is not None:
        mg = ctx.store.read(a["mg_name"])
        mg.mul_(a["decay"]).add_(g, alpha=1.0 - a["decay"])
        denom = ms - mg * mg
    mom.mul_(a["momentum"]).addcdiv_(g, (denom + a["eps"]).sqrt(), value=a["lr"])
    var.sub_(mom)
    return None


class AdamOptimizer(Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking: bool = False,
                 name: str = "Adam"):
        super().__init__(use_locking, name)
        self._lr, self._beta1, self._beta2, self._eps = learning_rate, beta1, beta2, epsilon
        self._beta1_power: Optional[Variable] = None
        self._beta2_power: Optional[Variable] = None

    def _create_slots(self, var_list):
        first = min(var_list, key=lambda v: v.var_name)
        if self._beta1_power is None:
            self._beta1_power = self._scalar_slot(first, self._beta1, "beta1_power")
            self._beta2_power = self._scalar_slot(first, self._beta2, "beta2_power")
        for v in var_list:
            self._zeros_slot(v, "m", self._name)
            self._zeros_slot(v, "v", self._name + "_1")

    def _get_beta_accumulators(self):
        return self._beta1_power, self._beta2_power

    def _apply_dense(self, grad, var, prep):
        m, v = self.get_slot(var, "m"), self.get_slot(var, "v")
        extra, lr = _lr_inputs(self._lr)
        return get_default_graph().create_node(
            "ApplyAdam", [grad, self._beta1_power._node, self._beta2_power._node] + extra,
            {"var_name": var.var_name, "m_name": m.var_name, "v_name": v.var_name, "lr": lr,
             "beta1": float(self._beta1), "beta2": float(self._beta2), "eps": float(self._eps)},
            "update_%s/ApplyAdam" % var.var_name.replace("/", "_"), device=var.device)

    def _finish(self, update_ops):
        g = get_default_graph()
        with g.control_dependencies(update_ops), _device.device(None), \
                _device.device(self._beta1_power.device or None):
            u1 = assign(self._beta1_power, self._beta1_power._node * self._beta1, name="update_beta1_power")
            u2 = assign(self._beta2_power, self._beta2_power._node * self._beta2, name="update_beta2_power")
        return list(update_ops) + [u1, u2]

    def fused_spec(self):
        return {"kind": "adam", "lr": _lr_value(self._lr), "beta1": float(self._beta1),
                "beta2": float(self._beta2), "eps": float(self._eps)}


@register_kernel("ApplyAdam", stateful=True)
def _k_apply_adam(ctx, node, grad, b1p, b2p, *extra):
    a = dict(node.attrs, lr=_lr_at_run_time(node, extra))
    var, m, v = ctx.store.read(a["var_name"]), ctx.store.read(a["m_name"]), ctx.store.read(a["v_name"])
    g = grad.to(device=var.device)
    b1p, b2p = float(b1p), float(b2p)
    lr_t = a["lr"] * (1.0 - b2p) ** 0.5 / (1.0 - b1p)
    if _our_kernels_apply(var):
        from ..ops import cuda_lib
        cuda_lib.apply_adam_(var, m, v, g, lr_t, a["beta1"], a["beta2"], a["eps"])
    else:
        g = g.to(var.dtype)
        if not native_runtime.cpu_optimizer_apply(2, var, m, v, g, lr_t, beta1=a["beta1"], beta2=a["beta2"], eps=a["eps"]):
            m.mul_(a["beta1"]).add_(g, alpha=1.0 - a["beta1"])
            v.mul_(a["beta2"]).addcmul_(g, g, value=1.0 - a["beta2"])
            var.sub_(lr_t * m / (v.sqrt() + a["eps"]))
    return None


def adam_reference_step(var, m, v, g, t: int, lr=0.001, beta1=0.9, beta2=0.999, eps=1e-8):
    """Closed-form TF-Adam step ``t`` (1-based) on plain tensors; used by tests as the oracle."""
    lr_t = lr * (1.0 - beta2 ** t) ** 0.5 / (1.0 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    var = var - lr_t * m / (v.sqrt() + eps)
    return var, m, v


# ---------------------------------------------------------------------------------------------------------------
# learning-rate schedules: scalar tensors computed from the global step
# ---------------------------------------------------------------------------------------------------------------
def exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase: bool = False, name=None) -> Tensor:
    """``learning_rate * decay_rate ^ (global_step / decay_steps)`` (integer division when ``staircase``)."""
    gs = convert_to_tensor(global_step)
    with _device.device(None), _device.device(gs.device or None):
        return get_default_graph().create_node(
            "LearningRateSchedule", [gs], {"kind": "exponential", "lr": float(learning_rate), "decay_steps": float(decay_steps),
                                           "decay_rate": float(decay_rate), "staircase": bool(staircase)},
            name or "ExponentialDecay", torch.float32, (), device=gs.device)


def piecewise_constant(x, boundaries: Sequence[float], values: Sequence[float], name=None) -> Tensor:
    """``values[0]`` while ``x <= boundaries[0]``, ``values[1]`` while ``x <= boundaries[1]``, ..., ``values[-1]`` after."""
    if len(values) != len(boundaries) + 1:
        raise ValueError("The length of boundaries should be 1 less than the length of values")
    xs = convert_to_tensor(x)
    with _device.device(None), _device.device(xs.device or None):
        return get_default_graph().create_node(
            "LearningRateSchedule", [xs], {"kind": "piecewise", "boundaries": [float(b) for b in boundaries],
                                           "values": [float(v) for v in values]},
            name or "PiecewiseConstant", torch.float32, (), device=xs.device)


def _schedule(kind: str, global_step, default_name: str, name=None, **attrs) -> Tensor:
    gs = convert_to_tensor(global_step)
    with _device.device(None), _device.device(gs.device or None):
        return get_default_graph().create_node("LearningRateSchedule", [gs], dict(attrs, kind=kind), name or default_name,
                                               torch.float32, (), device=gs.device)


def inverse_time_decay(learning_rate, global_step, decay_steps, decay_rate, staircase: bool = False, name=None) -> Tensor:
    """``learning_rate / (1 + decay_rate * global_step / decay_steps)`` (``floor`` of the ratio when ``staircase``)."""
    return _schedule("inverse_time", global_step, "InverseTimeDecay", name, lr=float(learning_rate), decay_steps=float(decay_steps),
                     decay_rate=float(decay_rate), staircase=bool(staircase))


def natural_exp_decay(learning_rate, global_step, decay_steps, decay_rate, staircase: bool = False, name=None) -> Tensor:
    """``learning_rate * exp(-decay_rate * global_step / decay_steps)``."""
    return _schedule("natural_exp", global_step, "NaturalExpDecay", name, lr=float(learning_rate), decay_steps=float(decay_steps),
                     decay_rate=float(decay_rate), staircase=bool(staircase))


def polynomial_decay(learning_rate, global_step, decay_steps, end_learning_rate=0.0001, power=1.0, cycle: bool = False, name=None) -> Tensor:
    """``(lr - end) * (1 - min(step, decay_steps) / decay_steps) ^ power + end``; ``cycle`` stretches ``decay_steps`` to the next
    multiple once it has been passed (TF's definition)."""
    return _schedule("polynomial", global_step, "PolynomialDecay", name, lr=float(learning_rate), decay_steps=float(decay_steps),
                     end=float(end_learning_rate), power=float(power), cycle=bool(cycle))


def cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0, name=None) -> Tensor:
    """``lr * ((1 - alpha) * 0.5 * (1 + cos(pi * min(step, decay_steps) / decay_steps)) + alpha)``."""
    return _schedule("cosine", global_step, "CosineDecay", name, lr=float(learning_rate), decay_steps=float(decay_steps), alpha=float(alpha))


@register_kernel("LearningRateSchedule")
def _k_lr_schedule(ctx, node, step):
    import math
    a, s = node.attrs, float(step)
    if a["kind"] in ("inverse_time", "natural_exp"):
        p = s / a["decay_steps"]
        if a["staircase"]:
            p = float(int(p))
        v = a["lr"] / (1.0 + a["decay_rate"] * p) if a["kind"] == "inverse_time" else a["lr"] * math.exp(-a["decay_rate"] * p)
        return torch.tensor(v, dtype=torch.float32)
    if a["kind"] == "polynomial":
        ds = a["decay_steps"]
        if a["cycle"]:
            ds = ds * max(1.0, math.ceil(s / ds))
        else:
            s = min(s, ds)
        return torch.tensor((a["lr"] - a["end"]) * (1.0 - s / ds) ** a["power"] + a["end"], dtype=torch.float32)
    if a["kind"] == "cosine":
        frac = min(s, a["decay_steps"]) / a["decay_steps"]
        return torch.tensor(a["lr"] * ((1.0 - a["alpha"]) * 0.5 * (1.0 + math.cos(math.pi * frac)) + a["alpha"]), dtype=torch.float32)
    if a["kind"] == "exponential":
        p = s / a["decay_steps"]
        if a["staircase"]:
            p = float(int(p))
        return torch.tensor(a["lr"] * a["decay_rate"] ** p, dtype=torch.float32)
    for b, v in zip(a["boundaries"], a["values"]):
        if s <= b:
            return torch.tensor(v, dtype=torch.float32)
    return torch.tensor(a["values"][-1], dtype=torch.float32)
