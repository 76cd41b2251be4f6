"""Automatic routing of ``tf.train`` programs onto the NVLink fabric.

The reference program builds its model under ``replica_device_setter`` and calls ``opt.minimize(loss, global_step)``
(``/root/reference/distributed_mnist.py:91-126``) -- nothing in it names an engine.  When the process's task is a WORKER
bound to a B200 (``DTF_GPU_INDEX`` / ``Server(gpu_index=...)``) and every trainable variable lives on ``/job:ps``,
``Optimizer.minimize`` (and therefore ``SyncReplicasOptimizer.minimize``) hands the step to
:class:`parallel.strategy.FabricPSStrategy` on its own:

* the graph is pattern-matched against the reference network -- ``-reduce_sum(y_ * log(clip(softmax(xw_plus_b(relu(
  xw_plus_b(x, W1, b1)), W2, b2)), eps, 1)))`` -- and, when it matches (any widths within the kernel's limits), the train
  op becomes ONE ``mlp_step_kernel`` launch per ``Session.run`` on the worker and ONE ``ps_apply_kernel`` per aggregate on
  the ps GPU (``parallel/ps_engine.py``, precision tf32, unicast fabric: the worker's TMA pulls W1 from the ps GPU's HBM);
* any other model keeps the generic fabric engine (pull kernel, autograd over our op kernels, push kernel);
* fetching the loss tensor in the same ``run`` returns the value the fused step computed (no second forward pass);
  fetching it alone -- the reference's validation every 1000 steps (``:160-165``) -- runs the engine's forward-only
  kernel over the fed rows.

``DTF_FABRIC=0`` keeps everything on the control-plane tier; ``DTF_FABRIC=1`` raises instead of silently staying on it when
a GPU worker's program cannot be routed.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional, Sequence

from ..framework.device import DeviceSpec
from ..framework.graph import Tensor, convert_to_tensor
from ..framework.variables import Variable, trainable_variables

__all__ = ["maybe_route_minimize", "match_reference_mlp"]


def _mode() -> str:
    return os.environ.get("DTF_FABRIC", "auto").lower()


def _producer(t: Tensor, *op_types: str) -> Optional[Tensor]:
    """``t`` itself if it is one of ``op_types``, looking through Identity nodes."""
    while t is not None and t.op_type == "Identity" and t.inputs:
        t = t.inputs[0]
    return t if (t is not None and t.op_type in op_types) else None


def _var_of(t: Tensor, by_node: Dict[int, Variable]) -> Optional[Variable]:
    while t is not None and t.op_type == "Identity" and t.inputs:
        t = t.inputs[0]
    return by_node.get(t.id) if t is not None else None


def match_reference_mlp(loss: Tensor, variables: Sequence[Variable]) -> Optional[Dict[str, Any]]:
    """Recognise the reference network (``distributed_mnist.py:106-113``).  Returns ``{"x", "y_": placeholders,
    "hid_w", "hid_b", "sm_w", "sm_b": variables, "clip_min": float}`` or ``None``."""
    by_node = {v._node.id: v for v in variables}
    fused = _producer(loss, "ClippedSoftmaxXentSum")       # the same loss as ONE node (dtf.nn.clipped_softmax_xent_sum)
    if fused is not None:
        labels, lo = _producer(fused.inputs[1], "Placeholder"), fused.attrs.get("clip_min", 1e-10)
        l2 = _producer(fused.inputs[0], "XwPlusB")
        if labels is None or l2 is None or not (0.0 <= float(lo) < 1e-3):
            return None
    else:
        neg = _producer(loss, "Neg")
        if neg is None:
            return None
        red = _producer(neg.inputs[0], "Sum")
        if red is None or red.attrs.get("axis") is not None:
            return None
        mul = _producer(red.inputs[0], "Mul")
        if mul is None:
            return None
        a, b = mul.inputs
        log = _producer(a, "Log") or _producer(b, "Log")
        if log is None:
            return None
        labels = b if _producer(a, "Log") is not None else a
        labels = _producer(labels, "Placeholder")
        clip = _producer(log.inputs[0], "ClipByValue")
        if labels is None or clip is None:
            return None
        lo, hi = clip.attrs.get("lo"), clip.attrs.get("hi")
        if lo is None or hi is None or float(hi) != 1.0 or not (0.0 <= float(lo) < 1e-3):
            return None
        sm = _producer(clip.inputs[0], "Softmax")
        if sm is None:
            return None
        l2 = _producer(sm.inputs[0], "XwPlusB")
        if l2 is None:
            return None
    relu = _producer(l2.inputs[0], "Relu")
    if relu is None:
        return None
    l1 = _producer(relu.inputs[0], "XwPlusB")
    if l1 is None:
        return None
    x = _producer(l1.inputs[0], "Placeholder")
    w1, b1 = _var_of(l1.inputs[1], by_node), _var_of(l1.inputs[2], by_node)
    w2, b2 = _var_of(l2.inputs[1], by_node), _var_of(l2.inputs[2], by_node)
    if x is None or None in (w1, b1, w2, b2) or len({id(v) for v in (w1, b1, w2, b2)}) != 4:
        return None
    if len(variables) != 4:
        return None                                        # other trainable variables: not (only) this network
    s1, s2 = [int(d) for d in w1.shape], [int(d) for d in w2.shape]
    if len(s1) != 2 or len(s2) != 2 or s1[1] != s2[0] or [int(d) for d in b1.shape] != [s1[1]] or [int(d) for d in b2.shape] != [s2[1]]:
        return None
    return {"x": x, "y_": labels, "hid_w": w1, "hid_b": b1, "sm_w": w2, "sm_b": b2, "clip_min": float(lo),
            "in_dim": s1[0], "hidden": s1[1], "classes": s2[1]}


def _worker_server():
    from .server import local_servers
    srvs = [s for s in local_servers() if s.job_name == "worker" and s.gpu_index is not None]
    return srvs[0] if len(srvs) == 1 else None


def maybe_route_minimize(optimizer, loss, global_step: Optional[Variable], var_list=None) -> Optional[Tensor]:
    """Called first thing by ``Optimizer.minimize``.  Returns the fabric train op, or ``None`` to build the ordinary
    graph-tier update."""
    mode = _mode()
    if mode in ("0", "off", "false") or getattr(optimizer, "_in_fabric_route", False):
        return None

    def decline(why: str):
        if mode in ("1", "on", "force", "true"):
            raise RuntimeError("DTF_FABRIC=1 but this program cannot run on the fabric: " + why)
        return None
    srv = _worker_server()
    if srv is None:
        return decline("no worker Server bound to a GPU in this process (set DTF_GPU_INDEX)") if mode != "auto" else None
    if global_step is None:
        return decline("minimize() without a global_step")
    try:
        optimizer.fused_spec()
    except Exception as e:      # noqa: BLE001 - e.g. a learning-rate schedule tensor: stays on the graph tier
        return decline("optimizer has no fused form (%s)" % e)
    vars_ = list(var_list) if var_list is not None else trainable_variables()
    if not vars_:
        return decline("no trainable variables")
    for v in vars_:
        if DeviceSpec.from_string(v.device).job != "ps":
            return decline("variable %s is not on a ps task" % v.var_name)
    from .strategy import FabricPSStrategy
    strategy = FabricPSStrategy(srv)
    optimizer._in_fabric_route = True
    try:
        train_op, _ = strategy.minimize(optimizer, loss, global_step, vars_)
    finally:
        optimizer._in_fabric_route = False
    optimizer._fabric_strategy = strategy
    inner = getattr(optimizer, "_opt", None)
    if inner is not None:
        inner._fabric_strategy = strategy
    print("dtf: minimize() routed onto the NVLink fabric (%s step, %s)" % (
        "fused MLP" if strategy.mlp is not None else "generic", "sync replicas" if optimizer.fused_spec().get("sync") else "async"))
    return train_op
