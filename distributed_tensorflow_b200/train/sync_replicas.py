"""``SyncReplicasOptimizer``: synchronous data-parallel training over parameter servers.

Capability parity (reference ``distributed_mnist.py:118-126``,
``example_between_graph.py:66-73``):

* per variable, a conditional accumulator on the variable's ps task; a worker
  step pushes ``(grad, local_step)`` into every accumulator -- gradients stamped
  older than the accumulator's global step are dropped as stale;
* the worker then dequeues one token from ``sync_token_q`` (blocking) and
  adopts it as its new ``local_step``;
* the chief runs a background loop: ``take_grad(replicas_to_aggregate)`` (mean
  of the fresh gradients) for every variable -> wrapped optimizer's apply
  (``global_step += 1`` once per aggregate) -> enqueue ``total_num_replicas``
  tokens carrying the new global step;
* ``replicas_to_aggregate < total_num_replicas`` gives backup workers;
* ``make_session_run_hook(is_chief)`` initialises ``local_step``, and on the
  chief pre-fills the token queue and starts the loop.

On the B200 fabric the same protocol runs on the device: gradient slots +
stamps in ps HBM written by worker kernels over NVLink, reduce+mean+apply in
one ps kernel, tokens as release/acquire flags (``parallel/ps_engine.py``).
"""
from __future__ import annotations

import threading
from typing import List, Optional, Tuple

import torch

from ..framework import device as _device
from ..framework import errors
from ..framework import ops as _ops
from ..framework.graph import GraphKeys, Tensor, convert_to_tensor, get_default_graph
from ..framework.ops import register_kernel
from ..framework.variables import (Variable, assign, global_variables, report_uninitialized_variables)
from ..parallel.ps_state import ConditionalAccumulator, FIFOQueue
from .coordinator import Coordinator, QueueRunner
from .hooks import SessionRunHook
from .optimizer import Optimizer

__all__ = ["SyncReplicasOptimizer"]


# -- kernels for the ps-resident resources ---------------------------------------------------------
def _acc(ctx, node) -> ConditionalAccumulator:
    name = node.attrs["acc_name"]
    return ctx.store.get_resource(name, lambda: _make_accumulator(name))


def _queue(ctx, node) -> FIFOQueue:
    name = node.attrs["queue_name"]
    return ctx.store.get_resource(name, lambda: _make_queue(name))


def _make_accumulator(name):
    from ..utils import native_runtime
    return native_runtime.make_accumulator(name)


def _make_queue(name):
    from ..utils import native_runtime
    return native_runtime.make_queue(name)


@register_kernel("AccumulatorApplyGrad", stateful=True)
def _k_acc_apply(ctx, node, grad, local_step):
    return torch.tensor(_acc(ctx, node).apply_grad(grad, int(local_step)))


@register_kernel("AccumulatorTakeGrad", stateful=True)
def _k_acc_take(ctx, node):
    return _acc(ctx, node).take_grad(node.attrs["num_required"], cancel=getattr(ctx, "cancel_event", None))


@register_kernel("AccumulatorSetGlobalStep", stateful=True)
def _k_acc_set(ctx, node, step):
    _acc(ctx, node).set_global_step(int(step))
    return None


@register_kernel("AccumulatorNumAccumulated", stateful=True)
def _k_acc_num(ctx, node):
    return torch.tensor(_acc(ctx, node).num_accumulated())


@register_kernel("QueueDequeue", stateful=True)
def _k_q_deq(ctx, node):
    v = _queue(ctx, node).dequeue(cancel=getattr(ctx, "cancel_event", None))
    return torch.tensor(int(v), dtype=torch.int64)


@register_kernel("QueueEnqueueMany", stateful=True)
def _k_q_enq(ctx, node, value):
    _queue(ctx, node).enqueue_many([int(value)] * int(node.attrs["count"]))
    return None


@register_kernel("QueueSize", stateful=True)
def _k_q_size(ctx, node):
    return torch.tensor(_queue(ctx, node).size())


@register_kernel("QueueClose", stateful=True)
def _k_q_close(ctx, node):
    _queue(ctx, node).close()
    return None


@register_kernel("QueueRenew", stateful=True)
def _k_q_renew(ctx, node):
    """Chief start-up: replace a CLOSED queue of this name (left behind by a previous chief on a long-lived ps) with
    a fresh one; an open queue is kept (its tokens belong to replicas that are still running)."""
    name = node.attrs["queue_name"]
    with ctx.store._lock:
        q = ctx.store.resources.get(name)
        if q is not None and getattr(q, "is_closed", lambda: False)():
            ctx.store.resources[name] = _make_queue(name)
    return None


class SyncReplicasOptimizer(Optimizer):
    def __init__(self, opt: Optimizer, replicas_to_aggregate: int, total_num_replicas: Optional[int] = None,
                 variable_averages=None, variables_to_average=None, use_locking: bool = False,
                 name: str = "sync_replicas"):
        if total_num_replicas is None:
            total_num_replicas = replicas_to_aggregate
        super().__init__(use_locking, name)
        print("SyncReplicasV2: replicas_to_aggregate=%s; total_num_replicas=%s"
              % (replicas_to_aggregate, total_num_replicas))
        self._opt = opt
        self._replicas_to_aggregate = int(replicas_to_aggregate)
        self._total_num_replicas = int(total_num_replicas)
        self._tokens_per_step = max(self._total_num_replicas, self._replicas_to_aggregate)
        self._gradients_applied = False
        self._global_step: Optional[Variable] = None
        self._local_step: Optional[Variable] = None
        self._sync_token_queue_name = "sync_token_q"
        self._chief_queue_runner: Optional[QueueRunner] = None
        self._accumulator_list: List[Tuple[str, str]] = []
        self.sync_op: Optional[Tensor] = None
        self.local_step_init_op: Optional[Tensor] = None
        self.chief_init_op: Optional[Tensor] = None
        self.ready_for_local_init_op: Optional[Tensor] = None

    # -- pass-throughs ----------------------------------------------------------------------------
    def compute_gradients(self, *args, **kwargs):
        return self._opt.compute_gradients(*args, **kwargs)

    def get_slot(self, *args, **kwargs):
        return self._opt.get_slot(*args, **kwargs)

    def get_slot_names(self, *args, **kwargs):
        return self._opt.get_slot_names(*args, **kwargs)

    def variables(self):
        return self._opt.variables()

    def fused_spec(self):
        spec = dict(self._opt.fused_spec())
        spec.update(sync=True, replicas_to_aggregate=self._replicas_to_aggregate,
                    total_num_replicas=self._total_num_replicas)
        return spec

    @property
    def replicas_to_aggregate(self) -> int:
        return self._replicas_to_aggregate

    @property
    def total_num_replicas(self) -> int:
        return self._total_num_replicas

    # -- the synchronous train op --------------------------------------------------------------------
    def apply_gradients(self, grads_and_vars, global_step: Optional[Variable] = None, name: Optional[str] = None) -> Tensor:
        if not grads_and_vars:
            raise ValueError("Must supply at least one variable")
        if global_step is None:
            raise ValueError("Global step is required to check staleness")
        g = get_default_graph()
        self._global_step = global_step
        worker_device = _device.current_device_for_ops()
        train_ops: List[Tensor] = []
        aggregated: List[Tuple[Optional[Tensor], Variable]] = []
        with g.name_scope(name or self._name):
            # worker-local step counter (NOT placed by the ps setter)
            with _device.device(None), _device.device(worker_device or None), g.name_scope(None):
                self._local_step = Variable(lambda: _ops.constant(0, dtype=torch.int64), trainable=False,
                                            collections=[GraphKeys.LOCAL_VARIABLES], name="sync_rep_local_step")
            self.local_step_init_op = assign(self._local_step, global_step._node, name="local_step_init")
            chief_init_ops = []
            for grad, var in grads_and_vars:
                if grad is None:
                    aggregated.append((None, var))
                    continue
                acc_name = var.var_name + "/grad_accum"
                dev = var.device
                self._accumulator_list.append((acc_name, dev))
                train_ops.append(g.create_node("AccumulatorApplyGrad",
                                               [convert_to_tensor(grad), self._local_step._node],
                                               {"acc_name": acc_name}, "AccumulatorApplyGradient", device=dev))
                take = g.create_node("AccumulatorTakeGrad", [], {"acc_name": acc_name,
                                                                 "num_required": self._replicas_to_aggregate},
                                     "AccumulatorTakeGradient", var.dtype, var.shape, device=dev)
                aggregated.append((take, var))
                chief_init_ops.append(g.create_node("AccumulatorSetGlobalStep", [global_step._node],
                                                    {"acc_name": acc_name}, "SetGlobalStep", device=dev))
            # chief-side: apply the aggregated gradients with the wrapped optimizer (global_step += 1 once)
            update_op = self._opt.apply_gradients(aggregated, global_step=global_step)
            qdev = global_step.device
            # worker-side: after pushing all grads, block on a token and adopt it as the local step
            with g.control_dependencies(train_ops):
                token = g.create_node("QueueDequeue", [], {"queue_name": self._sync_token_queue_name},
                                      "sync_token_q_Dequeue", torch.int64, (), device=qdev)
            train_op = assign(self._local_step, token, name="set_local_step")
            # chief-side: once the update ran, hand out tokens carrying the NEW global step
            with g.control_dependencies([update_op]):
                # everything between the aggregate and the tokens stays on the global step's device (TF colocates the sync op
                # with the global step): the chief's loop is ONE ps segment -- take_grad, apply, enqueue in one round trip
                with _device.device(None), _device.device(qdev or None):
                    step_after = _ops.identity(global_step._node, name="global_step_after_update")
                    self.sync_op = g.create_node("QueueEnqueueMany", [step_after],
                                                 {"queue_name": self._sync_token_queue_name,
                                                  "count": self._tokens_per_step}, "sync_token_q_EnqueueMany",
                                                 device=qdev)
            with _device.device(None), _device.device(qdev or None):
                # when the chief's coordinator stops, the token queue is closed (TF's QueueRunner does the same with
                # cancel_pending_enqueues): replicas still blocked in the dequeue get OutOfRangeError = a clean end of
                # training, instead of waiting for tokens nobody will produce any more
                close_op = g.create_node("QueueClose", [], {"queue_name": self._sync_token_queue_name},
                                         "sync_token_q_Close", device=qdev)
                chief_init_ops.append(g.create_node("QueueRenew", [], {"queue_name": self._sync_token_queue_name},
                                                    "sync_token_q_Renew", device=qdev))
            self._chief_queue_runner = QueueRunner(self._sync_token_queue_name, [self.sync_op], close_op=close_op)
            # A replica whose training loop ends cleanly (StopAtStepHook saw the last step) while another replica is still
            # inside a step would leave that one waiting for a token whose aggregate needs the departed replica's
            # gradient -- the init tokens let replicas run a step apart, and stale drops at the start use the spare
            # tokens up.  The departing replica therefore leaves one "farewell" token per other replica: whoever is
            # blocked finishes its run, sees the same global step and stops too.  (TF has this end-of-training hazard.)
            self._farewell_op = None
            if self._total_num_replicas > 1:
                self._farewell_op = g.create_node("QueueEnqueueMany", [self._global_step._node],
                                                  {"queue_name": self._sync_token_queue_name,
                                                   "count": int(self._total_num_replicas - 1)},
                                                  "sync_token_q_farewell", device=qdev)
            # like TF: the chief's init op also initialises ITS local_step (mnist_replica.py passes chief_init_op as the
            # chief's local_init_op and local_step_init_op as everybody else's)
            self.chief_init_op = _ops.group(self.local_step_init_op, *chief_init_ops, name="chief_init")
            self.ready_for_local_init_op = report_uninitialized_variables(global_variables())
            self._gradients_applied = True
            return train_op

    def get_chief_queue_runner(self) -> QueueRunner:
        if not self._gradients_applied:
            raise ValueError("Should be called after apply_gradients().")
        return self._chief_queue_runner

    def get_init_tokens_op(self, num_tokens: int = -1) -> Tensor:
        if not self._gradients_applied:
            raise ValueError("get_init_tokens_op() should be called after apply_gradients().")
        tokens_needed = self._replicas_to_aggregate - self._total_num_replicas
        if num_tokens == -1:
            num_tokens = self._replicas_to_aggregate
        elif num_tokens < tokens_needed:
            raise ValueError("Too few tokens to finish the first step: %d (given) vs %d (needed)"
                             % (num_tokens, tokens_needed))
        g = get_default_graph()
        if num_tokens <= 0:
            return _ops.no_op(name="no_init_tokens")
        qdev = self._global_step.device
        return g.create_node("QueueEnqueueMany", [self._global_step._node],
                             {"queue_name": self._sync_token_queue_name, "count": int(num_tokens)},
                             "sync_token_q_init", device=qdev)

    def make_session_run_hook(self, is_chief: bool, num_tokens: int = -1) -> "SyncReplicasOptimizerHook":
        return SyncReplicasOptimizerHook(self, is_chief, num_tokens)


class SyncReplicasOptimizerHook(SessionRunHook):
    """Initialises ``local_step``; on the chief also seeds the token queue and starts the aggregate loop."""

    def __init__(self, sync_optimizer: SyncReplicasOptimizer, is_chief: bool, num_tokens: int):
        self._sync_optimizer = sync_optimizer
        self._is_chief = is_chief
        self._num_tokens = num_tokens
        self._q_runner = None
        self._init_tokens_op = None
        self._local_init_op = None
        self._threads: List[threading.Thread] = []

    def _on_fabric(self) -> bool:
        # accumulators, token queue and the chief's aggregate loop are the fabric engine's device-side protocol there
        return getattr(self._sync_optimizer, "_fabric_strategy", None) is not None

    def begin(self):
        if not self._sync_optimizer._gradients_applied:
            raise ValueError("SyncReplicasOptimizer.apply_gradients should be called before using the hook.")
        if self._on_fabric():
            return
        self._local_init_op = self._sync_optimizer.local_step_init_op
        if self._is_chief:
            self._q_runner = self._sync_optimizer.get_chief_queue_runner()
            self._init_tokens_op = self._sync_optimizer.get_init_tokens_op(self._num_tokens)
            self._chief_init_op = self._sync_optimizer.chief_init_op

    def after_create_session(self, session, coord):
        if self._on_fabric():
            return
        raw = getattr(session, "raw_session", lambda: session)()
        raw.run(self._local_init_op)
        if self._is_chief:
            raw.run(self._chief_init_op)
            raw.run(self._init_tokens_op)
            self._threads = self._q_runner.create_threads(raw, coord=coord, daemon=True, start=True)

    def end(self, session):
        # Clean end of the chief's training loop: close the token queue NOW, while the session is still open (the
        # queue runner's close-on-stop thread races with the session teardown).  Replicas that are still running drain
        # the remaining tokens and then get OutOfRangeError from the dequeue = a clean end of their loop.
        if self._on_fabric():
            self._sync_optimizer._fabric_strategy.farewell()     # nobody will wait for this replica's gradient any more
            return
        raw = getattr(session, "raw_session", lambda: session)()
        if self._is_chief and self._q_runner is not None and self._q_runner.close_op is not None:
            try:
                raw.run(self._q_runner.close_op)
            except Exception:      # noqa: BLE001 - the ps may already be gone
                pass
        elif not self._is_chief and getattr(self._sync_optimizer, "_farewell_op", None) is not None:
            try:
                raw.run(self._sync_optimizer._farewell_op)       # see apply_gradients: nobody waits for our gradient
            except Exception:      # noqa: BLE001 - queue already closed by the chief / ps gone: nothing to release
                pass
