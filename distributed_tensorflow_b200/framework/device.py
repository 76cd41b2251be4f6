"""Device strings, device scopes and ``replica_device_setter``.

Capability parity:
* grammar ``/job:<j>/replica:<r>/task:<t>/(cpu|gpu|device:CPU|device:GPU):<n>``
  with partial specs that merge field-wise, inner scope winning
  (reference ``example_in_graph.py:32,53,57``, ``standalone.py:22,26,41,104,115``);
* ``replica_device_setter(cluster, worker_device)`` -- variables go to
  ``/job:ps/task:k`` round-robin **by creation order**, everything else to
  ``worker_device`` (reference ``distributed_mnist.py:91-94``,
  ``example_between_graph.py:43-45``).

B200 mapping: ``/gpu:n`` is CUDA ordinal ``n`` inside the owning task's
process.  When a task is bound to one B200 by the launcher (one process per
GPU), ``/gpu:0`` is *that* GPU.
"""
from __future__ import annotations

import re
import threading
from typing import Callable, List, Optional, Union

__all__ = ["DeviceSpec", "device", "replica_device_setter", "current_device", "apply_device_stack",
           "VARIABLE_OP_TYPES"]

_FIELD_RE = re.compile(r"^(job|replica|task|device|cpu|gpu):(.+)$", re.IGNORECASE)

# op types the setter treats as "parameters" (TF: Variable, VariableV2, VarHandleOp, ...)
VARIABLE_OP_TYPES = ("Variable", "VariableV2", "VarHandleOp", "AutoReloadVariable",
                     "MutableHashTable", "MutableHashTableV2")


class DeviceSpec:
    __slots__ = ("job", "replica", "task", "device_type", "device_index")

    def __init__(self, job: Optional[str] = None, replica: Optional[int] = None, task: Optional[int] = None,
                 device_type: Optional[str] = None, device_index: Optional[int] = None):
        self.job, self.replica, self.task = job, replica, task
        self.device_type = device_type.upper() if device_type else None
        self.device_index = device_index

    @classmethod
    def from_string(cls, spec: Optional[str]) -> "DeviceSpec":
        d = cls()
        if not spec:
            return d
        for part in str(spec).split("/"):
            if not part:
                continue
            m = _FIELD_RE.match(part)
            if not m:
                raise ValueError("malformed device specification %r (bad component %r)" % (spec, part))
            key, val = m.group(1).lower(), m.group(2)
            if key == "job":
                d.job = val
            elif key == "replica":
                d.replica = int(val)
            elif key == "task":
                d.task = int(val)
            elif key in ("cpu", "gpu"):
                d.device_type = key.upper()
                d.device_index = None if val == "*" else int(val)
            else:  # device:GPU:0
                typ, _, idx = val.partition(":")
                d.device_type = typ.upper()
                d.device_index = None if idx in ("", "*") else int(idx)
        return d

    def merge_from(self, inner: "DeviceSpec") -> "DeviceSpec":
        """Return self overridden field-by-field by ``inner`` (inner scope wins)."""
        return DeviceSpec(
            inner.job if inner.job is not None else self.job,
            inner.replica if inner.replica is not None else self.replica,
            inner.task if inner.task is not None else self.task,
            inner.device_type if inner.device_type is not None else self.device_type,
            inner.device_index if inner.device_index is not None else self.device_index,
        )

    def to_string(self) -> str:
        s = ""
        if self.job is not None:
            s += "/job:%s" % self.job
        if self.replica is not None:
            s += "/replica:%d" % self.replica
        if self.task is not None:
            s += "/task:%d" % self.task
        if self.device_type is not None:
            s += "/device:%s:%s" % (self.device_type, "*" if self.device_index is None else self.device_index)
        return s

    @property
    def task_key(self):
        """(job, task) or None when the spec does not name a remote task."""
        if self.job is None:
            return None
        return (self.job, 0 if self.task is None else self.task)

    def __eq__(self, other: object) -> bool:
        return isinstance(other, DeviceSpec) and self.to_string() == other.to_string()

    def __hash__(self) -> int:
        return hash(self.to_string())

    def __repr__(self) -> str:
        return "DeviceSpec(%r)" % self.to_string()


# ---------------------------------------------------------------------------
# device scope stack (per thread; graphs read it when a node is created)
# ---------------------------------------------------------------------------
_tls = threading.local()


def _stack() -> List[Union[str, Callable, None]]:
    st = getattr(_tls, "stack", None)
    if st is None:
        st = _tls.stack = []
    return st


class device:
    """``with dtf.device("/job:ps/task:0/cpu:0"):`` or ``with dtf.device(fn):``.

    ``fn`` is a device *function*: called with the node being created, it
    returns a device string.  ``None`` clears the scope (like TF).
    """

    def __init__(self, device_name_or_function: Union[str, Callable, None]):
        self._spec = device_name_or_function

    def __enter__(self):
        _stack().append(self._spec)
        return self

    def __exit__(self, *exc):
        _stack().pop()
        return False


def apply_device_stack(node) -> str:
    """Resolve the device of ``node`` from the active scopes, outermost first."""
    spec = DeviceSpec()
    for entry in _stack():
        if entry is None:
            spec = DeviceSpec()
        elif callable(entry):
            got = entry(node)
            if got:
                spec = spec.merge_from(DeviceSpec.from_string(got))
        else:
            spec = spec.merge_from(DeviceSpec.from_string(entry))
    return spec.to_string()


def current_device() -> str:
    class _Probe:  # a non-variable op
        op_type = "NoOp"
        name = "_probe"
        device = ""
    return apply_device_stack(_Probe())


# ---------------------------------------------------------------------------
# replica_device_setter
# ---------------------------------------------------------------------------
class _RoundRobin:
    def __init__(self, num_tasks: int):
        self._n, self._next = num_tasks, 0

    def __call__(self, node) -> int:
        t = self._next
        self._next = (self._next + 1) % self._n
        return t


class _GreedyLoad:
    """Alternative strategy: place each variable on the least-loaded ps (bytes)."""

    def __init__(self, num_tasks: int, load_fn: Callable):
        self._loads = [0] * num_tasks
        self._fn = load_fn

    def __call__(self, node) -> int:
        t = min(range(len(self._loads)), key=self._loads.__getitem__)
        self._loads[t] += int(self._fn(node))
        return t


class _ReplicaDeviceChooser:
    def __init__(self, ps_tasks: int, ps_device: str, worker_device: str, merge_devices: bool,
                 ps_ops, ps_strategy):
        self._ps_tasks, self._ps_device, self._worker_device = ps_tasks, ps_device, worker_device
        self._merge, self._ps_ops, self._strategy = merge_devices, tuple(ps_ops), ps_strategy

    def __call__(self, node) -> str:
        current = DeviceSpec.from_string(getattr(node, "device", "") or "")
        if not self._merge and current.to_string():
            return current.to_string()
        if self._ps_tasks and self._ps_device and node.op_type in self._ps_ops:
            ps = DeviceSpec.from_string(self._ps_device)
            if ps.task is None:
                ps.task = self._strategy(node)
            # fields already pinned on the node win over the setter (TF merge rule)
            return ps.merge_from(current).to_string()
        worker = DeviceSpec.from_string(self._worker_device or "")
        return worker.merge_from(current).to_string()


def replica_device_setter(ps_tasks: int = 0, ps_device: str = "/job:ps", worker_device: str = "/job:worker",
                          merge_devices: bool = True, cluster=None, ps_ops=None, ps_strategy=None):
    """Device function placing variables on ps tasks (round-robin) and the rest on the worker.

    With the MNIST model and 2 ps tasks the creation order
    ``global_step, hid_w, hid_b, sm_w, sm_b`` lands on ``ps0, ps1, ps0, ps1, ps0``
    (SURVEY A5).  Optimizer slots are colocated with their variable by the
    optimizers themselves, not by this function.
    """
    if cluster is not None:
        from ..parallel.cluster import ClusterSpec
        spec = cluster if isinstance(cluster, ClusterSpec) else ClusterSpec(cluster)
        ps_job = DeviceSpec.from_string(ps_device).job or "ps"
        ps_tasks = spec.num_tasks(ps_job) if ps_job in spec.jobs else 0
    if not ps_tasks:
        return None
    if ps_ops is None:
        ps_ops = VARIABLE_OP_TYPES
    if ps_strategy is None:
        ps_strategy = _RoundRobin(ps_tasks)
    if not callable(ps_strategy):
        raise TypeError("ps_strategy must be callable")
    return _ReplicaDeviceChooser(ps_tasks, ps_device, worker_device, merge_devices, ps_ops, ps_strategy)


def greedy_load_balancing_strategy(num_tasks: int, load_fn: Callable = None):
    if load_fn is None:
        def load_fn(node):
            shape = getattr(node, "shape", None) or ()
            n = 1
            for d in shape:
                n *= int(d or 1)
            return n * 4
    return _GreedyLoad(num_tasks, load_fn)


def current_device_for_ops() -> str:
    """Device a plain (non-variable) op would get under the active scopes -- the worker device
    under ``replica_device_setter``.  Used to pin worker-local state (e.g. ``sync_rep_local_step``)."""
    return current_device()
