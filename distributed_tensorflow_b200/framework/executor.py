"""Task-local graph executor and resource store.

A *task* (``/job:x/task:i``) owns a :class:`ResourceStore` (variables,
optimizer slots, gradient accumulators, token queues) and executes the graph
segments the session's master assigns to it.  This is the B200-native
counterpart of TF's per-task worker service: values are
``torch.Tensor`` on the task's device, kernels come from ``framework/ops.py``
(sm_100a kernels when the device is a GPU), reverse-mode gradients use the
autograd tape of the run.
"""
from __future__ import annotations

import threading
import time
from typing import Any, Dict, Iterable, List, Optional, Sequence, Set, Tuple

import torch

from . import fusion as _fusion
from . import ops as _ops
from .device import DeviceSpec
from .errors import FailedPreconditionError
from .graph import Tensor

__all__ = ["ResourceStore", "ExecContext", "execute", "needed_nodes", "schedule_segments"]


class ResourceStore:
    """Named mutable state of one task.  Thread-safe; shared by every session on the task."""

    def __init__(self, name: str = "local") -> None:
        self.name = name
        self._vars: Dict[str, torch.Tensor] = {}
        self._lock = threading.RLock()
        self.resources: Dict[str, Any] = {}       # accumulators, queues, engine handles
        self._uninit: Set[str] = set()            # bound (externally owned) storage that has not been assigned yet
        self._bound: Set[str] = set()
        self.incarnation = time.time_ns()

    # -- variables ---------------------------------------------------------------
    def read(self, name: str) -> torch.Tensor:
        try:
            if name in self._uninit:
                raise KeyError(name)
            return self._vars[name]
        except KeyError:
            raise FailedPreconditionError("Attempting to use uninitialized value %s" % name) from None

    def is_initialized(self, name: str) -> bool:
        return name in self._vars and name not in self._uninit

    def assign(self, name: str, value: torch.Tensor, device=None) -> torch.Tensor:
        with self._lock:
            cur = self._vars.get(name)
            if cur is not None and cur.shape == value.shape and cur.dtype == value.dtype:
                cur.copy_(value)            # keep storage identity (peer-mapped buffers stay valid)
                self._uninit.discard(name)
                return cur
            if cur is not None and name in self._bound:
                cur.copy_(value.to(cur.dtype).reshape(cur.shape))     # bound storage never gets replaced
                self._uninit.discard(name)
                return cur
            t = value.detach().clone()
            if device is not None and t.device != torch.device(device):
                t = t.to(device)
            self._vars[name] = t
            return t

    def assign_add(self, name: str, delta: torch.Tensor) -> torch.Tensor:
        with self._lock:
            cur = self.read(name)
            cur.add_(delta.to(device=cur.device, dtype=cur.dtype))
            return cur.clone() if cur.dim() == 0 else cur

    def bind(self, name: str, tensor: torch.Tensor, initialized: bool = True) -> None:
        """Adopt externally-owned storage (e.g. a slice of a fabric ps shard) as the variable.  A value that was
        already assigned under this name (the chief may have initialised before the fabric came up) is carried over."""
        with self._lock:
            old = self._vars.get(name)
            had = old is not None and name not in self._uninit
            if had:
                tensor.copy_(old.to(device=tensor.device, dtype=tensor.dtype).reshape(tensor.shape))
            self._vars[name] = tensor
            self._bound.add(name)
            if initialized or had:
                self._uninit.discard(name)
            else:
                self._uninit.add(name)

    def unbind(self, name: str) -> None:
        """Give up externally-owned storage (the fabric engine behind it is being torn down): the variable keeps its current
        value and initialisation state in storage of its own, so a later ``bind`` carries the value into the new engine."""
        with self._lock:
            if name in self._bound and name in self._vars:
                self._vars[name] = self._vars[name].detach().clone()
                self._bound.discard(name)

    def variable_names(self) -> List[str]:
        with self._lock:
            return sorted(self._vars)

    def snapshot(self, names: Optional[Iterable[str]] = None) -> Dict[str, torch.Tensor]:
        with self._lock:
            keys = self._vars.keys() if names is None else names
            return {k: self._vars[k].detach().to("cpu", copy=True) for k in keys if k in self._vars}

    def clear(self) -> None:
        with self._lock:
            self.__dict__.pop("_rng_streams", None)
            self._vars.clear()
            self._uninit.clear()
            self._bound.clear()
            for r in self.resources.values():
                close = getattr(r, "close", None)
                if close:
                    close()
            self.resources.clear()
            self.incarnation = time.time_ns()

    # -- generic resources -----------------------------------------------------------
    def get_resource(self, name: str, factory=None):
        with self._lock:
            r = self.resources.get(name)
            if r is None and factory is not None:
                r = self.resources[name] = factory()
            return r


class ExecContext:
    """Per-``Session.run`` state for one task."""

    def __init__(self, store: ResourceStore, task: Optional[Tuple[str, int]] = None, gpu_index: Optional[int] = None,
                 tracer=None, seed: Optional[int] = None, allow_soft_placement: bool = True):
        self.store = store
        self.task = task
        self.gpu_index = gpu_index            # CUDA ordinal bound to this task (None: use the spec's index)
        self.tracer = tracer
        self.values: Dict[int, Any] = {}
        self.leaves: Set[int] = set()         # node ids whose values must be autograd leaves
        self.allow_soft_placement = allow_soft_placement
        self._gen: Optional[torch.Generator] = None
        self.fusions = None                   # framework/fusion.py FusionState of this run (None: nothing planned)
        self._seed = seed
        self._dev_cache: Dict[str, torch.device] = {}
        self.force_device: Optional[torch.device] = None      # fabric strategy: run the whole sub-graph on the task's GPU

    def torch_device(self, node: Tensor) -> torch.device:
        if self.force_device is not None:
            return self.force_device
        key = node.device
        dev = self._dev_cache.get(key)
        if dev is None:
            spec = DeviceSpec.from_string(key)
            if spec.device_type == "GPU" and torch.cuda.is_available():
                idx = spec.device_index or 0
                if self.gpu_index is not None:
                    idx = self.gpu_index + idx
                dev = torch.device("cuda", idx % torch.cuda.device_count())
            elif spec.device_type == "GPU" and not self.allow_soft_placement:
                raise RuntimeError("cannot place %r on %s: no CUDA device" % (node.name, key))
            elif spec.device_type is None and self.gpu_index is not None and torch.cuda.is_available():
                dev = torch.device("cuda", self.gpu_index)      # task bound to a GPU: default device is that GPU
            else:
                dev = torch.device("cpu")
            self._dev_cache[key] = dev
        return dev

    def generator(self, node: Tensor) -> torch.Generator:
        if self._gen is None:
            self._gen = torch.Generator(device="cpu")
            seed = self._seed if self._seed is not None else node.graph.seed
            if seed is not None:
                self._gen.manual_seed(int(seed))
            else:
                self._gen.seed()
        return self._gen


def needed_nodes(fetch_nodes: Sequence[Tensor], fed: Set[int]) -> List[Tensor]:
    """Nodes reachable from the fetches (data + control edges), not expanding fed nodes; id order."""
    seen: Dict[int, Tensor] = {}
    stack = list(fetch_nodes)
    while stack:
        n = stack.pop()
        if n.id in seen:
            continue
        seen[n.id] = n
        if n.id in fed:
            continue
        stack.extend(n.inputs)
        stack.extend(n.control_inputs)
    # node ids are creation-ordered, and a node can only reference earlier nodes (or, for control
    # edges added post hoc by group(), explicitly listed ones) -> do a real topological sort.
    order: List[Tensor] = []
    state: Dict[int, int] = {}
    for root in sorted(seen.values(), key=lambda t: t.id):
        if state.get(root.id) == 2:
            continue
        st = [(root, iter(() if root.id in fed else list(root.inputs) + list(root.control_inputs)))]
        state[root.id] = 1
        while st:
            node, it = st[-1]
            advanced = False
            for dep in it:
                s = state.get(dep.id, 0)
                if s == 0:
                    state[dep.id] = 1
                    st.append((dep, iter(() if dep.id in fed else list(dep.inputs) + list(dep.control_inputs))))
                    advanced = True
                    break
                if s == 1:
                    raise ValueError("cycle in graph at %r" % dep.name)
            if not advanced:
                state[node.id] = 2
                order.append(node)
                st.pop()
    return order


def schedule_segments(order: Sequence[Tensor], task_of, fed: Set[int]) -> List[Tuple[Any, List[Tensor]]]:
    """Greedy list scheduling: keep executing ready nodes of the current task before switching,
    which yields the minimum number of cross-task hand-offs for chain-like graphs
    (ps reads -> worker fwd/bwd -> ps apply = three segments)."""
    order = [n for n in order if n.id not in fed]
    indeg: Dict[int, int] = {}
    users: Dict[int, List[Tensor]] = {}
    ids = {n.id for n in order}
    for n in order:
        deps = {d.id for d in list(n.inputs) + list(n.control_inputs) if d.id in ids}
        indeg[n.id] = len(deps)
        for d in deps:
            users.setdefault(d, []).append(n)
    ready: Dict[Any, List[Tensor]] = {}
    for n in order:
        if indeg[n.id] == 0:
            ready.setdefault(task_of(n), []).append(n)
    segments: List[Tuple[Any, List[Tensor]]] = []
    current = None
    remaining = len(order)
    while remaining:
        if current is None or not ready.get(current):
            # pick the task with the lowest-id ready node (stable, deterministic)
            current = min((t for t, l in ready.items() if l), key=lambda t: min(x.id for x in ready[t]))
            segments.append((current, []))
        bucket = ready[current]
        bucket.sort(key=lambda t: -t.id)
        n = bucket.pop()
        segments[-1][1].append(n)
        remaining -= 1
        for u in users.get(n.id, ()):
            indeg[u.id] -= 1
            if indeg[u.id] == 0:
                ready.setdefault(task_of(u), []).append(u)
    return segments


def _to_device(v, dev: torch.device):
    if isinstance(v, torch.Tensor) and v.device != dev:
        return v.to(dev, non_blocking=True)
    return v


def execute(nodes: Sequence[Tensor], ctx: ExecContext, want_grad: bool) -> None:
    """Evaluate ``nodes`` (already topologically ordered, inputs available in ``ctx.values``)."""
    values = ctx.values
    tracer = ctx.tracer
    fus = getattr(ctx, "fusions", None)
    fus_ids = fus.ids if fus else ()
    for node in nodes:
        if node.id in values:
            continue
        if node.id in fus_ids:
            # plan-time rewrites (framework/fusion.py): bias + ReLU in the GEMM epilogue, the clipped softmax cross-entropy chain
            # as one kernel; nodes interior to an active rewrite have no value of their own
            t0 = time.perf_counter_ns() if tracer is not None else 0
            handled, out = _fusion.try_execute(node, ctx, values, fus, ctx.torch_device(node), want_grad)
            if handled:
                if out is not None:
                    values[node.id] = out
                    if tracer is not None:
                        tracer.record(node, ctx, t0, time.perf_counter_ns(), out)
                continue
        kernel = _ops.KERNELS.get(node.op_type)
        if kernel is None:
            raise NotImplementedError("no kernel registered for op type %r (node %r)" % (node.op_type, node.name))
        args = []
        stateful = node.op_type in _ops.STATEFUL_OPS
        dev = None if stateful else ctx.torch_device(node)
        for i in node.inputs:
            v = values[i.id]
            args.append(v if dev is None else _to_device(v, dev))
        t0 = time.perf_counter_ns() if tracer is not None else 0
        if stateful or not want_grad:
            with torch.no_grad():
                out = kernel(ctx, node, *args)
        else:
            with torch.enable_grad():
                out = kernel(ctx, node, *args)
        if node.id in ctx.leaves and isinstance(out, torch.Tensor) and out.is_floating_point():
            out = out.detach().requires_grad_(True)
        values[node.id] = out
        if tracer is not None:
            tracer.record(node, ctx, t0, time.perf_counter_ns(), out)
