"""``Session``: the client handle that evaluates fetches.

* ``Session()`` -- single-process: one private task, job/task fields of device
  strings are ignored (reference ``standalone.py:30,136``,
  ``distributed_mnist_predict.py:35``).
* ``Session("grpc://host:port")`` -- distributed: the addressed task is the
  *master*; the client prunes the graph to what the fetches need, partitions
  it by ``/job/task`` and drives each task's segments (reference
  ``example_in_graph.py:49``, ``example_distributed_client.py:28``).  Tasks in
  the same process are called directly, others through the control-plane RPC.
* ``run(..., options=RunOptions(trace_level=FULL_TRACE), run_metadata=md)``
  collects per-op timing from every participating task into
  ``md.step_stats`` for :class:`dtf.timeline.Timeline` (reference
  ``example_in_graph.py:42-43,59``).
"""
from __future__ import annotations

import itertools
import sys
import threading
import uuid
from typing import Any, Dict, List, Optional, Sequence, Set, Tuple

import numpy as np
import torch

from ..framework import errors
from ..framework import ops as _ops
from ..framework.device import DeviceSpec
from ..framework.executor import ExecContext, ResourceStore, execute, needed_nodes, schedule_segments
from ..framework.graph import Graph, Tensor, get_default_graph
from ..framework.variables import Variable
from ..parallel.cluster import ClusterSpec
from ..parallel.rpc import RpcClient, parse_address
from ..parallel.server import Server, local_server_for, serialize_nodes

__all__ = ["Session", "InteractiveSession", "get_default_session", "RunOptions", "RunMetadata", "ConfigProto",
           "GPUOptions"]


class RunOptions:
    NO_TRACE, SOFTWARE_TRACE, HARDWARE_TRACE, FULL_TRACE = 0, 1, 2, 3

    def __init__(self, trace_level: int = 0, timeout_in_ms: int = 0):
        self.trace_level, self.timeout_in_ms = trace_level, timeout_in_ms


class RunMetadata:
    def __init__(self) -> None:
        self.step_stats: List[Dict[str, Any]] = []
        self.partition_graphs: List[Any] = []


class GPUOptions:
    """``per_process_gpu_memory_fraction`` caps this process's share of the 180 GB HBM3e
    (reference ``example_between_graph.py:76-79`` builds one and never uses it)."""

    def __init__(self, per_process_gpu_memory_fraction: float = 0.0, allow_growth: bool = True,
                 visible_device_list: str = ""):
        self.per_process_gpu_memory_fraction = per_process_gpu_memory_fraction
        self.allow_growth, self.visible_device_list = allow_growth, visible_device_list


class ConfigProto:
    def __init__(self, gpu_options: Optional[GPUOptions] = None, allow_soft_placement: bool = True,
                 log_device_placement: bool = False, device_filters: Optional[Sequence[str]] = None,
                 operation_timeout_in_ms: int = 0, intra_op_parallelism_threads: int = 0,
                 inter_op_parallelism_threads: int = 0):
        self.intra_op_parallelism_threads = int(intra_op_parallelism_threads)     # 0: the task's default budget
        self.inter_op_parallelism_threads = int(inter_op_parallelism_threads)     # accepted for parity (one executor thread)
        self.gpu_options = gpu_options or GPUOptions()
        self.allow_soft_placement, self.log_device_placement = allow_soft_placement, log_device_placement
        self.device_filters = list(device_filters or [])
        self.operation_timeout_in_ms = operation_timeout_in_ms

    def apply(self) -> None:
        if self.intra_op_parallelism_threads > 0:
            torch.set_num_threads(self.intra_op_parallelism_threads)
        frac = self.gpu_options.per_process_gpu_memory_fraction
        if frac and torch.cuda.is_available():
            for d in range(torch.cuda.device_count()):
                torch.cuda.set_per_process_memory_fraction(float(frac), d)


_tls = threading.local()


def _sess_stack() -> List["Session"]:
    st = getattr(_tls, "stack", None)
    if st is None:
        st = _tls.stack = []
    return st


def get_default_session() -> Optional["Session"]:
    st = _sess_stack()
    return st[-1] if st else None


class _Plan:
    __slots__ = ("fetch_nodes", "order", "segments", "leaves", "want_grad", "wanted", "consumers_task", "version", "fusions")


class Session:
    def __init__(self, target: str = "", graph: Optional[Graph] = None, config: Optional[ConfigProto] = None):
        self.graph = graph or get_default_graph()
        self.config = config or ConfigProto()
        self.config.apply()
        self.target = target or ""
        self._closed = False
        self.session_id = uuid.uuid4().hex
        self._graph_key = "%s/%d" % (self.session_id, id(self.graph))
        self._plans: Dict[Any, _Plan] = {}
        self._sent: Dict[Tuple[str, int], int] = {}       # remote task -> number of graph nodes already shipped
        self._clients: Dict[Tuple[str, int], RpcClient] = {}
        self._run_counter = itertools.count()
        self._lock = threading.RLock()
        if self.target:
            addr = "%s:%d" % parse_address(self.target)
            srv = local_server_for(addr)
            if srv is not None:
                self.cluster: Optional[ClusterSpec] = srv.cluster
                self.master_task: Optional[Tuple[str, int]] = srv.task
            else:
                info = RpcClient(addr).call("get_cluster")
                self.cluster = ClusterSpec(info["cluster"])
                self.master_task = tuple(info["task"])
            self._local_store = None
        else:
            self.cluster, self.master_task = None, None
            self._local_store = ResourceStore("local-session")
        self._default_ctx = None

    # -- context manager / default session -----------------------------------------------------
    def __enter__(self) -> "Session":
        _sess_stack().append(self)
        return self

    def __exit__(self, exc_type, exc, tb) -> bool:
        st = _sess_stack()
        if st and st[-1] is self:
            st.pop()
        self.close()
        return False

    def as_default(self):
        sess = self

        class _Ctx:
            def __enter__(self_inner):
                _sess_stack().append(sess)
                return sess

            def __exit__(self_inner, *a):
                _sess_stack().pop()
                return False
        return _Ctx()

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        if self.cluster is not None:
            for task in list(self._sent) + ([self.master_task] if self.master_task else []):
                self._call_best_effort(task, "release_session", self.session_id, self._graph_key)
        for c in self._clients.values():
            c.close()
        self._clients.clear()

    def cancel(self) -> None:
        """Unblock this session's pending blocking ops (token dequeue / take_grad) on every task."""
        if self.cluster is None:
            return
        for job, idx, _ in self.cluster.all_tasks():
            self._call_best_effort((job, idx), "cancel", self.session_id)

    def _call_best_effort(self, task, method, *args) -> None:
        """Clean-up broadcast (cancel / release): never waits for a task that is gone.  A normal call retries the
        connection for 30 s (tasks may still be starting); at shutdown a dead peer -- e.g. a killed backup worker --
        must not hold the survivors up."""
        try:
            srv = self._server(task)
            if srv is not None:
                getattr(srv, "rpc_" + method)(*args)
                return
            if task not in self._clients and not RpcClient(self.cluster.task_address(*task)).try_connect(timeout=0.3):
                return
            self._client(task).call(method, *args)
        except Exception:
            pass

    # -- task plumbing ------------------------------------------------------------------------------
    def _task_of(self, node: Tensor):
        if self.cluster is None:
            return None
        spec = DeviceSpec.from_string(node.device)
        if spec.job is None:
            return self.master_task
        return (spec.job, 0 if spec.task is None else spec.task)

    def _server(self, task) -> Optional[Server]:
        return local_server_for(self.cluster.task_address(*task))

    def _client(self, task) -> RpcClient:
        c = self._clients.get(task)
        if c is None:
            c = self._clients[task] = RpcClient(self.cluster.task_address(*task))
        return c

    def _call(self, task, method, *args):
        srv = self._server(task)
        if srv is not None:
            return getattr(srv, "rpc_" + method)(*args)
        return self._client(task).call(method, *args)

    # -- planning ----------------------------------------------------------------------------------
    def _resolve(self, f) -> Optional[Tensor]:
        if isinstance(f, Tensor):
            return f
        if isinstance(f, Variable):
            return f._node
        if isinstance(f, str):
            return self.graph.get_tensor_by_name(f)
        node = getattr(f, "_node", None)
        if node is not None:
            return node
        raise TypeError("cannot fetch %r (type %s)" % (f, type(f).__name__))

    def _flatten(self, fetches, out: List[Tensor]):
        """Returns a structure of indices into ``out`` mirroring ``fetches``."""
        if isinstance(fetches, (list, tuple)):
            return type(fetches)(self._flatten(f, out) for f in fetches) if not hasattr(fetches, "_fields") \
                else type(fetches)(*[self._flatten(f, out) for f in fetches])
        if isinstance(fetches, dict):
            return {k: self._flatten(v, out) for k, v in fetches.items()}
        out.append(self._resolve(fetches))
        return _Index(len(out) - 1)

    def _plan(self, fetch_nodes: List[Tensor], fed: Set[int]) -> _Plan:
        key = (tuple(n.id for n in fetch_nodes), frozenset(fed), self.graph.version)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        plan = _Plan()
        plan.fetch_nodes = fetch_nodes
        plan.order = needed_nodes(fetch_nodes, fed)
        plan.leaves = set()
        plan.want_grad = False
        for n in plan.order:
            if n.op_type == "Gradients":
                plan.want_grad = True
                ny = n.attrs["num_ys"]
                plan.leaves.update(x.id for x in n.inputs[ny:])
        if self.cluster is None:
            plan.segments = [(None, [n for n in plan.order if n.id not in fed])]      # one process: the whole plan is one segment
        else:
            plan.segments = schedule_segments(plan.order, self._task_of, fed)
        # which node values must leave their producing task: fetches + cross-task consumers
        seg_of: Dict[int, int] = {}
        for si, (_, nodes) in enumerate(plan.segments):
            for n in nodes:
                seg_of[n.id] = si
        wanted: List[Set[int]] = [set() for _ in plan.segments]
        for f in fetch_nodes:
            if f.id in seg_of:
                wanted[seg_of[f.id]].add(f.id)
        for si, (task, nodes) in enumerate(plan.segments):
            for n in nodes:
                for d in n.inputs:
                    sj = seg_of.get(d.id)
                    if sj is not None and plan.segments[sj][0] != task:
                        wanted[sj].add(d.id)
        plan.wanted = wanted
        from ..framework import fusion as _fusion
        task_of_id = {nid: plan.segments[si][0] for nid, si in seg_of.items()}
        plan.fusions = _fusion.plan_fusions(plan.order, {f.id for f in fetch_nodes}, plan.leaves, task_of_id, fed)
        if getattr(self.config, "log_device_placement", False):
            # TF logs the placement of every node the first time it is part of a run
            logged = self.__dict__.setdefault("_placement_logged", set())
            for n in plan.order:
                if n.id not in logged:
                    logged.add(n.id)
                    task = self._task_of(n)
                    where = n.device or ("/job:%s/task:%d" % task if task else "/job:localhost/replica:0/task:0/device:CPU:0")
                    print("%s: (%s): %s" % (n.name, n.op_type, where), file=sys.stderr, flush=True)
        if len(self._plans) > 256:
            self._plans.clear()
        self._plans[key] = plan
        return plan

    # -- run ------------------------------------------------------------------------------------------
    def run(self, fetches, feed_dict: Optional[Dict[Any, Any]] = None, options: Optional[RunOptions] = None,
            run_metadata: Optional[RunMetadata] = None):
        if self._closed:
            raise RuntimeError("Attempted to use a closed Session.")
        flat: List[Tensor] = []
        structure = self._flatten(fetches, flat)
        feeds: Dict[int, torch.Tensor] = {}
        if feed_dict:
            for k, v in feed_dict.items():
                node = self._resolve(k)
                t = v if isinstance(v, torch.Tensor) else _ops._to_torch(v, node.dtype if node.dtype is not None
                                                                           and not isinstance(v, torch.Tensor) else None)
                if node.dtype is not None and t.dtype != node.dtype and t.is_floating_point() == node.dtype.is_floating_point:
                    t = t.to(node.dtype)
                if node.shape is not None and t.dim() == len(node.shape):
                    for want, got in zip(node.shape, t.shape):
                        if want is not None and want != got:
                            raise ValueError("Cannot feed value of shape %s for Tensor %r, which has shape %s"
                                             % (tuple(t.shape), node.name, node.shape))
                feeds[node.id] = t
        trace = options is not None and options.trace_level != RunOptions.NO_TRACE
        # Fetch overrides (parallel/auto_fabric.py): a tensor whose value a fused engine step already produced -- the loss
        # of a FabricTrainStep fetched in the same run, or fetched alone for validation -- is answered by the engine instead
        # of being recomputed through the graph.  An override returns NotImplemented to decline.
        overrides = getattr(self.graph, "_fetch_overrides", None)
        taken: Dict[int, Any] = {}
        normal = flat
        if overrides:
            normal = [n for n in flat if n.id not in overrides]
        results: Dict[int, Any] = {}
        if normal or not flat:
            plan = self._plan(normal, set(feeds))
            results = self._execute_plan(plan, feeds, trace, run_metadata)
        if overrides and len(normal) != len(flat):
            fetch_ids = {n.id for n in flat}
            declined = []
            for n in flat:
                if n.id in overrides and n.id not in taken:
                    v = overrides[n.id](self, feeds, fetch_ids)
                    if v is NotImplemented:
                        declined.append(n)
                    else:
                        taken[n.id] = v
            if declined:
                plan2 = self._plan(declined, set(feeds))
                results = dict(results)
                results.update(self._execute_plan(plan2, feeds, trace, run_metadata))
        values = [_to_numpy(taken[n.id] if n.id in taken else results.get(n.id)) for n in flat]
        return _unflatten(structure, values)

    def _execute_plan(self, plan: _Plan, feeds: Dict[int, torch.Tensor], trace: bool,
                      run_metadata: Optional[RunMetadata]) -> Dict[int, Any]:
        # ---- single-process fast path ----
        if self.cluster is None:
            tracer = None
            if trace:
                from ..utils.timeline import StepTracer
                tracer = StepTracer("/job:localhost/replica:0/task:0")
            ctx = ExecContext(self._local_store, None, None, tracer, None, self.config.allow_soft_placement)
            ctx.leaves = plan.leaves
            if plan.fusions:
                from ..framework.fusion import FusionState
                ctx.fusions = FusionState(plan.fusions)
            ctx.cancel_event = threading.Event()
            for nid, v in feeds.items():
                ctx.values[nid] = v.detach().requires_grad_(True) if (nid in plan.leaves and v.is_floating_point()) else v
            for _, nodes in plan.segments:
                execute(nodes, ctx, plan.want_grad)
            if trace and run_metadata is not None:
                run_metadata.step_stats.extend(tracer.events())
            return ctx.values

        # ---- distributed: master drives per-task segments ----
        run_id = "%s-%d" % (self.session_id, next(self._run_counter))
        opts = {"leaves": list(plan.leaves), "want_grad": plan.want_grad, "trace": trace,
                "session_id": self.session_id, "graph_seed": self.graph.seed, "fusions": plan.fusions}
        master_values: Dict[int, Any] = dict(feeds)
        touched: List[Tuple[str, int]] = []
        last_segment = {task: si for si, (task, _) in enumerate(plan.segments)}     # where each task's run state can go
        try:
            for si, (task, nodes) in enumerate(plan.segments):
                if task not in touched:
                    touched.append(task)
                seg_ids = {n.id for n in nodes}
                inputs: Dict[int, Any] = {}
                for n in nodes:
                    for d in n.inputs:
                        if d.id not in seg_ids and d.id in master_values:
                            inputs[d.id] = master_values[d.id]
                want = sorted(plan.wanted[si])
                srv = self._server(task)
                if srv is not None:
                    out = srv.run_segment_local(run_id, nodes, inputs, want, opts)
                    # a value leaving its task is a COPY (TF's Send/Recv; what the RPC path does by serialising):
                    # without it a variable read would alias the ps's storage and a concurrent apply (the chief's
                    # aggregation thread, another worker) could modify it between this step's forward and backward
                    out = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}
                else:
                    sent = self._sent.get(task, 0)
                    all_nodes = self.graph.nodes
                    new_defs = serialize_nodes(all_nodes[sent:], lambda n, t=task: self._task_of(n) == t) \
                        if sent < len(all_nodes) else []
                    finish = last_segment[task] == si
                    out = self._client(task).call("run_segment", run_id, self._graph_key, new_defs,
                                                  [n.id for n in nodes], inputs, want, opts, finish)
                    self._sent[task] = len(all_nodes)
                    if finish:                       # the task released the run state in the same round trip
                        touched.remove(task)
                        if trace and run_metadata is not None and out["events"]:
                            run_metadata.step_stats.extend(out["events"])
                        out = out["values"]
                master_values.update(out)
        finally:
            for task in touched:
                try:
                    ev = self._call(task, "end_run", run_id)
                    if trace and run_metadata is not None and ev:
                        run_metadata.step_stats.extend(ev)
                except Exception:
                    pass
        return master_values

    # -- misc TF surface ----------------------------------------------------------------------------------
    def list_devices(self) -> List[str]:
        if self.cluster is None:
            devs = ["/job:localhost/replica:0/task:0/device:CPU:0"]
            devs += ["/job:localhost/replica:0/task:0/device:GPU:%d" % i for i in range(torch.cuda.device_count())]
            return devs
        return ["/job:%s/task:%d" % (j, i) for j, i, _ in self.cluster.all_tasks()]

    @property
    def sess_str(self) -> str:
        return self.target

    def local_store(self) -> Optional[ResourceStore]:
        if self._local_store is not None:
            return self._local_store
        srv = self._server(self.master_task) if self.master_task else None
        return srv.store if srv is not None else None


class InteractiveSession(Session):
    def __init__(self, target: str = "", graph=None, config=None):
        super().__init__(target, graph, config)
        _sess_stack().append(self)

    def close(self) -> None:
        st = _sess_stack()
        if st and st[-1] is self:
            st.pop()
        super().close()


class _Index:
    __slots__ = ("i",)

    def __init__(self, i: int):
        self.i = i


def _unflatten(structure, values):
    if isinstance(structure, _Index):
        return values[structure.i]
    if isinstance(structure, dict):
        return {k: _unflatten(v, values) for k, v in structure.items()}
    if hasattr(structure, "_fields"):
        return type(structure)(*[_unflatten(s, values) for s in structure])
    return type(structure)(_unflatten(s, values) for s in structure)


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        t = v.detach()
        if t.dtype == torch.bfloat16:
            t = t.float()
        a = t.cpu().numpy()
        if a.ndim == 0:
            return a[()]
        # a fetched value is the caller's own (TF copies into the result): a host tensor's numpy view would alias the live
        # variable / an executor value and change under the caller at the next apply
        return a.copy() if t.device.type == "cpu" else a
    if isinstance(v, list) and v and all(isinstance(x, torch.Tensor) or x is None for x in v):
        return [_to_numpy(x) for x in v]
    return v
