"""In-process task server (``dtf.train.Server``).

Capability parity: ``Server(cluster, job_name, task_index)`` starts
this task's service bound to its ClusterSpec address, ``.target`` is the
string a ``Session`` connects to and ``.join()`` blocks forever (the ps role) --
reference ``distributed_mnist.py:75-79``, ``example_between_graph.py:33-40``,
``example_in_graph.py:30,46``.  Every variable placed on the task lives in the
server's :class:`ResourceStore` and outlives client sessions.

The server executes graph *segments* that a session's master hands it
(:meth:`run_segment`); a segment is a topologically ordered list of nodes whose
device names this task.  Per-run state (values, autograd tape) stays on the
task between segments of the same run, so forward and backward of one
``Session.run`` can be separated by a ps round trip.
"""
from __future__ import annotations

import atexit
import os
import socket
import sys
import threading
import time
import weakref
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..framework import errors
from ..framework.executor import ExecContext, ResourceStore, execute
from .cluster import ClusterSpec
from .rpc import PeerAwareCancel, RpcClient, RpcServer, current_connection, parse_address, peer_closed

__all__ = ["Server", "NodeView", "serialize_nodes", "local_server_for", "local_servers"]

_LOCAL_SERVERS: Dict[Tuple[str, int], "weakref.ReferenceType[Server]"] = {}
_LOCAL_LOCK = threading.Lock()


def local_server_for(address: str) -> Optional["Server"]:
    try:
        key = parse_address(address)
    except ValueError:
        return None
    with _LOCAL_LOCK:
        ref = _LOCAL_SERVERS.get(key)
    srv = ref() if ref is not None else None
    return srv if srv is not None and srv.is_running else None


def local_servers() -> List["Server"]:
    """Every running Server of this process (a between-graph task normally has exactly one)."""
    with _LOCAL_LOCK:
        refs = list(_LOCAL_SERVERS.values())
    out = []
    for ref in refs:
        srv = ref()
        if srv is not None and srv.is_running:
            out.append(srv)
    return out


def _stop_local_servers_at_exit() -> None:
    """Registered with ``atexit`` (runs BEFORE the interpreter finalises): a task script that simply returns -- the
    reference's worker-0 client does, ``example_distributed_server.py:46-70`` -- leaves its in-process Server's accept and
    connection threads blocked in native socket calls.  CPython ends a daemon thread that comes back from such a call during
    finalisation with ``pthread_exit``; that forced unwind racing the process's static destructors showed up as a rare
    ``terminate called without an active exception`` (exit -6) AFTER the script's output was complete.  Closing the
    listeners / connections here lets every one of those threads return and be joined while the interpreter is whole."""
    # a peer's client session may still be using this task (in-graph replication: the reference's example_distributed_server.py
    # makes EVERY worker a client of every other one): leave only once those sessions were released or their clients are gone,
    # within DTF_EXIT_LINGER_S (default 10 s) -- a task that simply drops out fails its peers' in-flight runs with UnavailableError
    from .rpc import close_all_clients
    close_all_clients()           # this process's own client connections: sessions it never closed are not "peers still using us"
    try:
        linger = float(os.environ.get("DTF_EXIT_LINGER_S", "10"))
    except ValueError:
        linger = 10.0
    # DTF_EXIT_SERVE_S: keep serving for at least this long after the script returned (a task whose peers may not have reached it
    # yet: every worker of example_distributed_server.py is a client that exits when ITS run is done)
    try:
        serve_s = float(os.environ.get("DTF_EXIT_SERVE_S", "0"))
    except ValueError:
        serve_s = 0.0
    if serve_s > 0 and local_servers():
        t_end = time.time() + serve_s
        while time.time() < t_end and local_servers():
            time.sleep(0.1)
    deadline = time.time() + max(0.0, linger)
    while time.time() < deadline:
        busy = False
        for srv in local_servers():
            for sid, conn in list(srv._remote_sessions.items()):
                if getattr(conn, "closed", False) or peer_closed(conn):
                    srv._remote_sessions.pop(sid, None)
                else:
                    busy = True
        if not busy:
            break
        time.sleep(0.05)
    for srv in local_servers():
        try:
            rpc = srv._rpc
            if rpc is not None:
                rpc.close()
        except Exception:      # noqa: BLE001 - exit path
            pass
    deadline = time.time() + 2.0
    for t in threading.enumerate():
        if t.daemon and t is not threading.current_thread() and t.name.startswith("dtf-rpc"):
            t.join(max(0.0, deadline - time.time()))


atexit.register(_stop_local_servers_at_exit)


class _GraphStub:
    __slots__ = ("seed",)

    def __init__(self, seed=None):
        self.seed = seed


class NodeView:
    """Deserialised node: the attribute subset kernels read."""
    __slots__ = ("id", "name", "op_type", "inputs", "control_inputs", "attrs", "device", "dtype", "shape", "graph")

    def __init__(self, d: Dict[str, Any], graph: _GraphStub):
        self.id, self.name, self.op_type = d["id"], d["name"], d["op_type"]
        self.attrs, self.device, self.dtype, self.shape = d["attrs"], d["device"], d["dtype"], d["shape"]
        self.inputs: List[Any] = d["inputs"]           # ids until linked
        self.control_inputs: List[Any] = d["control_inputs"]
        self.graph = graph


def serialize_nodes(nodes: Sequence[Any], runs_here=None) -> List[Dict[str, Any]]:
    """Wire form of graph nodes for one remote task.  Attributes travel only for nodes that task will execute
    (``runs_here(node)``); the others are structure-only stubs (their attrs may hold client-side objects)."""
    out = []
    for n in nodes:
        keep = runs_here is None or runs_here(n)
        out.append({"id": n.id, "name": n.name, "op_type": n.op_type, "inputs": [i.id for i in n.inputs],
                    "control_inputs": [c.id for c in n.control_inputs], "attrs": dict(n.attrs) if keep else {},      # tensors inside are encoded by RpcClient.call
                    "device": n.device, "dtype": n.dtype, "shape": n.shape})
    return out


class _RunState:
    __slots__ = ("ctx", "want_grad", "created")

    def __init__(self, ctx: ExecContext, want_grad: bool):
        self.ctx, self.want_grad, self.created = ctx, want_grad, time.time()


class Server:
    def __init__(self, server_or_cluster_def, job_name: Optional[str] = None, task_index: Optional[int] = None,
                 protocol: Optional[str] = None, config=None, start: bool = True, gpu_index: Optional[int] = None):
        self.cluster = server_or_cluster_def if isinstance(server_or_cluster_def, ClusterSpec) \
            else ClusterSpec(server_or_cluster_def)
        if job_name is None:
            if len(self.cluster.jobs) != 1:
                raise ValueError("job_name is required when the cluster has several jobs")
            job_name = self.cluster.jobs[0]
        self.job_name = job_name
        self.task_index = int(task_index or 0)
        self.address = self.cluster.task_address(self.job_name, self.task_index)
        self.store = ResourceStore("/job:%s/task:%d" % (self.job_name, self.task_index))
        self.config = config
        # One process per GPU: DTF_GPU_INDEX (or LOCAL_RANK under torchrun) binds the task to a B200.
        if gpu_index is None:
            env = os.environ.get("DTF_GPU_INDEX")
            if env is not None and env != "":
                gpu_index = int(env)
        self.gpu_index = gpu_index if (gpu_index is not None and gpu_index >= 0 and torch.cuda.is_available()) else None
        self._set_intra_op_threads(config)
        self._graphs: Dict[str, Dict[int, NodeView]] = {}
        self._runs: Dict[str, _RunState] = {}
        self._cancel: Dict[str, threading.Event] = {}
        self._lock = threading.RLock()
        self._rpc: Optional[RpcServer] = None
        self._stopped = threading.Event()
        self._peers: Dict[Tuple[str, int], RpcClient] = {}
        # a task serves several peers from Python threads: a handler that becomes runnable (its reply is due) should not wait a
        # whole default GIL switch interval (5 ms) behind another handler's bookkeeping -- measured on the sync MNIST example
        # (1 ps + 2 workers): DTF_GIL_SWITCH_US=0 keeps the interpreter default
        us = float(os.environ.get("DTF_GIL_SWITCH_US", "200"))
        if us > 0 and sys.getswitchinterval() > us * 1e-6:
            sys.setswitchinterval(us * 1e-6)
        self._remote_sessions: Dict[str, Any] = {}           # open sessions of REMOTE clients that ran something here -> their connection
        self._fabric_gen: Dict[str, Dict[str, Any]] = {}      # fabric incarnations (rpc_fabric_generation); ps task 0 is asked
        self._fabric_gen_lock = threading.Lock()
        if start:
            self.start()

    def _set_intra_op_threads(self, config) -> None:
        """CPU thread budget of this task's kernels.  Several tasks of a cluster usually share one machine (the
        reference's localhost ps/worker layout, ``distributed_mnist.py:27-29``); if each of them ran the default
        one-thread-per-core pool the pools would thrash (measured: 18 vs 77 sync steps/s for 1 ps + 2 workers on
        8 cores).  Default: cores / tasks-on-this-host, at most 4 (the per-step tensors are small);
        ``ConfigProto(intra_op_parallelism_threads=n)`` or ``DTF_INTRA_OP_THREADS`` override it."""
        n = int(getattr(config, "intra_op_parallelism_threads", 0) or 0) if config is not None else 0
        if n <= 0:
            n = int(os.environ.get("DTF_INTRA_OP_THREADS", "0") or 0)
        if n <= 0:
            if os.environ.get("OMP_NUM_THREADS"):
                return                       # the user already chose
            local = {"127.0.0.1", "localhost", "0.0.0.0", "", socket.gethostname()}
            mine = parse_address(self.address)[0]
            tasks = 0
            for job in self.cluster.jobs:
                for t in self.cluster.task_indices(job):
                    host = parse_address(self.cluster.task_address(job, t))[0]
                    if host == mine or (host in local and mine in local):
                        tasks += 1
            n = max(1, min(4, (os.cpu_count() or 1) // max(tasks, 1)))
        try:
            torch.set_num_threads(n)
        except RuntimeError:
            pass
        self.intra_op_threads = n

    # -- lifecycle ----------------------------------------------------------------------
    def start(self) -> None:
        if self._rpc is not None:
            return
        self._rpc = RpcServer(self.address, self)
        with _LOCAL_LOCK:
            _LOCAL_SERVERS[parse_address(self.address)] = weakref.ref(self)

    @property
    def is_running(self) -> bool:
        return self._rpc is not None and not self._stopped.is_set()

    @property
    def target(self) -> str:
        return "grpc://%s" % self.address

    @property
    def task(self) -> Tuple[str, int]:
        return (self.job_name, self.task_index)

    def join(self, timeout: Optional[float] = None) -> None:
        """Block until :meth:`stop` (never, for a ps task started from the command line)."""
        self._stopped.wait(timeout)

    def stop(self) -> None:
        self._stopped.set()
        for ev in list(self._cancel.values()):
            ev.set()
        self.store.clear()
        if self._rpc is not None:
            self._rpc.close()
            self._rpc = None
        with _LOCAL_LOCK:
            _LOCAL_SERVERS.pop(parse_address(self.address), None)
        for c in self._peers.values():
            c.close()
        self._peers.clear()

    def __del__(self):
        try:
            if self._rpc is not None:
                self._rpc.close()
        except Exception:
            pass

    # -- segment execution (called in-process or through rpc_run_segment) ----------------------
    def cancel_event(self, session_id: str) -> threading.Event:
        with self._lock:
            ev = self._cancel.get(session_id)
            if ev is None:
                ev = self._cancel[session_id] = threading.Event()
            return ev

    def _run_state(self, run_id: str, opts: Dict[str, Any]) -> _RunState:
        with self._lock:
            st = self._runs.get(run_id)
            if st is None:
                tracer = None
                if opts.get("trace"):
                    from ..utils.timeline import StepTracer
                    tracer = StepTracer("/job:%s/task:%d" % self.task)
                ctx = ExecContext(self.store, self.task, self.gpu_index, tracer, opts.get("seed"))
                ctx.leaves = set(opts.get("leaves", ()))
                if opts.get("fusions"):
                    from ..framework.fusion import FusionState
                    ctx.fusions = FusionState(opts["fusions"])
                ctx.cancel_event = self.cancel_event(opts.get("session_id", ""))
                ctx.server = self
                st = self._runs[run_id] = _RunState(ctx, bool(opts.get("want_grad")))
            return st

    def run_segment_local(self, run_id: str, nodes: Sequence[Any], inputs: Dict[int, Any], want_ids: Sequence[int],
                          opts: Dict[str, Any]) -> Dict[int, Any]:
        if self._stopped.is_set():
            raise errors.AbortedError("server %s was stopped" % self.address)
        st = self._run_state(run_id, opts)
        ctx = st.ctx
        conn = current_connection()
        if conn is not None:         # remote client: blocking kernels also give up when that client dies
            ctx.cancel_event = PeerAwareCancel(self.cancel_event(opts.get("session_id", "")), conn)
            if opts.get("session_id"):
                self._remote_sessions[opts["session_id"]] = conn
        dev_default = None
        for nid, v in inputs.items():
            if isinstance(v, torch.Tensor):
                if nid in ctx.leaves and v.is_floating_point():
                    v = v.detach().requires_grad_(True)
            ctx.values[nid] = v
        execute(nodes, ctx, st.want_grad)
        out = {}
        for nid in want_ids:
            v = ctx.values.get(nid)
            out[nid] = v.detach() if isinstance(v, torch.Tensor) else v
        return out

    def end_run_local(self, run_id: str):
        with self._lock:
            st = self._runs.pop(run_id, None)
        if st is None or st.ctx.tracer is None:
            return None
        return st.ctx.tracer.events()

    # -- rpc surface -----------------------------------------------------------------------
    def rpc_ping(self):
        return {"task": self.task, "incarnation": self.store.incarnation, "pid": os.getpid()}

    def rpc_get_cluster(self):
        return {"cluster": self.cluster.as_dict(), "task": self.task}

    def rpc_run_segment(self, run_id, graph_key, new_nodedefs, node_ids, inputs, want_ids, opts, finish=False):
        """``finish``: this is the run's last segment on this task -- release the run state in the same round trip
        (the reply then is ``{"values": ..., "events": trace events or None}``), saving one RPC per step per task."""
        with self._lock:
            g = self._graphs.setdefault(graph_key, {})
            stub = _GraphStub(opts.get("graph_seed"))
            fresh = []
            for d in new_nodedefs:
                if d["id"] not in g:
                    nv = NodeView(d, stub)
                    g[nv.id] = nv
                    fresh.append(nv)
            for nv in fresh:
                nv.inputs = [g[i] for i in nv.inputs]
                nv.control_inputs = [g[i] for i in nv.control_inputs if i in g]
            nodes = [g[i] for i in node_ids]
        if not finish:
            return self.run_segment_local(run_id, nodes, inputs, want_ids, opts)
        try:
            values = self.run_segment_local(run_id, nodes, inputs, want_ids, opts)
        except BaseException:
            self.end_run_local(run_id)
            raise
        return {"values": values, "events": self.end_run_local(run_id)}

    def rpc_end_run(self, run_id):
        return self.end_run_local(run_id)

    def rpc_cancel(self, session_id):
        self.cancel_event(session_id).set()
        return True

    def rpc_release_session(self, session_id, graph_key=None):
        self._remote_sessions.pop(session_id, None)
        with self._lock:
            self._cancel.pop(session_id, None)
            if graph_key is not None:
                self._graphs.pop(graph_key, None)
        return True

    def rpc_snapshot(self, names=None):
        return self.store.snapshot(names)

    def rpc_variable_names(self):
        return self.store.variable_names()

    def rpc_restore(self, values: Dict[str, torch.Tensor]):
        for k, v in values.items():
            dev = None
            if self.gpu_index is not None:
                dev = torch.device("cuda", self.gpu_index)
            self.store.assign(k, v if dev is None else v.to(dev), device=dev)
        return sorted(values)

    def rpc_fabric_setup(self, spec):
        from .strategy import ps_fabric_setup
        return ps_fabric_setup(self, spec)

    def rpc_fabric_farewell(self, key):
        """A sync replica finished its training loop: release every worker's device-side token wait for good, so a replica
        still inside a step (whose aggregate would need the departed one's gradient) completes it, reads the final
        global step and stops too (the control-plane tier's farewell tokens, train/sync_replicas.py)."""
        svc = self.store.resources.get("fabric_service/" + key)
        if svc is not None:
            svc.farewell()
        return True

    def rpc_fabric_generation(self, base_key, at_least=0, rank=None):
        """Which incarnation of the fabric (buffers, rendezvous, apply service) the tasks of the job ``base_key`` should be
        in.  ps task 0 is the authority.  ``at_least``: a task that saw a failure at generation g asks for g + 1; every
        survivor of the same failure asks for the same number.  ``rank``: a task that asks to join a generation it is ALREADY
        a member of is a restarted process (its old incarnation's mappings are gone) -> the generation moves on and the
        surviving members learn it from their periodic liveness call."""
        with self._fabric_gen_lock:
            st = self._fabric_gen.setdefault(base_key, {"gen": 0, "members": set()})
            if int(at_least) > st["gen"]:
                st["gen"], st["members"] = int(at_least), set()
            elif rank is not None and rank in st["members"]:
                st["gen"], st["members"] = st["gen"] + 1, set()
            if rank is not None:
                st["members"].add(rank)
            return st["gen"]

    def rpc_fabric_current_generation(self, base_key):
        with self._fabric_gen_lock:
            st = self._fabric_gen.get(base_key)
            return st["gen"] if st else 0

    def rpc_fabric_teardown(self, base_key):
        """Stop this task's apply service(s) of the job ``base_key``: the graph variables keep their values in storage of
        their own (``VariableStore.unbind``), the engine's HBM and peer mappings are released."""
        from .strategy import ps_fabric_teardown
        return ps_fabric_teardown(self, base_key)

    def rpc_fabric_step_stats(self, key):
        """Timeline events of this ps task's fabric engine (device-side ``ps_apply`` ring): what a traced ``Session.run`` on
        a worker merges into ``RunMetadata.step_stats`` next to its own step-kernel phases."""
        svc = self.store.resources.get("fabric_service/" + key)
        if svc is None or svc.engine is None or not hasattr(svc.engine, "step_stats"):
            return []
        return svc.engine.step_stats()

    def rpc_reset(self):
        self.store.clear()
        return True

    def rpc_shutdown(self):
        threading.Thread(target=self.stop, daemon=True).start()
        return True

    # -- peers --------------------------------------------------------------------------------
    def peer(self, job: str, task: int) -> RpcClient:
        key = (job, int(task))
        c = self._peers.get(key)
        if c is None:
            c = self._peers[key] = RpcClient(self.cluster.task_address(job, task))
        return c

    @staticmethod
    def create_local_server(config=None, start=True) -> "Server":
        import socket
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        return Server({"local": ["127.0.0.1:%d" % port]}, job_name="local", task_index=0, config=config, start=start)
