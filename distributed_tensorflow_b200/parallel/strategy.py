"""``FabricPSStrategy``: run a tf.train-style ps/worker program (graph API, ``MonitoredTrainingSession``, hooks,
``Saver``) with the parameters, gradients and tokens on the NVLink fabric instead of the control-plane RPC.

What changes for the user program (``examples/distributed_mnist.py --engine=fabric``)::

    server = dtf.train.Server(cluster, job_name, task_index)
    strategy = dtf.fabric.FabricPSStrategy(server)               # ps tasks: then just server.join() as usual
    with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device=...)):
        ... build the model exactly as before ...
        train_op, loss = strategy.minimize(opt, loss, global_step)   # instead of opt.minimize(...)
    with dtf.train.MonitoredTrainingSession(master=server.target, is_chief=..., hooks=hooks) as sess:
        sess.run([train_op, global_step, loss], feed_dict={...})

How it works:

* ``minimize`` lists the trainable variables (name, shape, ps task from their device strings), asks every ps
  ``Server`` over the control plane to join the fabric (``rpc_fabric_setup``) and builds a
  :class:`GenericPSEngine` on the worker; the ps side builds the same engine, **binds each graph variable (and
  ``global_step``) to its slice of the engine's HBM buffers** in the task's resource store -- so initialisers,
  ``Saver.save/restore``, hooks reading ``global_step`` all keep working through the ordinary graph path -- and starts
  the apply service loop (one fused ``ps_apply`` kernel per aggregate / per push).
* ``train_op`` is a ``FabricTrainStep`` node: per ``Session.run`` it waits for the worker's token, pulls the
  parameters from the ps GPUs (peer loads), evaluates the user's loss sub-graph locally with those parameters as
  autograd leaves (our sm_100a kernels through ``ops/native.py``), and pushes the gradients into its slot on the ps
  (NVLink stores + stamp + arrival counter).  ``SyncReplicasOptimizer(opt, R, N)`` selects sync mode with
  ``replicas_to_aggregate=R``; a plain optimizer selects async mode (staleness measured on the ps).

Ranks are the ClusterSpec tasks in ``all_tasks()`` order (ps first), rendezvous happens over
``torch.distributed`` (gloo: ranks may share a GPU) at the first task's host, port ``+ DTF_FABRIC_PORT_OFFSET``.
"""
from __future__ import annotations

import os
import threading
import time
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch

from ..framework import device as _device
from ..framework import fusion as _fusion
from ..framework import ops as _ops
from ..framework.device import DeviceSpec
from ..framework.executor import ExecContext, execute, needed_nodes
from ..framework.graph import Tensor, convert_to_tensor, get_default_graph
from ..framework.ops import register_kernel
from ..framework.variables import Variable, trainable_variables
from .cluster import ClusterSpec

__all__ = ["FabricPSStrategy", "fabric_rank_of", "init_fabric_process_group"]

_PG_LOCK = threading.Lock()
_DEBUG = os.environ.get("DTF_DEBUG", "0") == "1"


def _dbg(msg: str) -> None:
    if _DEBUG:
        print("[dtf.fabric %.3f pid %d] %s" % (time.time() % 1000, os.getpid(), msg), flush=True)


def fabric_rank_of(cluster: ClusterSpec, job: str, task: int) -> Tuple[int, int]:
    tasks = [(j, i) for j, i, _ in cluster.all_tasks()]
    return tasks.index((job, int(task))), len(tasks)


_PG_GEN: List[Optional[int]] = [None]       # generation of the process group THIS module created (None: not ours / none)


def init_fabric_process_group(cluster: ClusterSpec, job: str, task: int, generation: int = 0) -> Tuple[int, int]:
    """Join the fabric's torch.distributed group of ``generation`` (idempotent).  gloo: only a store + barriers are needed.
    A new generation (the fabric re-forms after a task failure, ``FabricPSStrategy._abort``) leaves the old group -- it has
    a dead member -- and rendezvouses on its own port (``+ generation``) with whoever is alive or restarted by then."""
    import torch.distributed as dist
    rank, world = fabric_rank_of(cluster, job, task)
    with _PG_LOCK:
        if dist.is_initialized() and _PG_GEN[0] is not None and _PG_GEN[0] != generation:
            _dbg("leaving process group of generation %d" % _PG_GEN[0])
            try:
                dist.destroy_process_group()
            except Exception:      # noqa: BLE001 - peers of the old group may be gone
                pass
            _PG_GEN[0] = None
        if not dist.is_initialized():
            host, _, port = cluster.all_tasks()[0][2].rpartition(":")
            if host in ("localhost", "", "0.0.0.0"):
                host = "127.0.0.1"
            port = int(port) + int(os.environ.get("DTF_FABRIC_PORT_OFFSET", "1000")) + int(generation)
            import datetime
            _dbg("joining process group rank %d/%d at %s:%d (generation %d)" % (rank, world, host, port, generation))
            dist.init_process_group("gloo", init_method="tcp://%s:%d" % (host, port), rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=float(os.environ.get("DTF_FABRIC_TIMEOUT", "300"))))
            _PG_GEN[0] = generation
            _dbg("process group up")
    return rank, world


def _engine_cfg(spec: Dict[str, Any], num_ps: int, num_workers: int):
    from .ps_engine import EngineConfig
    opt = dict(spec["optimizer"])
    return EngineConfig(num_ps=num_ps, num_workers=num_workers, sync=bool(opt.get("sync", False)),
                        replicas_to_aggregate=opt.get("replicas_to_aggregate"), optimizer=opt,
                        timeout_ns=int(spec.get("timeout_ns") or
                                       float(os.environ.get("DTF_FABRIC_STEP_TIMEOUT", "20")) * 1e9))


def build_engine(cluster: ClusterSpec, job: str, task: int, spec: Dict[str, Any], gpu_index: Optional[int]):
    from .fabric import Fabric
    from .generic_engine import GenericPSEngine
    rank, world = init_fabric_process_group(cluster, job, task, int(spec.get("generation", 0)))
    dev = gpu_index if gpu_index is not None else 0
    torch.cuda.set_device(dev)
    import torch.distributed as dist
    fabric = Fabric(world, {rank: dev}, store=dist.distributed_c10d._get_default_store(),
                    prefix="dtf_fabric/%s" % spec["key"])
    num_ps = cluster.num_tasks("ps")
    cfg = _engine_cfg(spec, num_ps, world - num_ps)
    _dbg("building engine (rank %d, device %d)" % (rank, dev))
    if spec.get("mlp"):
        # the reference network: ONE fused step kernel per worker step, ONE apply kernel per aggregate (ps_engine.py)
        from .ps_engine import MLPSpec, PSTrainEngine
        m = spec["mlp"]
        shard_of = {n: sh for n, _, sh in spec["params"]}
        cfg.shards = {role: shard_of[m["roles"][role]] for role in ("hid_w", "hid_b", "sm_w", "sm_b")}
        cfg.precision, cfg.nvls, cfg.clip_min = "tf32", False, float(m["clip_min"])
        eng = PSTrainEngine(MLPSpec(in_dim=m["in_dim"], hidden=m["hidden"], classes=m["classes"], batch=m["batch"]), cfg, fabric)
        eng.names = [m["roles"][r] for r in ("hid_w", "hid_b", "sm_w", "sm_b")]
        _dbg("fused MLP engine built; preparing")
        eng.prepare()
        return eng
    eng = GenericPSEngine([(n, tuple(s)) for n, s, _ in spec["params"]], cfg, fabric,
                          shards=[sh for _, _, sh in spec["params"]])
    _dbg("engine built; preparing")
    eng.prepare()
    _dbg("engine ready")
    return eng


class _PsService:
    """ps-task side: owns the engine, binds graph variables to its buffers, runs the apply loop."""

    def __init__(self, server, spec: Dict[str, Any]):
        self.server, self.spec = server, spec
        self.engine = None
        self.ready = threading.Event()
        self.error: Optional[BaseException] = None
        self.applies = 0
        self._farewell = threading.Event()
        self._farewell_done = False
        self._stop = threading.Event()
        self._bound_names: List[str] = []
        self.thread = threading.Thread(target=self._run, name="dtf-fabric-ps", daemon=True)
        self.thread.start()

    def _run(self):
        try:
            srv = self.server
            eng = build_engine(srv.cluster, srv.job_name, srv.task_index, self.spec, srv.gpu_index)
            self.engine = eng
            rank = next(iter(eng.ranks))
            s = eng.ps_ranks.index(rank)
            rk = eng.ranks[rank]
            # graph variables on this task live in the engine's master buffer from now on
            def bind(gname, tensor):
                srv.store.bind(gname, tensor, initialized=False)        # (a value assigned earlier is carried over)
                self._bound_names.append(gname)
            if self.spec.get("mlp"):
                for role, gname in self.spec["mlp"]["roles"].items():
                    if eng.layout[role].shard == s:
                        bind(gname, eng.var_tensor(rank, role))
                if s == 0:
                    bind(self.spec["global_step"], eng.global_step_tensor(rank))
            else:
                for name in eng.names:
                    if eng.layout[name][0] == s:
                        bind(name, eng._view(rk.bufs["gmaster%d" % s], name))
                if s == 0:
                    bind(self.spec["global_step"], rk.bufs["gctl0"].tensor(torch.int64, eng.off["global_step"], 1).view(()))
            # a re-formed fabric (an earlier generation of this job was torn down on this task): the optimizer's own state
            # -- slots, Adam's beta powers -- continues where that generation stopped, like the variables do
            saved = srv.store.resources.pop("fabric_opt_state/" + self.spec.get("base_key", self.spec["key"]), None)
            if saved and hasattr(eng, "load_optimizer_state"):
                try:
                    _dbg("optimizer state carried over: %s" % (eng.load_optimizer_state(saved),))
                except Exception:      # noqa: BLE001 - the slots restart from zero, as before
                    import traceback
                    traceback.print_exc()
            self.ready.set()
            _dbg("ps service loop starts")
            per_round = 1 if eng.cfg.sync else eng.cfg.num_workers
            while not self._stop.is_set() and srv.is_running:
                for _ in range(per_round):
                    eng.ps_apply(rank, idle_ok=True)
                rk.stream.synchronize()
                self.applies += per_round
                if self._farewell.is_set() and not self._farewell_done and hasattr(eng, "release_all_tokens"):
                    eng.release_all_tokens(rank)
                    self._farewell_done = True
        except BaseException as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            self.error = e
            self.ready.set()

    def stop(self):
        self._stop.set()

    close = stop           # VariableStore.clear() closes its resources

    def farewell(self):
        self._farewell.set()

    def teardown(self) -> None:
        """The fabric re-forms (a task failed or restarted): stop the apply loop, move the graph variables out of the engine's
        HBM (they keep their values), release the buffers and the peer mappings.  Idempotent."""
        self._stop.set()
        if self.thread is not threading.current_thread():
            self.thread.join(15.0)
        eng, self.engine = self.engine, None
        for name in self._bound_names:
            self.server.store.unbind(name)
        self._bound_names = []
        if eng is not None and hasattr(eng, "optimizer_state"):
            try:                   # slots / beta powers live only in the engine: keep them for the next generation
                st = eng.optimizer_state()
                if st:
                    self.server.store.resources["fabric_opt_state/" + self.spec.get("base_key", self.spec["key"])] = st
            except Exception:      # noqa: BLE001 - a GPU in an error state: the next generation's slots start from zero
                import traceback
                traceback.print_exc()
        if eng is not None:
            try:
                eng.close()
            except Exception:      # noqa: BLE001
                import traceback
                traceback.print_exc()
        _dbg("ps service %s torn down" % self.spec["key"])


def ps_fabric_teardown(server, base_key: str, keep: Optional[str] = None) -> bool:
    """Tear down this task's services of the job ``base_key`` (all generations but ``keep``)."""
    for k, svc in list(server.store.resources.items()):
        if k.startswith("fabric_service/") and isinstance(svc, _PsService) and \
                svc.spec.get("base_key", svc.spec["key"]) == base_key and svc.spec["key"] != keep:
            svc.teardown()
            server.store.resources.pop(k, None)
    return True


def ps_fabric_setup(server, spec: Dict[str, Any]) -> bool:
    """Called through ``Server.rpc_fabric_setup`` on every ps task (idempotent per spec key).  A spec of a NEW generation of
    the same job first retires the older generations' services on this task."""
    ps_fabric_teardown(server, spec.get("base_key", spec["key"]), keep=spec["key"])
    svc = server.store.get_resource("fabric_service/" + spec["key"], lambda: _PsService(server, spec))
    if not svc.ready.wait(120.0):
        raise RuntimeError("ps task %s did not join the fabric within 120 s" % (server.task,))
    if svc.error is not None:
        raise RuntimeError("ps fabric setup failed: %r" % (svc.error,))
    return True


class FabricPSStrategy:
    def __init__(self, server):
        self.server = server
        self.cluster: ClusterSpec = server.cluster
        self.engine = None
        self.loss: Optional[Tensor] = None
        self._spec: Optional[Dict[str, Any]] = None
        self._primed = False
        self.mlp: Optional[Dict[str, Any]] = None         # the reference network recognised in the loss sub-graph (auto_fabric)
        self._last_loss: Optional[float] = None
        self._step_node_id: Optional[int] = None
        self._pinned: List[Any] = []
        self._spec_base: Optional[Dict[str, Any]] = None
        self._gen: Optional[int] = None          # fabric incarnation this worker is in
        self._min_gen = 0                         # ... and the least one it will join next (bumped by _abort)
        self._last_live = 0.0
        self._live_every = float(os.environ.get("DTF_FABRIC_LIVENESS_SECS", "2.0"))

    # -- graph construction ---------------------------------------------------------------------------------------
    def minimize(self, optimizer, loss, global_step: Variable, var_list: Optional[Sequence[Variable]] = None
                 ) -> Tuple[Tensor, Tensor]:
        if self.server.job_name != "worker":
            raise RuntimeError("FabricPSStrategy.minimize() is for worker tasks; ps tasks just server.join()")
        g = get_default_graph()
        vars_ = list(var_list) if var_list is not None else trainable_variables()
        params = []
        for v in vars_:
            spec = DeviceSpec.from_string(v.device)
            if spec.job != "ps":
                raise ValueError("variable %s is not placed on a ps task (device %r): build the model under "
                                 "replica_device_setter" % (v.var_name, v.device))
            params.append((v.var_name, [int(d) for d in v.shape], int(spec.task or 0)))
        fs = dict(optimizer.fused_spec())
        import hashlib
        import json
        # the key must be identical in every worker process (it names the fabric buffers and the ps-side service)
        key = hashlib.md5(json.dumps([params, sorted((k, str(v)) for k, v in fs.items())]).encode()).hexdigest()[:12]
        # the job's identity; the spec of one fabric INCARNATION (generation, batch-specialised engine) is derived from it in
        # _ensure_engine
        self._spec_base = {"key": "f" + key, "params": params, "optimizer": fs, "global_step": global_step.var_name}
        self._spec = dict(self._spec_base)
        loss_t = convert_to_tensor(loss)
        from .auto_fabric import match_reference_mlp
        m = match_reference_mlp(loss_t, vars_)
        if m is not None and m["hidden"] <= 128 and m["classes"] <= 16 and len({sh for _, _, sh in params}) <= 4:
            self.mlp = m          # the batch size is only known at the first run: the engine spec is completed there
        order = needed_nodes([loss_t], set())
        placeholders = [n for n in order if n.op_type == "Placeholder"]
        with _device.device(None), _device.device(loss_t.device or None):
            step = g.create_node("FabricTrainStep", placeholders,
                                 {"strategy": self, "loss": loss_t, "var_nodes": [v._node for v in vars_],
                                  "order": order,
                                  # plan-time rewrites of the worker sub-graph (framework/fusion.py): only the loss leaves it
                                  "fusions": _fusion.plan_fusions(order, {loss_t.id}, {v._node.id for v in vars_},
                                                                  {n.id: 0 for n in order}, set())},
                                 "fabric_train_step", loss_t.dtype, ())
            self.loss = _ops.identity(step, name="fabric_loss")
        self._step_node_id = step.id
        # the user's own loss tensor, fetched next to the train op (or alone, for validation), is answered by the engine
        overrides = g.__dict__.setdefault("_fetch_overrides", {})
        overrides[loss_t.id] = self._loss_fetch
        self._loss_t = loss_t
        return step, self.loss

    def _loss_fetch(self, session, feeds: Dict[int, Any], fetch_ids) -> Any:
        """Session fetch override of the loss tensor: the value the fused step of THIS run computed, or -- fetched without
        the train op (validation, reference distributed_mnist.py:160-165) -- the engine's forward-only kernel."""
        if self._step_node_id in fetch_ids and self._last_loss is not None:
            return torch.tensor(self._last_loss, dtype=torch.float32)
        if self.mlp is not None and self.engine is not None and self.mlp["x"].id in feeds and self.mlp["y_"].id in feeds:
            ev = self.engine.evaluate(feeds[self.mlp["x"].id].float(), feeds[self.mlp["y_"].id].float())
            return torch.tensor(ev["loss"], dtype=torch.float32)
        return NotImplemented

    # -- runtime ---------------------------------------------------------------------------------------------------
    def _ps_call(self, t: int, method: str, *args):
        """Control-plane call to ps task ``t`` (in-process server: direct call)."""
        from .server import local_server_for
        srv = local_server_for(self.cluster.task_address("ps", t))
        if srv is not None:
            return getattr(srv, "rpc_" + method)(*args)
        return self.server.peer("ps", t).call(method, *args)

    def _ensure_engine(self, batch: Optional[int] = None) -> None:
        if self.engine is not None:
            return
        from ..framework import errors
        base = self._spec_base["key"]
        rank, _ = fabric_rank_of(self.cluster, self.server.job_name, self.server.task_index)
        try:
            gen = int(self._ps_call(0, "fabric_generation", base, self._min_gen, rank))
        except (OSError, EOFError, ConnectionError) as e:
            raise errors.UnavailableError("ps task 0 is unreachable (%s: %s)" % (type(e).__name__, e))
        spec = dict(self._spec_base)
        spec["base_key"], spec["generation"] = base, gen
        spec["key"] = base + ("g%d" % gen if gen else "")
        if self.mlp is not None:
            if batch is None or batch > 128:
                self.mlp = None       # the fused step handles <= 128 rows per worker step: generic engine instead
            else:
                m = self.mlp
                spec["mlp"] = {"roles": {r: m[r].var_name for r in ("hid_w", "hid_b", "sm_w", "sm_b")},
                               "in_dim": m["in_dim"], "hidden": m["hidden"], "classes": m["classes"],
                               "batch": int(batch), "clip_min": m["clip_min"]}
                spec["key"] += "m%d" % int(batch)
        self._spec, self._gen = spec, gen
        self._farewell_sent = False
        # (1) every ps task joins the fabric (control-plane RPC; blocks until its buffers are exported)
        threads, errs = [], []
        for t in range(self.cluster.num_tasks("ps")):
            def call(t=t):
                try:
                    self._ps_call(t, "fabric_setup", spec)
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)
            th = threading.Thread(target=call, daemon=True)
            th.start()
            threads.append(th)
        # (2) this worker joins too (the handle exchange needs all ranks)
        try:
            self.engine = build_engine(self.cluster, self.server.job_name, self.server.task_index, spec, self.server.gpu_index)
            for th in threads:
                th.join(180.0)
            if errs:
                raise errs[0]
            if self.mlp is not None:
                me = next(iter(self.engine.ranks))
                gs = self.engine.adopt_global_step(me)      # restored checkpoint / re-formed fabric: tokens continue from gs
                _dbg("fused MLP worker adopted global_step %d" % gs)
            self._primed = False
        except BaseException as e:  # noqa: BLE001 - a rendezvous that never completed, a ps that went away mid-setup, ...
            if isinstance(e, (KeyboardInterrupt, SystemExit)):
                raise
            self._abort("building generation %d of the fabric failed: %s: %s" % (gen, type(e).__name__, e))
        self._last_live = time.time()
        if not getattr(self, "_atexit", False):
            import atexit
            atexit.register(self._report)
            self._atexit = True

    # -- failure handling (reference example_between_graph.py:99: MonitoredTrainingSession recovers from AbortedError) -----
    def _abort(self, why: str) -> None:
        """A peer of this fabric incarnation is gone (a device-side wait timed out, the rendezvous failed, ps task 0 reports
        a newer generation): leave it -- release this worker's engine, ask the reachable ps tasks to retire their services
        (the graph variables keep their values) -- and raise the error ``MonitoredTrainingSession`` recovers from:
        ``UnavailableError`` when a ps task does not answer on the control plane, ``AbortedError`` otherwise.  The next
        ``train_step`` forms generation + 1 with whoever is alive or restarted by then."""
        from ..framework import errors
        eng, self.engine = self.engine, None
        self._min_gen = max(self._min_gen, (self._gen if self._gen is not None else -1) + 1)
        self._primed = False
        if eng is not None:
            try:
                eng.close()
            except Exception:      # noqa: BLE001
                pass
        down = []
        for t in range(self.cluster.num_tasks("ps")):
            try:
                self._ps_call(t, "fabric_teardown", self._spec_base["key"])
            except Exception:      # noqa: BLE001
                down.append(t)
        self.aborts = getattr(self, "aborts", 0) + 1
        msg = "fabric generation %s of job %s aborted on worker %d: %s" % (self._gen, self._spec_base["key"],
                                                                           self.server.task_index, why)
        print("dtf.fabric: " + msg, flush=True)
        if down:
            raise errors.UnavailableError(msg + " (ps task(s) %s unreachable)" % down)
        raise errors.AbortedError(msg)

    def _after_step(self, t0: float) -> None:
        """Failure detection without a per-step cost: a step that took long had a device-side wait run into its timeout (a
        peer stopped pushing / applying) -> read the error words; every ``DTF_FABRIC_LIVENESS_SECS`` ask ps task 0 whether
        the fabric has moved on to a newer generation (a restarted task re-joined)."""
        now = time.time()
        eng = self.engine
        timeout_s = eng.cfg.timeout_ns / 1e9
        if now - t0 > min(1.0, timeout_s / 4):
            try:
                eng.check_errors()
            except RuntimeError as e:
                if not getattr(self, "_farewell_sent", False):
                    self._abort(str(e))
        if now - self._last_live > self._live_every:
            self._last_live = now
            try:
                cur = int(self._ps_call(0, "fabric_current_generation", self._spec_base["key"]))
            except Exception as e:      # noqa: BLE001
                self._abort("ps task 0 does not answer (%s)" % type(e).__name__)
            if cur > self._gen:
                self._abort("the job moved on to fabric generation %d (a task restarted)" % cur)

    def farewell(self) -> None:
        """This replica's training loop is over (sync mode): ask every ps shard to release the device-side token waits of
        the replicas that are still inside a step (see ``Server.rpc_fabric_farewell``).  Idempotent, best effort."""
        if self.engine is None or getattr(self, "_farewell_sent", False) or not self._spec["optimizer"].get("sync"):
            return
        self._farewell_sent = True
        for t in range(self.cluster.num_tasks("ps")):
            try:
                from .server import local_server_for
                srv = local_server_for(self.cluster.task_address("ps", t))
                if srv is not None:
                    srv.rpc_fabric_farewell(self._spec["key"])
                else:
                    self.server.peer("ps", t).call("fabric_farewell", self._spec["key"])
            except Exception:      # noqa: BLE001 - the ps may already be gone
                pass

    def _trace_into(self, tracer) -> None:
        """FULL_TRACE run on the fused path: the device-side rings (this worker's step-kernel phases, every ps shard's
        ``ps_apply`` launches) become timeline events of the run, one process row per task -- the whole step per
        ``/job:worker/task:i`` and ``/job:ps/task:k`` (reference example_in_graph.py:65-68)."""
        events = list(self.engine.step_stats()) if hasattr(self.engine, "step_stats") else []
        for t in range(self.cluster.num_tasks("ps")):
            try:
                from .server import local_server_for
                srv = local_server_for(self.cluster.task_address("ps", t))
                ev = srv.rpc_fabric_step_stats(self._spec["key"]) if srv is not None else \
                    self.server.peer("ps", t).call("fabric_step_stats", self._spec["key"])
                events += ev
            except Exception:      # noqa: BLE001 - tracing is best effort
                pass
        seen = self.__dict__.setdefault("_traced", set())
        for e in events:
            k = (e["task"], e["name"], e["start_us"])
            if k not in seen:
                seen.add(k)
                tracer.add_event(e)

    def _report(self) -> None:
        """One line at process exit: which engine ran this worker's steps and how many kernels of ours it launched."""
        self.farewell()
        try:
            from ..ops import cuda_lib
            kind = "fused MLP step (mlp_step_kernel + ps_apply_kernel)" if self.mlp is not None else "generic fabric engine"
            print("dtf.fabric: worker %d ran %d steps on the %s; %d kernel launches of ours in this process" % (
                self.server.task_index, getattr(self, "_steps", 0), kind, cuda_lib.launch_count()), flush=True)
        except Exception:      # noqa: BLE001
            pass

    def _prime(self) -> None:
        """First step after (re)initialisation: adopt the ps's global_step as this worker's token base."""
        eng = self.engine
        rank = next(iter(eng.ranks))
        gs = eng.adopt_global_step(rank)
        _dbg("worker primed at global_step %d" % gs)
        self._primed = True

    def train_step(self, ctx, node, feeds: Sequence[torch.Tensor]) -> torch.Tensor:
        _dbg("train_step enter") if self.engine is None else None
        if self.mlp is not None:
            ph_ids = [p.id for p in node.inputs]
            x = feeds[ph_ids.index(self.mlp["x"].id)]
            y = feeds[ph_ids.index(self.mlp["y_"].id)]
            self._ensure_engine(int(x.shape[0]))
            if self.mlp is not None:
                t0 = time.time()
                out = self._mlp_step(x, y)
                self._after_step(t0)
                if ctx.tracer is not None:
                    self._trace_into(ctx.tracer)
                return out
        self._ensure_engine()
        if not self._primed:
            self._prime()
        eng = self.engine
        rank = next(iter(eng.ranks))
        rk = eng.ranks[rank]
        a = node.attrs
        var_nodes: List[Tensor] = a["var_nodes"]
        ph = list(node.inputs)

        def loss_fn(leaves, *unused):
            sub = ExecContext(ctx.store, ctx.task, eng.ranks[rank].device.index, None, None)
            sub.force_device = rk.device
            for pnode, val in zip(ph, feeds):
                sub.values[pnode.id] = val.to(rk.device, non_blocking=True)
            for vn in var_nodes:
                sub.values[vn.id] = leaves[vn.attrs["var_name"]]
            if a.get("fusions"):
                sub.fusions = _fusion.FusionState(a["fusions"])
            execute([n for n in a["order"] if n.id not in sub.values], sub, True)
            return sub.values[a["loss"].id]
        t0 = time.time()
        loss = eng.worker_step(rank, loss_fn)
        if time.time() - t0 > 0.5 or time.time() - self._last_live > self._live_every:
            float(loss)                      # the step's kernels have run (a timed-out wait has set its error word by now)
            self._after_step(t0)
        return loss


def _mlp_step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """One fused worker step (mlp_step_kernel): the fed batch goes through pinned staging (H2D on the copy stream), the
    loss comes back with the step; push / aggregate / apply / token are the engine's device-side protocol."""
    eng = self.engine
    if int(x.shape[0]) != eng.spec.batch:
        raise ValueError("the fused MLP step was built for batches of %d rows, got %d (feed a constant batch size, or set "
                         "DTF_FABRIC=0)" % (eng.spec.batch, int(x.shape[0])))
    if not self._pinned:
        for _ in range(2):
            self._pinned.append((torch.empty((eng.spec.batch, eng.spec.in_dim), dtype=torch.float32).pin_memory(),
                                 torch.empty((eng.spec.batch, eng.spec.classes), dtype=torch.float32).pin_memory()))
        self._pin_i = 0
    px, py = self._pinned[self._pin_i]
    self._pin_i ^= 1
    px.copy_(x)
    py.copy_(y)
    loss = eng.step(px, py, sync_loss=True)
    self._steps = getattr(self, "_steps", 0) + 1
    self._last_loss = float(loss)
    return torch.tensor(self._last_loss, dtype=torch.float32)


FabricPSStrategy._mlp_step = _mlp_step


@register_kernel("FabricTrainStep", stateful=True)
def _k_fabric_train_step(ctx, node, *feeds):
    with torch.enable_grad():
        return node.attrs["strategy"].train_step(ctx, node, feeds)
