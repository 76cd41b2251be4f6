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


def init_fabric_process_group(cluster: ClusterSpec, job: str, task: int) -> Tuple[int, int]:
    """Join the fabric's torch.distributed group (idempotent).  gloo: only a store + barriers are needed."""
    import torch.distributed as dist
    rank, world = fabric_rank_of(cluster, job, task)
    with _PG_LOCK:
        if not dist.is_initialized():
            host, _, port = cluster.all_tasks()[0][2].rpartition(":")
            if host in ("localhost", "", "0.0.0.0"):
                host = "127.0.0.1"
            port = int(port) + int(os.environ.get("DTF_FABRIC_PORT_OFFSET", "1000"))
            import datetime
            _dbg("joining process group rank %d/%d at %s:%d" % (rank, world, host, port))
            dist.init_process_group("gloo", init_method="tcp://%s:%d" % (host, port), rank=rank, world_size=world,
                                    timeout=datetime.timedelta(seconds=float(os.environ.get("DTF_FABRIC_TIMEOUT", "300"))))
            _dbg("process group up")
    return rank, world


def _engine_cfg(spec: Dict[str, Any], num_ps: int, num_workers: int):
    from .ps_engine import EngineConfig
    opt = dict(spec["optimizer"])
    return EngineConfig(num_ps=num_ps, num_workers=num_workers, sync=bool(opt.get("sync", False)),
                        replicas_to_aggregate=opt.get("replicas_to_aggregate"), optimizer=opt,
                        timeout_ns=int(spec.get("timeout_ns", 20_000_000_000)))


def build_engine(cluster: ClusterSpec, job: str, task: int, spec: Dict[str, Any], gpu_index: Optional[int]):
    from .fabric import Fabric
    from .generic_engine import GenericPSEngine
    rank, world = init_fabric_process_group(cluster, job, task)
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
            if self.spec.get("mlp"):
                for role, gname in self.spec["mlp"]["roles"].items():
                    if eng.layout[role].shard == s:
                        srv.store.bind(gname, eng.var_tensor(rank, role), initialized=False)
                if s == 0:
                    srv.store.bind(self.spec["global_step"], eng.global_step_tensor(rank), initialized=False)
            else:
                for name in eng.names:
                    if eng.layout[name][0] == s:
                        srv.store.bind(name, eng._view(rk.bufs["gmaster%d" % s], name), initialized=False)
                if s == 0:
                    gs = rk.bufs["gctl0"].tensor(torch.int64, eng.off["global_step"], 1).view(())
                    srv.store.bind(self.spec["global_step"], gs, initialized=False)
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

    def farewell(self):
        self._farewell.set()


def ps_fabric_setup(server, spec: Dict[str, Any]) -> bool:
    """Called through ``Server.rpc_fabric_setup`` on every ps task (idempotent per spec key)."""
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
        self._spec = {"key": "f" + key, "params": params, "optimizer": fs, "global_step": global_step.var_name}
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
                                  "order": order}, "fabric_train_step", loss_t.dtype, ())
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
    def _ensure_engine(self, batch: Optional[int] = None) -> None:
        if self.engine is not None:
            return
        if self.mlp is not None:
            if batch is None or batch > 128:
                self.mlp = None       # the fused step handles <= 128 rows per worker step: generic engine instead
            else:
                m = self.mlp
                self._spec["mlp"] = {"roles": {r: m[r].var_name for r in ("hid_w", "hid_b", "sm_w", "sm_b")},
                                     "in_dim": m["in_dim"], "hidden": m["hidden"], "classes": m["classes"],
                                     "batch": int(batch), "clip_min": m["clip_min"]}
                self._spec["key"] += "m%d" % int(batch)
        # (1) every ps task joins the fabric (control-plane RPC; blocks until its buffers are exported)
        threads, errs = [], []
        for t in range(self.cluster.num_tasks("ps")):
            def call(t=t):
                try:
                    from .server import local_server_for
                    addr = self.cluster.task_address("ps", t)
                    srv = local_server_for(addr)
                    if srv is not None:
                        srv.rpc_fabric_setup(self._spec)
                    else:
                        self.server.peer("ps", t).call("fabric_setup", self._spec)
                except BaseException as e:  # noqa: BLE001
                    errs.append(e)
            th = threading.Thread(target=call, daemon=True)
            th.start()
            threads.append(th)
        # (2) this worker joins too (the handle exchange needs all ranks)
        self.engine = build_engine(self.cluster, self.server.job_name, self.server.task_index, self._spec,
                                   self.server.gpu_index)
        for th in threads:
            th.join(180.0)
        if errs:
            raise errs[0]
        import atexit
        atexit.register(self._report)

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
                out = self._mlp_step(x, y)
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
            execute([n for n in a["order"] if n.id not in sub.values], sub, True)
            return sub.values[a["loss"].id]
        loss = eng.worker_step(rank, loss_fn)
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
