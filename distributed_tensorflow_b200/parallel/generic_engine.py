"""Model-agnostic fabric parameter server: any set of named parameters, any loss.

Same device-resident protocol as :mod:`parallel.ps_engine` (gradient slots + stamps + arrival counters in ps
HBM, one fused ``ps_apply`` kernel per aggregate, tokens in worker mailboxes, sync or async) but the worker's
compute is arbitrary: the caller supplies ``loss_fn(params, *batch) -> loss`` written with the framework's ops
(``ops/native.py``: tcgen05 GEMM / conv-as-GEMM / fused xent, autograd-wrapped).  Per step a worker runs

    wait_token (+ pull_shadow: peer loads of the ps's fp32 parameters into the local replica)
    -> forward/backward (our kernels + autograd glue) -> gradients packed into one flat buffer
    -> push_grad (vectorised NVLink stores into its slot on the ps + stamp + arrivals)

and the ps runs ``ps_apply`` (N-way reduce + mean + SGD/Momentum/Adam + publish + tokens).  This is what runs
ResNet-18 under the ps API (BASELINE.json config 5): see ``models/resnet.py``.

Variables are sharded across ps ranks round-robin in creation order, exactly like ``replica_device_setter``.
"""
from __future__ import annotations

import ctypes
import ctypes as _ct
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import torch

from ..ops import cuda_lib
from ..ops.cuda_lib import MAX_WORKERS, PsApplyArgs, round_up
from .fabric import Fabric, FabricBuffer
from .ps_engine import EngineConfig, _KIND, _Rank

__all__ = ["GenericPSEngine"]

_PUSH_CTAS = 64


class GenericPSEngine:
    def __init__(self, param_shapes: Sequence[Tuple[str, Tuple[int, ...]]], cfg: EngineConfig, fabric: Fabric,
                 shards: Optional[Sequence[int]] = None):
        self.cfg, self.fabric = cfg, fabric
        self.lib = cuda_lib.load()
        self.world = fabric.world_size
        if cfg.colocated:
            assert self.world == 1 and cfg.num_ps == 1 and cfg.num_workers == 1
            self.ps_ranks, self.worker_ranks = [0], [0]
        elif getattr(cfg, "ps_on_workers", False):
            # every rank a worker, ps shard s on worker s's GPU and stream (see PSTrainEngine): the caller runs
            # worker_step(r) then ps_apply(r) on those ranks
            assert self.world == cfg.num_workers and 1 <= cfg.num_ps <= self.world
            self.ps_ranks = list(range(cfg.num_ps))
            self.worker_ranks = list(range(self.world))
        else:
            assert self.world == cfg.num_ps + cfg.num_workers
            self.ps_ranks = list(range(cfg.num_ps))
            self.worker_ranks = list(range(cfg.num_ps, self.world))
        if cfg.num_workers > MAX_WORKERS:
            raise ValueError("at most %d workers" % MAX_WORKERS)
        # layout: (shard, offset, numel) per variable, round-robin in creation order, 64-element aligned
        self.layout: Dict[str, Tuple[int, int, Tuple[int, ...]]] = {}
        sizes = [0] * cfg.num_ps
        for i, (name, shape) in enumerate(param_shapes):
            shard = (i % cfg.num_ps) if shards is None else int(shards[i])      # explicit placement (device strings)
            off = round_up(sizes[shard], 64)
            n = 1
            for d in shape:
                n *= int(d)
            self.layout[name] = (shard, off, tuple(int(d) for d in shape))
            sizes[shard] = off + n
        self.shard_elems = [round_up(max(s, 64), 64) for s in sizes]
        self.names = [n for n, _ in param_shapes]
        self.R = cfg.replicas_to_aggregate or cfg.num_workers
        self.opt = dict(cfg.optimizer)
        self.kind = _KIND[self.opt["kind"]]
        self.ctl_bytes = self.lib.dtf_sizeof_ps_control()
        self.mb_bytes = self.lib.dtf_sizeof_mailbox()
        self.off = {k: self.lib.dtf_offsetof_ctl(i) for i, k in enumerate(
            ["global_step", "param_version", "beta1_power", "beta2_power", "dropped_stale", "applied_total",
             "staleness_hist", "staleness_sum", "err", "w", "w_stride", "consumed"])}
        self.ranks: Dict[int, _Rank] = {r: _Rank(r, d) for r, d in fabric.local_ranks.items()}
        for rk in self.ranks.values():
            with torch.cuda.device(rk.device):
                rk.stream = torch.cuda.Stream(rk.device)
        W = cfg.num_workers
        # ---- NVLS mode (cfg.nvls): symmetric buffers.  Gradients stay in the WORKERS' HBM and the ps sums them in the
        # switch (multimem.ld_reduce): it ingests n floats per step instead of W*n.  The fp32 parameters are published
        # with ONE multimem.st stream into every GPU's replica: the ps emits n floats instead of W*n. ----
        want = cfg.nvls
        self.nvls, self.nvls_multicast = False, False
        self.sym_grads: List[Any] = []
        self.sym_repl: List[Any] = []
        if want and self.world > 1:
            level = fabric.nvls_level()
            if want == "auto" and level < 2:
                want = False
            elif level == 0:
                raise RuntimeError("nvls=True needs CUDA VMM (POSIX fd export); not available on this machine")
        if want and self.world > 1:
            self.nvls = True
            for s in range(cfg.num_ps):
                self.sym_grads.append(fabric.alloc_symmetric("gsgrads%d" % s, self.shard_elems[s] * 4))
                self.sym_repl.append(fabric.alloc_symmetric("gsrepl%d" % s, self.shard_elems[s] * 4))
            self.nvls_multicast = self.sym_grads[0].multicast
        self.push_ctas = 1 if self.nvls else _PUSH_CTAS
        for r, rk in self.ranks.items():
            if r in self.ps_ranks:
                s = self.ps_ranks.index(r)
                n = self.shard_elems[s]
                for name, nbytes in (("gctl%d" % s, self.ctl_bytes), ("gmaster%d" % s, n * 4), ("ggrads%d" % s, n * 4 * W),
                                     ("gslot_m%d" % s, n * 4), ("gslot_v%d" % s, n * 4), ("gshadow%d" % s, n * 2)):
                    if self.nvls and name.startswith("ggrads"):
                        rk.bufs[name] = self.sym_grads[s].local(r)       # stays zero: the ps contributes nothing
                    else:
                        rk.bufs[name] = fabric.alloc(r, name, nbytes)
                for name in ("gctl%d" % s, "gmaster%d" % s) + (() if self.nvls else ("ggrads%d" % s,)):
                    fabric.publish(r, name)
            if r in self.worker_ranks:
                w = self.worker_ranks.index(r)
                rk.bufs["gmailbox_w%d" % w] = fabric.alloc(r, "gmailbox_w%d" % w, self.mb_bytes * cfg.num_ps)
                rk.bufs["gmisc_w%d" % w] = fabric.alloc(r, "gmisc_w%d" % w, 256)
                fabric.publish(r, "gmailbox_w%d" % w)
                for s in range(cfg.num_ps):
                    if self.nvls:
                        rk.bufs["greplica%d_w%d" % (s, w)] = self.sym_repl[s].local(r)
                        rk.bufs["glocalgrad%d_w%d" % (s, w)] = self.sym_grads[s].local(r)
                    else:
                        rk.bufs["greplica%d_w%d" % (s, w)] = fabric.alloc(r, "greplica%d_w%d" % (s, w), self.shard_elems[s] * 4)
                        rk.bufs["glocalgrad%d_w%d" % (s, w)] = fabric.alloc(r, "glocalgrad%d_w%d" % (s, w), self.shard_elems[s] * 4)
        self.peer: Dict[Tuple[int, str], FabricBuffer] = {}
        for r in self.ranks:
            if r in self.worker_ranks:
                for s, pr in enumerate(self.ps_ranks):
                    for base in ("gctl", "gmaster") + (() if self.nvls else ("ggrads",)):
                        self.peer[(r, "%s%d" % (base, s))] = fabric.peer(r, pr, "%s%d" % (base, s))
            if r in self.ps_ranks:
                s = self.ps_ranks.index(r)
                for w, wr in enumerate(self.worker_ranks):
                    self.peer[(r, "gmailbox_w%d" % w)] = fabric.peer(r, wr, "gmailbox_w%d" % w)
                    if self.nvls:
                        self.peer[(r, "gsgrads%d_w%d" % (s, w))] = self.sym_grads[s].peer(r, wr)
                        self.peer[(r, "gsrepl%d_w%d" % (s, w))] = self.sym_repl[s].peer(r, wr)
        self._p: Dict[int, PsApplyArgs] = {}
        for r, rk in self.ranks.items():
            if r not in self.ps_ranks:
                continue
            s = self.ps_ranks.index(r)
            n = self.shard_elems[s]
            a = PsApplyArgs()
            a.ctl, a.master = rk.bufs["gctl%d" % s].ptr, rk.bufs["gmaster%d" % s].ptr
            a.slot_m, a.slot_v, a.shadow = rk.bufs["gslot_m%d" % s].ptr, rk.bufs["gslot_v%d" % s].ptr, rk.bufs["gshadow%d" % s].ptr
            for w in range(W):
                a.grad[w] = self.peer[(r, "gsgrads%d_w%d" % (s, w))].ptr if self.nvls else rk.bufs["ggrads%d" % s].ptr + w * n * 4
                a.mailbox[w] = self.peer[(r, "gmailbox_w%d" % w)].ptr + s * self.mb_bytes
            if self.nvls and self.nvls_multicast:
                a.grad_mc = self.sym_grads[s].mc(r)
                a.master_mc = self.sym_repl[s].mc(r)
            a.n, a.num_workers, a.replicas_to_aggregate = n, W, self.R
            a.ctas_per_push = self.push_ctas
            a.mode, a.kind = (0 if cfg.sync else 1), self.kind
            a.lr, a.momentum = float(self.opt["lr"]), float(self.opt.get("momentum", 0.0))
            a.beta1, a.beta2 = float(self.opt.get("beta1", 0.9)), float(self.opt.get("beta2", 0.999))
            a.eps = float(self.opt.get("eps", self.opt.get("epsilon", 1e-8)))
            a.nesterov, a.publish_replicas, a.num_zero = int(bool(self.opt.get("nesterov", False))), 0, 0
            a.timeout_ns, a.grid = cfg.timeout_ns, 0
            a.system_scope = 0 if cfg.colocated else 1
            self._p[r] = a

    # -- parameters ------------------------------------------------------------------------------------------
    def _view(self, buf: FabricBuffer, name: str) -> torch.Tensor:
        shard, off, shape = self.layout[name]
        n = 1
        for d in shape:
            n *= d
        return buf.tensor(torch.float32, off * 4, n).view(shape)

    def prepare(self) -> None:
        """Zero the buffers and publish beta powers WITHOUT parameter values (they arrive through bound variables)."""
        self.init_params({})

    def adopt_global_step(self, rank: int) -> int:
        """Worker: take the ps's current global_step as the base of its token / stamp sequence (fresh start or
        restore from a checkpoint) and reset its local step count."""
        rk = self.ranks[rank]
        w = self.worker_ranks.index(rank)
        rk.stream.synchronize()
        gs = int(self.peer[(rank, "gctl0")].tensor(torch.int64, self.off["global_step"], 1).cpu()[0])
        mb = rk.bufs["gmailbox_w%d" % w].tensor(torch.int64)
        per = self.mb_bytes // 8
        for s in range(self.cfg.num_ps):
            mb[s * per] = gs
            mb[s * per + 1] = gs
        torch.cuda.synchronize(rk.device)
        rk.base, rk.step = gs, 0
        return gs

    def init_params(self, values: Dict[str, torch.Tensor]) -> None:
        for r, rk in self.ranks.items():
            if r in self.ps_ranks:
                s = self.ps_ranks.index(r)
                with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
                    for base in ("gctl", "gmaster", "ggrads", "gslot_m", "gslot_v"):
                        rk.bufs["%s%d" % (base, s)].tensor(torch.uint8).zero_()
                    for name in self.names:
                        if self.layout[name][0] == s and name in values:
                            self._view(rk.bufs["gmaster%d" % s], name).copy_(values[name].to(rk.device).float())
                    b = rk.bufs["gctl%d" % s].tensor(torch.float32, self.off["beta1_power"], 2)
                    b[0] = float(self.opt.get("beta1", 0.9))
                    b[1] = float(self.opt.get("beta2", 0.999))
                    if self.nvls:
                        self.publish_params(r)
                rk.stream.synchronize()
            if r in self.worker_ranks:
                w = self.worker_ranks.index(r)
                with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
                    rk.bufs["gmailbox_w%d" % w].tensor(torch.uint8).zero_()
                    rk.bufs["gmisc_w%d" % w].tensor(torch.uint8).zero_()
                rk.stream.synchronize()
            rk.step = 0
            rk.base = 0
        self.fabric.barrier()

    def publish_params(self, rank: int) -> None:
        """NVLS mode, ps side: copy the fp32 master into every worker's replica -- one multimem.st stream when the box
        has NVLS (the switch fans out), otherwise one peer store per worker (``fabric_bcast`` kernel)."""
        rk = self.ranks[rank]
        s = self.ps_ranks.index(rank)
        n = self.shard_elems[s]
        W = self.cfg.num_workers
        peers = (_ct.c_void_p * 16)(*[self.peer[(rank, "gsrepl%d_w%d" % (s, w))].ptr for w in range(W)])
        mc = self.sym_repl[s].mc(rank) if self.nvls_multicast else None
        with torch.cuda.device(rk.device):
            rc = self.lib.dtf_fabric_bcast(rk.bufs["gmaster%d" % s].ptr, mc, peers, W, n * 4, 0, rk.stream.cuda_stream)
        assert rc == 0, rc
        cuda_lib._bump()

    def state_dict(self) -> Dict[str, torch.Tensor]:
        out = {}
        for r, rk in self.ranks.items():
            if r in self.ps_ranks:
                s = self.ps_ranks.index(r)
                rk.stream.synchronize()
                for name in self.names:
                    if self.layout[name][0] == s:
                        out[name] = self._view(rk.bufs["gmaster%d" % s], name).detach().cpu().clone()
                if s == 0:
                    out["global_step"] = torch.tensor(self.read_ctl(0, "global_step"), dtype=torch.int64)
        return out

    def optimizer_state(self) -> Dict[str, torch.Tensor]:
        """Optimizer slots of the LOCAL ps shards under their TF names + the beta powers (shard 0): what a fabric
        re-formation carries over next to the graph variables."""
        out: Dict[str, torch.Tensor] = {}
        kind = self.opt.get("kind", "sgd")
        if kind == "sgd":
            return out
        for r, rk in self.ranks.items():
            if r not in self.ps_ranks:
                continue
            s = self.ps_ranks.index(r)
            rk.stream.synchronize()
            for name in self.names:
                if self.layout[name][0] != s:
                    continue
                out[name + "/" + ("Adam" if kind == "adam" else "Momentum")] = \
                    self._view(rk.bufs["gslot_m%d" % s], name).detach().cpu().clone()
                if kind == "adam":
                    out[name + "/Adam_1"] = self._view(rk.bufs["gslot_v%d" % s], name).detach().cpu().clone()
            if kind == "adam":
                b = rk.bufs["gctl%d" % s].tensor(torch.float32, self.off["beta1_power"], 2).cpu()
                out["beta_powers/%d" % s] = b.clone()
        return out

    def load_optimizer_state(self, state: Dict[str, torch.Tensor]) -> List[str]:
        done: List[str] = []
        suffix = {"Momentum": "gslot_m", "Adam": "gslot_m", "Adam_1": "gslot_v"}
        for r, rk in self.ranks.items():
            if r not in self.ps_ranks:
                continue
            s = self.ps_ranks.index(r)
            with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
                for key, val in state.items():
                    if key == "beta_powers/%d" % s:
                        rk.bufs["gctl%d" % s].tensor(torch.float32, self.off["beta1_power"], 2).copy_(val.to(rk.device).float())
                        done.append(key)
                        continue
                    if "/" not in key:
                        continue
                    name, slot = key.rsplit("/", 1)
                    if name not in self.layout or self.layout[name][0] != s or slot not in suffix:
                        continue
                    view = self._view(rk.bufs["%s%d" % (suffix[slot], s)], name)
                    if tuple(view.shape) != tuple(val.shape):
                        continue
                    view.copy_(val.to(rk.device).float())
                    done.append(key)
            rk.stream.synchronize()
        return done

    def read_ctl(self, shard: int, fld: str, count: int = 1):
        rk = self.ranks[self.ps_ranks[shard]]
        t = rk.bufs["gctl%d" % shard].tensor(torch.int64, self.off[fld], count).cpu()
        return int(t[0]) if count == 1 else t.tolist()

    # -- worker side -------------------------------------------------------------------------------------------
    def pull(self, rank: int) -> Dict[str, torch.Tensor]:
        """Wait for this step's token, then copy the ps's fp32 parameters into the local replica (peer loads)."""
        rk = self.ranks[rank]
        w = self.worker_ranks.index(rank)
        st = rk.stream.cuda_stream
        params = {}
        with torch.cuda.device(rk.device):
            for s in range(self.cfg.num_ps):
                rc = self.lib.dtf_wait_token(rk.bufs["gmailbox_w%d" % w].ptr + s * self.mb_bytes, getattr(rk, "base", 0) + rk.step, None,
                                             self.cfg.timeout_ns, rk.bufs["gmisc_w%d" % w].ptr + 16, st)
                assert rc == 0, rc
                if self.cfg.sync and self.R < self.cfg.num_workers:
                    # backup workers: this worker's previous push may still be waiting in its slot (a straggler holds the
                    # next token already); the slot / local gradient buffer is rewritten only once the ps folded it in or
                    # dropped it as stale (TF's ConditionalAccumulator copies under a lock instead)
                    per_push = 1 if self.nvls else _PUSH_CTAS
                    rc = self.lib.dtf_wait_token(self.peer[(rank, "gctl%d" % s)].ptr + self.off["consumed"] + w * 8,
                                                 rk.step * per_push, None, self.cfg.timeout_ns,
                                                 rk.bufs["gmisc_w%d" % w].ptr + 16, st)
                    assert rc == 0, rc
                if not self.nvls:        # NVLS mode: the ps already stored the new parameters into this replica
                    rep = rk.bufs["greplica%d_w%d" % (s, w)]
                    rc = self.lib.dtf_pull_shadow(self.peer[(rank, "gmaster%d" % s)].ptr, rep.ptr, self.shard_elems[s] * 4, 148, st)
                    assert rc == 0, rc
            cuda_lib._bump((1 if self.nvls else 2) * self.cfg.num_ps)
            for name in self.names:
                s = self.layout[name][0]
                params[name] = self._view(rk.bufs["greplica%d_w%d" % (s, w)], name)
        return params

    def push(self, rank: int, grads: Dict[str, torch.Tensor]) -> None:
        rk = self.ranks[rank]
        w = self.worker_ranks.index(rank)
        with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
            for name, g in grads.items():
                s = self.layout[name][0]
                self._view(rk.bufs["glocalgrad%d_w%d" % (s, w)], name).copy_(g)
        self._push_signal(rank)

    def _push_signal(self, rank: int) -> None:
        """The local gradient buffers are filled (stream order): move them to the ps (unicast) or just signal (NVLS: they
        already sit in this worker's copy of the symmetric buffer), one stamp + arrival per shard."""
        rk = self.ranks[rank]
        w = self.worker_ranks.index(rank)
        st = rk.stream.cuda_stream
        with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
            for s in range(self.cfg.num_ps):
                n = self.shard_elems[s]
                if self.nvls:
                    rc = self.lib.dtf_push_grad(None, None, 0, self.peer[(rank, "gctl%d" % s)].ptr,
                                                rk.bufs["gmailbox_w%d" % w].ptr + s * self.mb_bytes, w,
                                                0 if self.cfg.sync else 1, 1, 1, st)
                else:
                    dst = self.peer[(rank, "ggrads%d" % s)].ptr + w * n * 4
                    rc = self.lib.dtf_push_grad(rk.bufs["glocalgrad%d_w%d" % (s, w)].ptr, dst, n, self.peer[(rank, "gctl%d" % s)].ptr,
                                                rk.bufs["gmailbox_w%d" % w].ptr + s * self.mb_bytes, w,
                                                0 if self.cfg.sync else 1, 1, _PUSH_CTAS, st)
                assert rc == 0, rc
            cuda_lib._bump(self.cfg.num_ps)
        rk.step += 1

    def worker_step(self, rank: int, loss_fn: Callable[..., torch.Tensor], *batch, graph: bool = False) -> torch.Tensor:
        """wait token (+ pull) -> forward / backward of ``loss_fn(params, *batch)`` on our kernels -> push (+ signal).
        ``graph=True``: after two eager steps the forward/backward/gradient-staging part is captured ONCE into a CUDA
        graph and replayed (``loss_fn`` must be shape-static and free of host synchronisation); the token wait, the pull
        and the push stay eager because their targets are host-side step numbers."""
        if graph:
            return self._worker_step_graphed(rank, loss_fn, batch)
        rk = self.ranks[rank]
        params = self.pull(rank)
        with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
            leaves = {k: v.detach().requires_grad_(True) for k, v in params.items()}
            loss = loss_fn(leaves, *batch)
            gl = torch.autograd.grad(loss, [leaves[k] for k in self.names], allow_unused=True)
            grads = {k: (g if g is not None else torch.zeros_like(leaves[k])) for k, g in zip(self.names, gl)}
        self.push(rank, grads)
        return loss.detach()

    def _worker_step_graphed(self, rank: int, loss_fn, batch) -> torch.Tensor:
        """One launch-bound model step as ONE graph launch.  The parameter replicas, the local gradient buffers and the
        input staging tensors have fixed addresses, so the captured kernels (TMA descriptors included) stay valid across
        replays; a ResNet-18 step is ~600 small launches + autograd bookkeeping when run eagerly.
        Validated on hardware in round 2 (``bench.py --model resnet18 --graph-step 1``: 11.6 -> 5.4 ms per step at the time)."""
        rk = self.ranks[rank]
        st = self.__dict__.setdefault("_step_graphs", {}).setdefault(rank, {"eager": 0, "graph": None})
        if st["graph"] is None and st["eager"] < 2:
            st["eager"] += 1             # first-launch attribute setup, allocator pools, workspaces: not capturable
            return self.worker_step(rank, loss_fn, *batch)
        params = self.pull(rank)
        w = self.worker_ranks.index(rank)
        with torch.cuda.device(rk.device), torch.cuda.stream(rk.stream):
            if st["graph"] is None:
                st["inputs"] = [torch.empty_like(b, device=rk.device) for b in batch]
                for dst, src in zip(st["inputs"], batch):
                    dst.copy_(src)
                rk.stream.synchronize()
                g = torch.cuda.CUDAGraph()
                before = cuda_lib.launch_count()
                with torch.cuda.graph(g, stream=rk.stream, capture_error_mode="thread_local"):
                    leaves = {k: v.detach().requires_grad_(True) for k, v in params.items()}
                    loss = loss_fn(leaves, *st["inputs"])
                    gl = torch.autograd.grad(loss, [leaves[k] for k in self.names], allow_unused=True)
                    for name, gr in zip(self.names, gl):
                        view = self._view(rk.bufs["glocalgrad%d_w%d" % (self.layout[name][0], w)], name)
                        if gr is None:
                            view.zero_()
                        else:
                            view.copy_(gr)
                    st["loss"] = loss.detach()
                st["kernels"] = cuda_lib.launch_count() - before
                cuda_lib._bump(-st["kernels"])          # the capture itself launched nothing
                st["graph"] = g
            else:
                for dst, src in zip(st["inputs"], batch):
                    dst.copy_(src, non_blocking=True)
            st["graph"].replay()
            cuda_lib._bump(st["kernels"])
        self._push_signal(rank)
        return st["loss"]

    def ps_apply(self, rank: int, idle_ok: bool = False) -> None:
        rk = self.ranks[rank]
        a = self._p[rank]
        a.idle_ok = int(idle_ok)
        if idle_ok:
            a.timeout_ns = 20_000_000          # service loop: come back to the host every 20 ms when nothing arrives
        with torch.cuda.device(rk.device):
            rc = self.lib.dtf_ps_apply(ctypes.byref(a), rk.stream.cuda_stream)
        assert rc == 0, rc
        cuda_lib._bump()

    def step(self, loss_fn: Callable[..., torch.Tensor], batches: Dict[int, Tuple]) -> Dict[int, torch.Tensor]:
        """One step for every local rank; ``batches[rank]`` is the worker's batch tuple."""
        losses = {}
        for r in self.worker_ranks:
            if r in self.ranks:
                losses[r] = self.worker_step(r, loss_fn, *batches[r])
        for r in self.ps_ranks:
            if r in self.ranks:
                for _ in range(1 if self.cfg.sync else self.cfg.num_workers):
                    self.ps_apply(r)
        return losses

    def check_errors(self) -> None:
        for r, rk in self.ranks.items():
            rk.stream.synchronize()
            if r in self.worker_ranks:
                w = self.worker_ranks.index(r)
                e = int(rk.bufs["gmisc_w%d" % w].tensor(torch.int32, 16, 1).cpu()[0])
                if e:
                    raise RuntimeError("worker %d: device-side wait timed out (code %d)" % (w, e))
            if r in self.ps_ranks:
                s = self.ps_ranks.index(r)
                e = int(rk.bufs["gctl%d" % s].tensor(torch.int32, self.off["err"], 1).cpu()[0])
                if e:
                    raise RuntimeError("ps shard %d: device-side wait timed out (code %d)" % (s, e))

    def synchronize(self) -> None:
        for rk in self.ranks.values():
            rk.stream.synchronize()

    def close(self) -> None:
        self.synchronize()
        self.fabric.close()
