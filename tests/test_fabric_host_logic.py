"""Host-side logic of the fabric tier that does not need a GPU: store binding, rank mapping, strategy graph
construction, engine layouts (SURVEY A5 placement carried into the fabric)."""
import pytest
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.framework.executor import ResourceStore
from distributed_tensorflow_b200.parallel.ps_engine import MLPSpec, _layout
from distributed_tensorflow_b200.parallel.strategy import fabric_rank_of


def test_store_bind_keeps_storage_and_tracks_initialisation():
    st = ResourceStore()
    buf = torch.zeros(8)
    view = buf[2:6].view(2, 2)
    st.bind("w", view, initialized=False)
    assert not st.is_initialized("w")
    with pytest.raises(dtf.errors.FailedPreconditionError):
        st.read("w")
    st.assign("w", torch.tensor([[1.0, 2.0], [3.0, 4.0]]))
    assert st.is_initialized("w") and buf.tolist() == [0, 0, 1, 2, 3, 4, 0, 0]     # written THROUGH the binding
    # a value assigned before the fabric came up is carried into the bound storage
    st.assign("b", torch.tensor([7.0, 8.0]))
    buf2 = torch.zeros(2)
    st.bind("b", buf2, initialized=False)
    assert st.is_initialized("b") and buf2.tolist() == [7.0, 8.0]
    st.assign("b", torch.tensor([1.0, 1.0]))
    assert buf2.tolist() == [1.0, 1.0] and st.read("b") is buf2
    # 0-d int64 (global_step living in a control block)
    ctl = torch.zeros(4, dtype=torch.int64)
    st.bind("global_step", ctl[1:2].view(()), initialized=False)
    st.assign("global_step", torch.tensor(41))
    assert int(ctl[1]) == 41


def test_fabric_rank_mapping_ps_first():
    c = dtf.train.ClusterSpec({"worker": ["h:3", "h:4"], "ps": ["h:1", "h:2"]})
    assert [fabric_rank_of(c, j, t)[0] for j, t in (("ps", 0), ("ps", 1), ("worker", 0), ("worker", 1))] == [0, 1, 2, 3]
    assert fabric_rank_of(c, "worker", 1)[1] == 4


def test_mlp_engine_layout_follows_round_robin_placement():
    lay, sizes = _layout(MLPSpec(), 2)
    # creation order global_step, hid_w, hid_b, sm_w, sm_b over 2 ps -> ps0, ps1, ps0, ps1, ps0
    assert {k: v.shard for k, v in lay.items()} == {"hid_w": 1, "hid_b": 0, "sm_w": 1, "sm_b": 0}
    assert lay["hid_w"].pitch % 8 == 0 and lay["hid_w"].pitch >= 100          # 16-byte bf16 rows for TMA
    assert all(v.offset % 64 == 0 for v in lay.values())
    lay1, sizes1 = _layout(MLPSpec(hidden=64), 1)
    assert sizes1[0] >= 784 * 64 + 64 + 64 * 16 + 16


def test_strategy_minimize_builds_fabric_train_step(ports):
    p = ports(3)
    cluster = dtf.train.ClusterSpec({"ps": ["127.0.0.1:%d" % p[0], "127.0.0.1:%d" % p[1]], "worker": ["127.0.0.1:%d" % p[2]]})
    server = dtf.train.Server(cluster, "worker", 0)
    try:
        strategy = dtf.fabric.FabricPSStrategy(server)
        with dtf.device(dtf.train.replica_device_setter(cluster=cluster, worker_device="/job:worker/task:0/cpu:0")):
            gs = dtf.train.get_or_create_global_step()
            w = dtf.Variable(dtf.zeros([4, 3]), name="w")
            b = dtf.Variable(dtf.zeros([3]), name="b")
            x = dtf.placeholder(dtf.float32, [None, 4])
            loss = dtf.reduce_sum(dtf.square(dtf.nn.xw_plus_b(x, w, b)))
            opt = dtf.train.SyncReplicasOptimizer(dtf.train.MomentumOptimizer(0.1, 0.9), 1, 1)
            train_op, floss = strategy.minimize(opt, loss, gs)
        spec = strategy._spec
        assert [(n, s, sh) for n, s, sh in spec["params"]] == [("w", [4, 3], 1), ("b", [3], 0)]   # w->ps1, b->ps0
        assert spec["optimizer"]["kind"] == "momentum" and spec["optimizer"]["sync"] and spec["global_step"] == "global_step"
        assert spec["key"].startswith("f") and len(spec["key"]) == 13      # content hash: same in every worker process
        assert train_op.op_type == "FabricTrainStep" and [i.op_type for i in train_op.inputs] == ["Placeholder"]
        assert train_op.device == "/job:worker/task:0/device:CPU:0"
        with pytest.raises(RuntimeError):
            dtf.fabric.FabricPSStrategy(dtf.train.Server.create_local_server()).minimize(opt, loss, gs)
    finally:
        server.stop()


def test_fd_passing_between_processes():
    """VMM / multicast handles travel between task processes as POSIX fds over a unix socket (parallel/fdshare.py)."""
    import multiprocessing as mp
    import os
    from distributed_tensorflow_b200.parallel.fdshare import FdServer, fetch_fd
    srv = FdServer()
    r, w = os.pipe()
    os.write(w, b"sym-buffer")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_fd_child, args=(srv.path, q))
    p.start()
    srv.register("srepl0/mem/1", r)            # registered AFTER the client connected: the server waits for it
    assert q.get(timeout=60) == b"sym-buffer"
    p.join(timeout=30)
    with pytest.raises(Exception):
        fetch_fd(srv.path + ".nope", "x", timeout=0.3)
    srv.close()


def _fd_child(path, q):
    import os
    from distributed_tensorflow_b200.parallel.fdshare import fetch_fd
    fd = fetch_fd(path, "srepl0/mem/1")
    q.put(os.read(fd, 10))


def test_native_step_plan_layout_and_host_ops():
    """The ctypes mirror of DtfStepOp matches the C struct, and a plan of host-side ops reports failures by index."""
    import ctypes
    from distributed_tensorflow_b200.ops import cuda_lib
    if not cuda_lib.available():
        pytest.skip("kernel library not built")
    lib = cuda_lib.load()
    assert lib.dtf_sizeof_step_op() == ctypes.sizeof(cuda_lib.StepOp)
    bad = cuda_lib.StepPlan([cuda_lib.StepOp(kind=99)], -1, None)
    with pytest.raises(RuntimeError, match="op 0"):
        bad.run()


def test_deferred_loss_step_ordering_with_fake_plans():
    """``step(..., sync_loss="deferred")`` host logic without a GPU: plans are replaced by recorders.  Per step: the copy
    plan of this parity (unless prefetched), the compute plan, the prefetch copy of the other parity, then the loss
    read-back plan of this parity; a handle that was never read is resolved before its landing buffer is reused."""
    import numpy as np
    from types import SimpleNamespace
    from distributed_tensorflow_b200.parallel.ps_engine import PendingLoss, PSTrainEngine

    log = []

    class Plan:
        def __init__(self, name):
            self.name, self.ops = name, [SimpleNamespace(p1=None) for _ in range(3)]

        def run(self):
            log.append(self.name)

    class Event:
        def __init__(self, name):
            self.name, self.syncs = name, 0

        def query(self):
            return False

        def synchronize(self):
            self.syncs += 1
            log.append("sync:" + self.name)

    eng = PSTrainEngine.__new__(PSTrainEngine)
    eng.cfg = SimpleNamespace(sync=True, num_workers=1, colocated=True)
    eng.spec = SimpleNamespace(batch=4, in_dim=8, classes=2)
    eng.worker_ranks, eng.ps_ranks, eng.head_ctas = [0], [0], 2
    eng.ranks = {0: SimpleNamespace(step=0)}
    eng._is_pinned_f32 = lambda t: True
    hosts = [np.array([1.0, 2.0, 50.0], dtype=np.float32), np.array([10.0, 20.0, 50.0], dtype=np.float32)]
    events = [Event("l0"), Event("l1")]
    eng._native_plans = {"copy": {0: [Plan("copy0"), Plan("copy1")]}, "compute": {0: [Plan("compute0"), Plan("compute1")]},
                         "ps": {}, "loss": {}, "pending": {}, "keep": [], "runs": 10, "parity": 0, "prefetched": None, "graphed": True,
                         "loss_async": {0: [(Plan("loss0"), hosts[0], None, events[0]), (Plan("loss1"), hosts[1], None, events[1])]}}
    import torch
    xs = [torch.zeros(4, 8) + i for i in range(4)]
    ys = [torch.zeros(4, 2) for _ in range(4)]
    h0 = eng.step(xs[0], ys[0], sync_loss="deferred", prefetch=(xs[1], ys[1]))
    assert isinstance(h0, PendingLoss) and not h0.done()
    assert log == ["copy0", "compute0", "copy1", "loss0"]
    h1 = eng.step(xs[1], ys[1], sync_loss="deferred", prefetch=(xs[2], ys[2]))          # batch 1 was prefetched: no second copy
    assert log[4:] == ["compute1", "copy0", "loss1"]
    assert h0.result() == 3.0 and log[-1] == "sync:l0" and float(h0) == 3.0 and events[0].syncs == 1
    del log[:]
    h2 = eng.step(xs[2], ys[2], sync_loss="deferred")                                    # parity 0 again; h0 already read
    assert log == ["compute0", "loss0"]
    del log[:]
    hosts[1][:2] = [7.0, 8.0]
    h3 = eng.step(xs[3], ys[3], sync_loss="deferred")          # parity 1: h1 was never read -> resolved before reuse
    assert log == ["copy1", "compute1", "sync:l1", "loss1"]
    assert h1.result() == 15.0 and h1.done()
    assert h2.result() == 3.0 and h3.result() == 15.0
    assert eng.ranks[0].step == 4


def test_train_loop_host_logic_with_a_fake_native_loop():
    """``PSTrainEngine.train_loop`` without a GPU: the first steps of a process go through ``step()`` until the plans are
    graphed, the rest is ONE ``dtf_run_loop`` call whose arguments (plans, streams, batch walk, parity, prefetch state, loss
    rows) are checked here; the loss rows the native loop fills come back summed per step."""
    import ctypes
    import numpy as np
    import torch
    from types import SimpleNamespace
    from distributed_tensorflow_b200.ops.cuda_lib import LoopArgs, StepOp
    from distributed_tensorflow_b200.parallel.ps_engine import PSTrainEngine

    calls = []

    class Plan:
        def __init__(self, n, stream):
            self.n, self.ops, self.stream = n, (StepOp * n)(), stream

    class Lib:
        def dtf_run_loop(self, ref):
            a = ctypes.cast(ref, ctypes.POINTER(LoopArgs)).contents
            calls.append({k: getattr(a, k) for k in ("device", "steps", "depth", "parity", "prefetched", "x_op", "y_op", "n_ps", "copy_stream",
                                                     "stream", "ps_stream", "x_base", "y_base", "x_stride", "y_stride", "nbatches", "first",
                                                     "batch_step", "loss_src", "loss_bytes", "loss_row_bytes")})
            calls[-1]["n_copy"], calls[-1]["copy_ops"] = list(a.n_copy), list(a.copy_ops)
            rows = np.ctypeslib.as_array(ctypes.cast(a.loss_host, ctypes.POINTER(ctypes.c_float)), shape=(a.steps, 16))
            for i in range(a.steps):
                rows[i, :] = 100.0                       # beyond head_ctas: must not be summed
                rows[i, :2] = [i, 0.5]
            a.kernels, a.parity = 2 * a.steps, (a.parity + a.steps) & 1
            return 0

    eng = PSTrainEngine.__new__(PSTrainEngine)
    eng.cfg = SimpleNamespace(sync=True, num_workers=2, colocated=False)
    eng.spec = SimpleNamespace(batch=4, in_dim=8, classes=2)
    eng.worker_ranks, eng.ps_ranks, eng.head_ctas, eng.lib = [0, 1], [0], 2, Lib()
    eng.ranks = {1: SimpleNamespace(step=0, device=SimpleNamespace(index=3))}
    eng._w = {1: {"loss_ptr": 0xbeef}}
    eng._is_pinned_f32 = lambda t: isinstance(t, torch.Tensor)
    cp, cm, ps = [Plan(4, 0x58), Plan(4, 0x58)], [Plan(3, 0x57), Plan(3, 0x57)], Plan(1, 0x59)
    stepped = []
    eng._native_plans = plans = {"copy": {1: cp}, "compute": {1: cm}, "ps": {1: ps}, "loss": {}, "pending": {}, "keep": [], "runs": 2,
                                 "parity": 0, "prefetched": None, "loss_async": {}}

    def fake_step(x, y, sync_loss=True, **kw):
        stepped.append(int(x[0, 0]))
        plans["runs"] += 1
        plans["parity"] ^= 1
        if plans["runs"] == 4:
            plans["graphed"] = True
        return 7.0
    eng.step = fake_step
    xb = torch.arange(5, dtype=torch.float32).view(5, 1, 1).expand(5, 4, 8).contiguous()
    yb = torch.zeros(5, 4, 2)
    out = eng.train_loop(xb, yb, steps=6, first=3, stride=2, depth=3)
    assert stepped == [3, 0]                              # batches (3 + 2 i) % 5 of the two eager steps
    assert len(calls) == 1
    c = calls[0]
    assert (c["device"], c["steps"], c["depth"], c["parity"], c["prefetched"], c["x_op"], c["y_op"]) == (3, 4, 3, 0, 0, 1, 2)
    assert (c["n_ps"], c["copy_stream"], c["stream"], c["ps_stream"]) == (1, 0x58, 0x57, 0x59)
    assert (c["x_base"], c["y_base"], c["x_stride"], c["y_stride"]) == (xb.data_ptr(), yb.data_ptr(), 4 * 8 * 4, 4 * 2 * 4)
    assert (c["nbatches"], c["first"], c["batch_step"]) == (5, (3 + 2 * 2) % 5, 2)
    assert (c["loss_src"], c["loss_bytes"], c["loss_row_bytes"]) == (0xbeef, 8, 64)
    assert c["n_copy"] == [4, 4] and c["copy_ops"] == [ctypes.addressof(cp[0].ops), ctypes.addressof(cp[1].ops)]
    np.testing.assert_allclose(out, [7.0, 7.0, 0.5, 1.5, 2.5, 3.5])
    assert eng.ranks[1].step == 4 and plans["runs"] == 8 and plans["parity"] == 0 and plans["prefetched"] is None
    # a batch that step() prefetched for exactly this parity is not copied again
    plans["prefetched"] = (xb[1].data_ptr(), yb[1].data_ptr(), 0)
    eng.train_loop(xb, yb, steps=1, first=1)
    assert calls[-1]["prefetched"] == 1 and calls[-1]["steps"] == 1
    # prefetch_next: the native loop reports that the batch after the last step is travelling; step() must recognise it
    class Lib2(Lib):
        def dtf_run_loop(self, ref):
            rc = Lib.dtf_run_loop(self, ref)
            a = ctypes.cast(ref, ctypes.POINTER(LoopArgs)).contents
            a.prefetched = a.prefetch_next
            return rc
    eng.lib = Lib2()
    par0 = plans["parity"]
    eng.train_loop(xb, yb, steps=3, first=4, stride=2, prefetch_next=True)
    bn = (4 + 3 * 2) % 5
    assert plans["prefetched"] == (xb[bn].data_ptr(), yb[bn].data_ptr(), (par0 + 3) & 1)
    eng.lib = Lib()
    plans["prefetched"] = None
    # several local workers (in-graph replication) or unpinned inputs: step() per step
    eng.ranks[0] = SimpleNamespace(step=0, device=SimpleNamespace(index=2))
    n = len(calls)
    out = eng.train_loop(xb, yb, steps=3)
    assert len(calls) == n and list(out) == [7.0] * 3


def test_untimed_alignment_step_with_prefetch_hands_the_next_batch_to_the_following_step():
    """bench.py's alignment steps call ``step(x, y, sync_loss=False, prefetch=next)``: no loss read-back, the next batch's copy is
    issued into the other buffer set and the following step (or ``train_loop``) skips its own copy."""
    from types import SimpleNamespace
    import torch
    from distributed_tensorflow_b200.parallel.ps_engine import PSTrainEngine
    log = []

    class Plan:
        def __init__(self, name):
            self.name, self.ops = name, [SimpleNamespace(p1=None) for _ in range(3)]

        def run(self):
            log.append(self.name)
    eng = PSTrainEngine.__new__(PSTrainEngine)
    eng.cfg = SimpleNamespace(sync=True, num_workers=1, colocated=True)
    eng.spec = SimpleNamespace(batch=4, in_dim=8, classes=2)
    eng.worker_ranks, eng.ps_ranks, eng.head_ctas = [0], [0], 2
    eng.ranks = {0: SimpleNamespace(step=0)}
    eng._is_pinned_f32 = lambda t: True
    eng._native_plans = {"copy": {0: [Plan("copy0"), Plan("copy1")]}, "compute": {0: [Plan("compute0"), Plan("compute1")]},
                         "ps": {}, "loss": {0: (Plan("loss"), __import__("numpy").array([1.0, 2.0], dtype="float32"), None)},
                         "pending": {}, "keep": [], "runs": 10, "parity": 0, "prefetched": None, "graphed": True, "loss_async": {}}
    xs = [torch.zeros(4, 8) + i for i in range(3)]
    ys = [torch.zeros(4, 2) for _ in range(3)]
    assert eng.step(xs[0], ys[0], sync_loss=False, prefetch=(xs[1], ys[1])) is None
    assert log == ["copy0", "compute0", "copy1"]
    assert eng._native_plans["prefetched"] == (xs[1].data_ptr(), ys[1].data_ptr(), 1)
    del log[:]
    assert eng.step(xs[1], ys[1], sync_loss=True, prefetch=(xs[2], ys[2])) == 3.0          # the prefetched batch: no copy1 again
    assert log == ["compute1", "copy0", "loss"]


def test_optimizer_state_round_trip_of_the_mlp_engine_without_a_gpu():
    """``PSTrainEngine.optimizer_state`` is the non-variable part of ``state_dict`` (slots under TF's names, beta powers);
    ``load_optimizer_state`` writes it back into the shard that owns each variable and skips foreign / mis-shaped entries."""
    from types import SimpleNamespace
    from distributed_tensorflow_b200.parallel.ps_engine import PSTrainEngine

    def make():
        eng = PSTrainEngine.__new__(PSTrainEngine)
        store = {("master", "hid_w"): torch.arange(6.0).reshape(2, 3), ("slot_m", "hid_w"): torch.zeros(2, 3),
                 ("slot_v", "hid_w"): torch.zeros(2, 3), ("master", "sm_b"): torch.ones(3), ("slot_m", "sm_b"): torch.zeros(3),
                 ("slot_v", "sm_b"): torch.zeros(3)}
        ctl = torch.zeros(8)
        ctl[2], ctl[3] = 0.9, 0.999
        rk = SimpleNamespace(device=None, stream=None, sync=lambda: None,
                             bufs={"ctl0": SimpleNamespace(tensor=lambda dt, off, n: ctl[off:off + n])})
        eng.ranks, eng.ps_ranks, eng.kind = {0: rk}, [0], 2
        eng.off = {"beta1_power": 2, "global_step": 0}
        eng.layout = {"hid_w": SimpleNamespace(shard=0, shape=(2, 3), rows=2, cols=3, name="hid_w"),
                      "sm_b": SimpleNamespace(shard=0, shape=(3,), rows=1, cols=3, name="sm_b"),
                      "other": SimpleNamespace(shard=1, shape=(3,), rows=1, cols=3, name="other")}
        # like the engine's: a [rows, cols] window of a PITCHED buffer (writes must go through the view, a reshape would copy)
        pitched = {k: torch.zeros(lay_rows, 8) for k, lay_rows in ((("slot_m", "hid_w"), 2), (("slot_v", "hid_w"), 2), (("master", "hid_w"), 2),
                                                                   (("slot_m", "sm_b"), 1), (("slot_v", "sm_b"), 1), (("master", "sm_b"), 1))}
        for k, buf in pitched.items():
            buf[:, :3] = store[k].reshape(buf.shape[0], 3)
            store[k] = buf[:, :3]
        eng._var_view = lambda rk_, base, lay: store[(base, lay.name)]
        eng.read_ctl = lambda shard, fld, count=1: 41
        return eng, store, ctl
    a, sa, ca = make()
    sa[("slot_m", "hid_w")] += 0.25
    sa[("slot_v", "sm_b")] += 4.0
    ca[2] = 0.9 ** 5
    st = a.optimizer_state()
    assert sorted(st) == ["beta1_power", "beta2_power", "hid_w/Adam", "hid_w/Adam_1", "sm_b/Adam", "sm_b/Adam_1"]      # no variables, no global_step
    b, sb, cb = make()
    st["other/Adam"] = torch.ones(3)                          # owned by another shard
    st["sm_b/Adam"] = torch.ones(5)                           # wrong shape
    done = b.load_optimizer_state(st)
    assert "other/Adam" not in done and "sm_b/Adam" not in done and "beta1_power" in done
    assert sb[("slot_m", "hid_w")].tolist() == [[0.25] * 3] * 2 and sb[("slot_v", "sm_b")].reshape(-1).tolist() == [4.0] * 3
    assert sb[("slot_m", "sm_b")].reshape(-1).tolist() == [0.0] * 3 and float(cb[2]) == pytest.approx(0.9 ** 5) and float(cb[3]) == pytest.approx(0.999)
    assert sb[("master", "hid_w")].tolist() == torch.arange(6.0).reshape(2, 3).tolist()       # variables untouched
