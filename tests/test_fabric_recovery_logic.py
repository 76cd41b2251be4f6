"""Host-side logic of the fabric tier's failure handling (no GPU): fabric generations on ps task 0, variables moving out of
and back into engine storage, the strategy's abort path (AbortedError / UnavailableError, reference
example_between_graph.py:99 "handles AbortedError in case of preempted PS") and the per-generation process group."""
import os
import socket

import pytest
import torch

import distributed_tensorflow_b200 as dtf
from distributed_tensorflow_b200.framework import errors
from distributed_tensorflow_b200.framework.executor import ResourceStore
from distributed_tensorflow_b200.parallel import strategy as S
from distributed_tensorflow_b200.parallel.cluster import ClusterSpec


def _free_ports(n):
    socks = [socket.socket() for _ in range(n)]
    for s in socks:
        s.bind(("127.0.0.1", 0))
    ports = [s.getsockname()[1] for s in socks]
    for s in socks:
        s.close()
    return ports


def test_generation_authority_rules():
    p = _free_ports(1)[0]
    srv = dtf.train.Server(ClusterSpec({"ps": ["127.0.0.1:%d" % p]}), "ps", 0)
    try:
        g = srv.rpc_fabric_generation
        assert g("job", 0, 1) == 0 and g("job", 0, 2) == 0 and srv.rpc_fabric_current_generation("job") == 0
        # two survivors of one failure at generation 0 ask for 1: both get 1
        assert g("job", 1, 1) == 1 and g("job", 1, 2) == 1
        # a restarted task asks with at_least = 0: it is already a member of generation 1 -> the job moves to 2 ...
        assert g("job", 0, 2) == 2
        # ... which the surviving member learns from its liveness call and then joins
        assert srv.rpc_fabric_current_generation("job") == 2 and g("job", 2, 1) == 2
        # a restarted task that was NOT yet a member of the current generation simply joins it
        assert g("job", 0, 3) == 2
        assert g("other", 0, 1) == 0                       # per job
    finally:
        srv.stop()


def test_unbind_keeps_the_value_and_rebind_carries_it_over():
    store = ResourceStore()
    engine_buf = torch.zeros(8)
    store.assign("w", torch.arange(4.0))                  # the chief initialised before the fabric came up
    store.bind("w", engine_buf[:4], initialized=False)
    assert engine_buf[:4].tolist() == [0.0, 1.0, 2.0, 3.0] and store.is_initialized("w")
    engine_buf[:4] += 10.0                                # the ps applied updates in engine storage
    store.unbind("w")
    engine_buf.fill_(-1.0)                                # the engine's memory is released / reused
    assert store.read("w").tolist() == [10.0, 11.0, 12.0, 13.0] and store.is_initialized("w")
    new_buf = torch.zeros(4)
    store.bind("w", new_buf, initialized=False)           # generation + 1
    assert new_buf.tolist() == [10.0, 11.0, 12.0, 13.0]


class _FakeEngine:
    def __init__(self, err=None):
        self.err, self.closed = err, False
        self.cfg = type("C", (), {"timeout_ns": int(2e9)})()

    def check_errors(self):
        if self.err:
            raise RuntimeError(self.err)

    def close(self):
        self.closed = True


class _FakeServer:
    job_name, task_index, gpu_index = "worker", 0, 0

    def __init__(self):
        self.cluster = ClusterSpec({"ps": ["127.0.0.1:1", "127.0.0.1:2"], "worker": ["127.0.0.1:3", "127.0.0.1:4"]})


def _strategy(calls, unreachable=()):
    st = S.FabricPSStrategy(_FakeServer())
    st._spec_base = {"key": "fjob", "optimizer": {"sync": True}}
    st._spec, st._gen = dict(st._spec_base), 0

    def ps_call(t, method, *args):
        calls.append((t, method) + args)
        if t in unreachable:
            raise ConnectionRefusedError("ps %d" % t)
        return {"fabric_current_generation": st.__dict__.get("_reported_gen", 0)}.get(method, True)
    st._ps_call = ps_call
    return st


def test_a_timed_out_device_wait_becomes_aborted_error_and_the_next_generation():
    import time
    calls = []
    st = _strategy(calls)
    eng = st.engine = _FakeEngine(err="worker 0: device-side wait timed out (code 1)")
    st._last_live = time.time()
    st._after_step(time.time())                           # a fast step: nothing is read, nothing raised
    assert not calls and st.engine is eng
    with pytest.raises(errors.AbortedError, match="generation 0 .* timed out"):
        st._after_step(time.time() - 1.5)                 # a step that sat in a device-side timeout
    assert eng.closed and st.engine is None and st._min_gen == 1 and st.aborts == 1
    assert [c[:2] for c in calls] == [(0, "fabric_teardown"), (1, "fabric_teardown")] and calls[0][2] == "fjob"


def test_an_unreachable_ps_is_unavailable_error_and_a_newer_generation_aborts_too():
    import time
    calls = []
    st = _strategy(calls, unreachable=(1,))
    st.engine = _FakeEngine(err="worker 0: device-side wait timed out (code 1)")
    st._last_live = time.time()
    with pytest.raises(errors.UnavailableError, match=r"ps task\(s\) \[1\] unreachable"):
        st._after_step(time.time() - 1.5)
    # liveness: ps task 0 reports that a restarted task moved the job to generation 3
    st = _strategy(calls)
    st.engine, st._gen, st._reported_gen = _FakeEngine(), 2, 3
    st._last_live = time.time() - 10.0
    with pytest.raises(errors.AbortedError, match="moved on to fabric generation 3"):
        st._after_step(time.time())
    assert st._min_gen == 3
    # both are what MonitoredTrainingSession recovers from
    from distributed_tensorflow_b200.train.monitored_session import _RECOVERABLE
    assert errors.AbortedError in _RECOVERABLE and errors.UnavailableError in _RECOVERABLE


def test_process_group_follows_the_generation(monkeypatch):
    import torch.distributed as dist
    if dist.is_initialized():
        pytest.skip("a process group of this test process is already up")
    base = _free_ports(1)[0]
    cluster = ClusterSpec({"ps": ["127.0.0.1:%d" % base]})
    monkeypatch.setenv("DTF_FABRIC_PORT_OFFSET", "0")
    monkeypatch.setenv("DTF_FABRIC_TIMEOUT", "30")
    try:
        assert S.init_fabric_process_group(cluster, "ps", 0, 0) == (0, 1) and S._PG_GEN[0] == 0
        pg0 = dist.distributed_c10d._get_default_group()
        assert S.init_fabric_process_group(cluster, "ps", 0, 0) == (0, 1)
        assert dist.distributed_c10d._get_default_group() is pg0              # same generation: idempotent
        S.init_fabric_process_group(cluster, "ps", 0, 1)                          # new generation: new group, port + 1
        assert S._PG_GEN[0] == 1 and dist.is_initialized() and dist.distributed_c10d._get_default_group() is not pg0
        dist.barrier()
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()
        S._PG_GEN[0] = None


def test_optimizer_slots_live_through_a_fabric_re_formation_on_the_ps_task(monkeypatch):
    """Slots and Adam's beta powers exist only in the engine's HBM.  When the fabric re-forms (a worker died, the ps task
    survived) the ps service of the old generation hands them to the next one, as the variable store does for the
    variables: the optimizer continues instead of restarting its moments from zero."""
    import time
    from types import SimpleNamespace
    from distributed_tensorflow_b200.framework.executor import ResourceStore
    built = []

    class Eng:
        def __init__(self):
            self.ranks = {0: SimpleNamespace(stream=SimpleNamespace(synchronize=lambda: None))}
            self.ps_ranks, self.cfg = [0], SimpleNamespace(sync=True, num_workers=2)
            self.layout = {"hid_w": SimpleNamespace(shard=0)}
            self.w, self.gs = torch.zeros(3), torch.zeros((), dtype=torch.int64)
            self.slots = {"hid_w/Adam": torch.zeros(3), "hid_w/Adam_1": torch.zeros(3), "beta1_power": torch.tensor(0.9),
                          "beta2_power": torch.tensor(0.999)}
            self.loaded, self.closed = None, False
            built.append(self)

        def var_tensor(self, rank, role):
            return self.w

        def global_step_tensor(self, rank):
            return self.gs

        def ps_apply(self, rank, idle_ok=False):
            time.sleep(0.005)

        def optimizer_state(self):
            return {k: v.clone() for k, v in self.slots.items()}

        def load_optimizer_state(self, st):
            self.loaded = {k: v.clone() for k, v in st.items()}
            return sorted(st)

        def close(self):
            self.closed = True
    monkeypatch.setattr(S, "build_engine", lambda *a, **k: Eng())
    srv = SimpleNamespace(cluster=None, job_name="ps", task_index=0, gpu_index=0, is_running=True, task=("ps", 0),
                          store=ResourceStore("/job:ps/task:0"))
    spec0 = {"key": "fjob", "base_key": "fjob", "mlp": {"roles": {"hid_w": "hid_w:0"}}, "global_step": "global_step:0"}
    assert S.ps_fabric_setup(srv, spec0)
    e0 = built[0]
    assert e0.loaded is None                                   # first generation: nothing to carry over
    srv.store.assign("hid_w:0", torch.zeros(3))                # the chief's init op
    e0.w += 5.0                                                # training moved the variable and the moments
    e0.slots["hid_w/Adam"] += 1.5
    e0.slots["hid_w/Adam_1"] += 2.5
    e0.slots["beta1_power"] = torch.tensor(0.9 ** 7)
    spec1 = dict(spec0, key="fjob@1")
    assert S.ps_fabric_setup(srv, spec1)                       # generation 1 retires generation 0 on this task
    e1 = built[1]
    assert e0.closed and "fabric_service/fjob" not in srv.store.resources
    assert e1.loaded["hid_w/Adam"].tolist() == [1.5] * 3 and e1.loaded["hid_w/Adam_1"].tolist() == [2.5] * 3
    assert float(e1.loaded["beta1_power"]) == pytest.approx(0.9 ** 7)
    assert e1.w.tolist() == [5.0] * 3                          # the variable itself came through the store's rebind
    assert "fabric_opt_state/fjob" not in srv.store.resources  # consumed
    S.ps_fabric_teardown(srv, "fjob")
    assert e1.closed and "fabric_opt_state/fjob" in srv.store.resources     # kept for a later generation
