"""bench.py's end-to-end arms (``run_e2e``) against a stand-in engine on CPU: every arm runs, walks the batches it claims to
walk, keeps the step counter continuous across the untimed / timed boundary, reports the best arm, and a failing native loop is
dropped without losing the other arms' numbers.  (The real engine's ``step`` / ``train_loop`` are covered by
tests/test_fabric_host_logic.py, tests/test_step_exec_host.py and the hardware tiers.)"""
import argparse
import importlib.util
import math
import os
import time
from types import SimpleNamespace

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec_ = importlib.util.spec_from_file_location("dtf_bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec_)
spec_.loader.exec_module(bench)


class FakePending:
    def __init__(self, v):
        self.v = v

    def result(self):
        return self.v


class FakeEngine:
    """Costs: step() 30 us of host time, train_loop 10 us per step -- so the native arm must win."""

    def __init__(self, nb, fail_native=False):
        self.ranks = {0: SimpleNamespace(device=None, stream=None)}
        self.worker_ranks, self.head_ctas, self.nb = [0], 7, nb
        self.seen, self.prefetches, self.fail_native = [], [], fail_native
        self.loop_calls = []

    def _spin(self, us):
        t = time.perf_counter()
        while (time.perf_counter() - t) * 1e6 < us:
            pass

    def step(self, x=None, y=None, sync_loss=True, prefetch=None, source="dataset"):
        self._spin(30)
        if x is None:                                # a ps-only process: one apply per step, no batch
            self.seen.append(-1)
            return None
        b = int(x[0, 0])
        self.seen.append(b)
        if prefetch is not None:
            self.prefetches.append((b, int(prefetch[0][0, 0])))
        loss = 1.0 / (1 + len(self.seen))
        if sync_loss == "deferred":
            return FakePending(loss)
        return loss if sync_loss else None

    def train_loop(self, xb, yb, steps, first=0, stride=1, depth=2, prefetch_next=False):
        if self.fail_native:
            raise RuntimeError("native loop exploded")
        assert tuple(xb.shape[1:]) == (100, 784) and xb.shape[0] == self.nb and depth == 4 and prefetch_next
        self.loop_calls.append((steps, first, stride))
        out = []
        for i in range(steps):
            self._spin(10)
            self.seen.append(int(xb[(first + i * stride) % self.nb, 0, 0]))
            out.append(1.0 / (1 + len(self.seen)))
        return np.asarray(out, np.float32)

    def join_streams(self):
        pass

    def synchronize(self):
        pass

    def check_errors(self):
        pass


class WallTimer:
    def start(self, eng):
        return time.perf_counter()

    def stop(self, t0, eng):
        return (time.perf_counter() - t0) * 1e3


def _run(fail_native=False, K=20):
    nb = 50
    images = np.repeat(np.arange(nb, dtype=np.float32), 100)[:, None] * np.ones((1, 784), np.float32)     # row value = its batch index
    labels = np.zeros((nb * 100, 10), np.float32)
    args = argparse.Namespace(num_train=nb * 100, e2e_prefetch=1, e2e_pipeline=1, e2e_native_loop=1, e2e_depth=4, min_ms=5.0, max_reps=50)
    eng = FakeEngine(nb, fail_native)
    spec = SimpleNamespace(batch=100, in_dim=784, classes=10)
    e2e = bench.run_e2e(eng, args, spec, K, 1, True, 1, False, images, labels, barrier=lambda: None,
                        allmax=lambda v: [float(x) for x in v], timer=WallTimer(), pin=torch.from_numpy)
    return e2e, eng


def test_every_arm_runs_walks_consecutive_batches_and_the_best_one_is_reported():
    e2e, eng = _run()
    assert {"synchronous", "pipelined", "native_loop"} <= set(e2e) and "native_loop_error" not in e2e
    assert e2e["native_loop"]["value"] > e2e["synchronous"]["value"] and e2e["value"] == e2e["native_loop"]["value"]
    assert e2e["api"].startswith("PSTrainEngine.train_loop(") and e2e["native_loop"]["depth"] == 4
    assert e2e["h2d_bytes_per_step"] == 100 * 794 * 4 and e2e["d2h_bytes_per_step"] == 28 and e2e["steps"] == 20
    assert math.isfinite(e2e["last_loss"]) and e2e["reps"] >= 3
    # one continuous walk over the batches: step t trains on batch t % nb, across alignment steps, repetitions and arms
    assert eng.seen == [t % 50 for t in range(len(eng.seen))]
    assert eng.prefetches and all(p == (b + 1) % 50 for b, p in eng.prefetches)     # every prefetch names the next batch
    assert eng.loop_calls and all(c[0] == 20 and c[2] == 1 for c in eng.loop_calls)


def test_a_failing_native_loop_is_dropped_and_the_other_arms_stand():
    e2e, eng = _run(fail_native=True)
    assert "native_loop" not in e2e and "native loop exploded" in e2e["native_loop_error"]
    assert e2e["value"] == max(e2e["synchronous"]["value"], e2e["pipelined"]["value"]) and math.isfinite(e2e["last_loss"])


def test_a_ps_only_rank_steps_without_batches_and_skips_the_worker_only_arms():
    args = argparse.Namespace(num_train=5000, e2e_prefetch=1, e2e_pipeline=1, e2e_native_loop=1, e2e_depth=4, min_ms=2.0, max_reps=20)
    eng = FakeEngine(50)
    eng.worker_ranks = [1]                                   # rank 0 of this process hosts the ps shard only
    spec = SimpleNamespace(batch=100, in_dim=784, classes=10)
    e2e = bench.run_e2e(eng, args, spec, 20, 1, False, 2, False, None, None, barrier=lambda: None,
                        allmax=lambda v: [float(x) for x in v], timer=WallTimer(), pin=torch.from_numpy)
    assert e2e["value"] > 0 and "native_loop" not in e2e and "pipelined" not in e2e and set(eng.seen) == {-1}


def test_every_rank_a_worker_topology_runs_all_three_arms_on_a_multi_rank_job():
    nb = 50
    images = np.repeat(np.arange(nb, dtype=np.float32), 100)[:, None] * np.ones((1, 784), np.float32)
    labels = np.zeros((nb * 100, 10), np.float32)
    args = argparse.Namespace(num_train=nb * 100, e2e_prefetch=1, e2e_pipeline=1, e2e_native_loop=1, e2e_depth=4, min_ms=2.0, max_reps=20)
    eng = FakeEngine(nb)
    eng.ranks = {3: SimpleNamespace(device=None, stream=None)}
    eng.worker_ranks = [0, 1, 2, 3]                          # this process is worker 3 of 4: batches 3, 7, 11, ...
    spec = SimpleNamespace(batch=100, in_dim=784, classes=10)
    e2e = bench.run_e2e(eng, args, spec, 20, 4, True, 4, True, images, labels, barrier=lambda: None,
                        allmax=lambda v: [float(x) for x in v], timer=WallTimer(), pin=torch.from_numpy)
    assert {"synchronous", "pipelined", "native_loop"} <= set(e2e)
    assert eng.seen == [(4 * t + 3) % 50 for t in range(len(eng.seen))]
    assert all(c[2] == 4 for c in eng.loop_calls)
    assert e2e["value"] == pytest.approx(4 * 100 * 20 / (e2e["ms_per_step"] * 20 / 1e3))     # whole-job samples/s


def test_a_hung_native_loop_trips_the_dead_man_timer_which_emits_the_record_measured_so_far(monkeypatch):
    """The native arm enqueues from C: a stream / event wait that never returns cannot be cancelled from Python.  The timer's
    callback receives the e2e record of the arms measured before it (with the reason under native_loop_error) and ends the
    process; here the exit is replaced by releasing the stand-in loop."""
    import threading
    released, emitted, exits = threading.Event(), [], []

    class Hanging(FakeEngine):
        def train_loop(self, *a, **k):
            released.wait(20)
            raise RuntimeError("released by the test")

    def fake_exit(code):
        exits.append(code)
        released.set()
    real = bench.DeadMan
    monkeypatch.setattr(bench, "DeadMan", lambda s, f, exit_fn=None: real(s, f, exit_fn=fake_exit))
    nb = 50
    images = np.repeat(np.arange(nb, dtype=np.float32), 100)[:, None] * np.ones((1, 784), np.float32)
    labels = np.zeros((nb * 100, 10), np.float32)
    args = argparse.Namespace(num_train=nb * 100, e2e_prefetch=1, e2e_pipeline=1, e2e_native_loop=1, e2e_depth=4, min_ms=2.0, max_reps=20,
                              e2e_native_timeout=0.3)
    spec = SimpleNamespace(batch=100, in_dim=784, classes=10)
    e2e = bench.run_e2e(Hanging(nb), args, spec, 20, 1, True, 1, False, images, labels, barrier=lambda: None,
                        allmax=lambda v: [float(x) for x in v], timer=WallTimer(), pin=torch.from_numpy, failsafe=emitted.append)
    assert exits == [0] and len(emitted) == 1
    rec = emitted[0]
    assert "dead-man timer" in rec["native_loop_error"] and "native_loop" not in rec
    assert rec["value"] == max(rec["synchronous"]["value"], rec["pipelined"]["value"]) and math.isfinite(rec["last_loss"])
    assert "native_loop" not in e2e                       # (the released loop raised: the arm is dropped on the normal path too)


def test_the_timer_is_cancelled_when_the_arm_returns():
    fired = []
    g = bench.DeadMan(0.2, lambda: fired.append(1), exit_fn=lambda c: fired.append(("exit", c)))
    g.cancel()
    time.sleep(0.4)
    assert fired == [] and not g.fired
    assert bench.DeadMan(0.0, lambda: None)._t is None and bench.DeadMan(5.0, None)._t is None      # disabled: no thread
