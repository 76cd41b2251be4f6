"""csrc/step_exec.cu (the native step executor behind ``engine.step(x_pinned, y_pinned)``) compiled by g++ against
tests/emu/step_exec_stubs.h: the CUDA runtime calls and kernel launchers record a trace.  Pins the argument mapping of every
op kind, the kernel count, the error index, device switching and the graph-capture sequence."""
import ctypes
import os
import shutil
import subprocess

import pytest

from distributed_tensorflow_b200.ops.cuda_lib import (OP_CONVERT, OP_D2H, OP_GEMM, OP_GRAPH, OP_H2D, OP_HEAD, OP_PS_APPLY, OP_SIGNAL, OP_STAGE,
                                                     OP_SYNC, OP_WAIT_TOKEN, StepOp)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OP_EVENT_RECORD, OP_EVENT_WAIT = 11, 12


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu") / "libstep_host.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-DDTF_HOST_EMU", "-I" + os.path.join(ROOT, "tests", "emu"), "-x", "c++", "-shared",
                    "-fPIC", "-o", so, os.path.join(ROOT, "distributed_tensorflow_b200", "csrc", "step_exec.cu")], check=True)
    lib = ctypes.CDLL(so)
    lib.dtf_run_ops.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.dtf_capture_ops.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p),
                                    ctypes.POINTER(ctypes.c_int)]
    lib.step_emu_trace.restype = ctypes.c_char_p
    lib.step_emu_reset.argtypes = [ctypes.c_int, ctypes.c_int]
    return lib


def _run(host, ops, device=-1, stream=0x57):
    arr = (StepOp * len(ops))(*ops)
    k = ctypes.c_int(0)
    rc = host.dtf_run_ops(arr, len(ops), device, stream, ctypes.byref(k))
    return rc, k.value, host.step_emu_trace().decode().splitlines()


def test_every_op_kind_maps_its_arguments(host):
    assert host.dtf_sizeof_step_op() == ctypes.sizeof(StepOp)
    host.step_emu_reset(0, 0)
    ops = [StepOp(kind=OP_EVENT_WAIT, p0=0xe0),
           StepOp(kind=OP_H2D, p0=0xd0, p1=0xa0, i0=313600),
           StepOp(kind=OP_CONVERT, p0=0xd0, p1=0xd1, i0=784, i1=784, i2=100, i3=784, i4=784),
           StepOp(kind=OP_EVENT_RECORD, p0=0xe1),
           StepOp(kind=OP_WAIT_TOKEN, p0=0xb0, p1=0xb1, p2=0xb2, i0=3, u0=5000),
           StepOp(kind=OP_GEMM, p0=0x100), StepOp(kind=OP_HEAD, p0=0x200),
           StepOp(kind=OP_SIGNAL, p0=0xc0, p1=0xc1, i0=4, i1=1),
           StepOp(kind=OP_STAGE, p0=0x1, p1=0x2, p2=0x3, p3=0x4, p4=0x5, i0=550, i1=100, i2=784, i3=10, i4=7, i5=2),
           StepOp(kind=OP_PS_APPLY, p0=0x300),
           StepOp(kind=OP_D2H, p0=0xa1, p1=0xd2, i0=32), StepOp(kind=OP_GRAPH, p0=0x900, i0=5), StepOp(kind=OP_SYNC)]
    rc, kernels, trace = _run(host, ops)
    assert rc == 0 and kernels == 7 + 5                       # convert, wait, gemm, head, signal, stage, apply + 5 inside the graph
    assert trace == [
        "event_wait ev=0xe0 stream=0x57",
        "memcpy h2d dst=0xd0 src=0xa0 n=313600 stream=0x57",
        "convert in=0xd0 ld_in=784 out=0xd1 ld_out=784 rows=100 cols=784 pad=784",
        "event_record ev=0xe1 stream=0x57",
        "wait_token mb=0xb0 target=3 ptr=0xb1 timeout=5000 err=0xb2",
        "gemm args=0x100 stream=0x57", "head args=0x200 stream=0x57",
        "push_grad src=(nil) dst=(nil) n=0 ctl=0xc0 mb=0xc1 rank=4 from_version=1 write_stamp=1 grid=1",
        "stage images=0x1 labels=0x2 nb=550 B=100 D=784 C=10 stride=7 offset=2 ctr=0x3 x16=0x4 lab=0x5",
        "ps_apply args=0x300 stream=0x57",
        "memcpy d2h dst=0xa1 src=0xd2 n=32 stream=0x57",
        "graph_launch exec=0x900 stream=0x57", "sync stream=0x57"]


def test_error_index_device_switch_and_capture(host):
    host.step_emu_reset(4, 77)                                # the GEMM launcher fails with code 77
    rc, kernels, trace = _run(host, [StepOp(kind=OP_HEAD, p0=1), StepOp(kind=OP_GEMM, p0=2), StepOp(kind=OP_PS_APPLY, p0=3)], device=3)
    assert rc == 2 * 100000 + 77 and kernels == 2             # (index + 1) * 100000 + code; nothing after the failure runs
    assert trace[0] == "setdevice 3" and trace[-1] == "setdevice 0" and not any(l.startswith("ps_apply") for l in trace)
    host.step_emu_reset(0, 0)
    assert _run(host, [StepOp(kind=99)])[0] == 100000 + 99
    host.step_emu_reset(0, 0)
    ops = (StepOp * 2)(StepOp(kind=OP_GEMM, p0=1), StepOp(kind=OP_HEAD, p0=2))
    ex, nk = ctypes.c_void_p(), ctypes.c_int(0)
    assert host.dtf_capture_ops(ops, 2, -1, 0x57, ctypes.byref(ex), ctypes.byref(nk)) == 0
    assert ex.value == 0xe1 and nk.value == 2
    assert host.step_emu_trace().decode().splitlines() == ["begin_capture stream=0x57", "gemm args=0x1 stream=0x57", "head args=0x2 stream=0x57",
                                                           "end_capture stream=0x57", "instantiate graph=0x6a", "graph_destroy 0x6a"]


def _loop_args(host, steps, depth, prefetched=0, ps=True, parity=0, first=3, batch_step=2, nb=5):
    from distributed_tensorflow_b200.ops.cuda_lib import LoopArgs
    assert host.dtf_sizeof_loop_args() == ctypes.sizeof(LoopArgs)
    host.dtf_loop_release_events()            # the loop caches its events per device across calls: every test starts without them
    host.step_emu_reset(0, 0)
    host.dtf_run_loop.argtypes = [ctypes.POINTER(LoopArgs)]
    keep = []
    a = LoopArgs()
    a.device, a.steps, a.depth, a.parity, a.prefetched, a.x_op, a.y_op = 2, steps, depth, parity, prefetched, 1, 2
    for p in range(2):
        cops = (StepOp * 4)(StepOp(kind=OP_EVENT_WAIT, p0=0xd0 + p), StepOp(kind=OP_H2D, p0=0x1000 + p, i0=400),
                            StepOp(kind=OP_H2D, p0=0x2000 + p, i0=40), StepOp(kind=OP_EVENT_RECORD, p0=0xa0 + p))
        mops = (StepOp * 3)(StepOp(kind=OP_EVENT_WAIT, p0=0xa0 + p), StepOp(kind=OP_GRAPH, p0=0x900 + p, i0=2),
                            StepOp(kind=OP_EVENT_RECORD, p0=0xd0 + p))
        keep += [cops, mops]
        a.n_copy[p], a.n_compute[p] = 4, 3
        a.copy_ops[p], a.compute_ops[p] = ctypes.addressof(cops), ctypes.addressof(mops)
    if ps:
        pops = (StepOp * 1)(StepOp(kind=OP_PS_APPLY, p0=0x300))
        keep.append(pops)
        a.n_ps, a.ps_ops, a.ps_stream = 1, ctypes.addressof(pops), 0x59
    a.copy_stream, a.stream = 0x58, 0x57
    a.x_base, a.y_base, a.x_stride, a.y_stride = 0x100000, 0x200000, 0x1000, 0x100
    a.nbatches, a.first, a.batch_step = nb, first, batch_step
    a.loss_src, a.loss_bytes, a.loss_host, a.loss_row_bytes = 0xbeef, 28, 0x300000, 64
    return a, keep


def test_native_loop_orders_copies_kernels_loss_reads_and_bounds_the_run_ahead(host):
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=4, depth=2)
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    t = host.step_emu_trace().decode().splitlines()
    batches = [(3 + 2 * i) % 5 for i in range(4)]                     # 3, 0, 2, 4

    def copy(par, b):
        return ["event_wait ev=0x%x stream=0x58" % (0xd0 + par),
                "memcpy h2d dst=0x%x src=0x%x n=400 stream=0x58" % (0x1000 + par, 0x100000 + b * 0x1000),
                "memcpy h2d dst=0x%x src=0x%x n=40 stream=0x58" % (0x2000 + par, 0x200000 + b * 0x100),
                "event_record ev=0x%x stream=0x58" % (0xa0 + par)]

    def compute(par):
        return ["event_wait ev=0x%x stream=0x57" % (0xa0 + par), "graph_launch exec=0x%x stream=0x57" % (0x900 + par),
                "event_record ev=0x%x stream=0x57" % (0xd0 + par), "ps_apply args=0x300 stream=0x59"]

    def loss(i, ev):
        return ["memcpy d2h dst=0x%x src=0xbeef n=28 stream=0x57" % (0x300000 + 64 * i), "event_record ev=0x%x stream=0x57" % ev]

    want = ["setdevice 2", "event_create ev=0xe000 flags=2", "event_create ev=0xe001 flags=2"]
    want += copy(0, batches[0])
    want += compute(0) + copy(1, batches[1]) + loss(0, 0xe000)
    want += compute(1) + copy(0, batches[2]) + loss(1, 0xe001)
    want += ["event_sync ev=0xe000"] + compute(0) + copy(1, batches[3]) + loss(2, 0xe000)     # step 2 waits for the loss of step 0
    want += ["event_sync ev=0xe001"] + compute(1) + loss(3, 0xe001)                            # no copy after the last step
    want += ["event_sync ev=0xe001", "setdevice 0"]                      # the events stay cached for the next call
    assert t == want
    assert (a.parity, a.prefetched, a.kernels, a.waited) == (0, 0, 4 * 3, 3)
    host.step_emu_reset(0, 0)                                           # a second call reuses the cached events
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    t2 = host.step_emu_trace().decode().splitlines()
    assert not any(l.startswith(("event_create", "event_destroy")) for l in t2) and t2[1:] == want[3:]
    assert host.dtf_loop_release_events() == 2


def test_native_loop_prefetched_first_batch_odd_parity_short_runs_and_errors(host):
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=1, depth=4, prefetched=1, ps=False, parity=1)
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    t = host.step_emu_trace().decode().splitlines()
    assert t == ["setdevice 2", "event_create ev=0xe000 flags=2",
                 "event_wait ev=0xa1 stream=0x57", "graph_launch exec=0x901 stream=0x57", "event_record ev=0xd1 stream=0x57",
                 "memcpy d2h dst=0x300000 src=0xbeef n=28 stream=0x57", "event_record ev=0xe000 stream=0x57",
                 "event_sync ev=0xe000", "setdevice 0"]
    assert (a.parity, a.kernels) == (0, 2)
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=0, depth=2)
    assert host.dtf_run_loop(ctypes.byref(a)) == 0 and a.kernels == 0
    assert not any(l.startswith(("memcpy", "graph")) for l in host.step_emu_trace().decode().splitlines())
    for bad in (dict(depth=0), dict(depth=65), dict(nb=0)):
        a, keep = _loop_args(host, steps=2, **{"depth": 2, **bad})
        assert host.dtf_run_loop(ctypes.byref(a)) == -1
    a, keep = _loop_args(host, steps=2, depth=2)
    a.x_op = 9
    assert host.dtf_run_loop(ctypes.byref(a)) == -2
    a, keep = _loop_args(host, steps=3, depth=2)
    host.step_emu_reset(100, 2)                                        # event creation fails: nothing is enqueued
    assert host.dtf_run_loop(ctypes.byref(a)) == 2
    assert not any(l.startswith(("memcpy", "graph", "event_destroy")) for l in host.step_emu_trace().decode().splitlines())


def test_native_loop_prefetch_next_copies_the_batch_after_the_last_step(host):
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=2, depth=2, ps=False)
    a.prefetch_next = 1
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    t = host.step_emu_trace().decode().splitlines()
    h2d = [l for l in t if l.startswith("memcpy h2d") and "n=400" in l]
    batches = [(3 + 2 * i) % 5 for i in range(3)]
    assert [int(l.split("src=")[1].split()[0], 16) for l in h2d] == [0x100000 + b * 0x1000 for b in batches]     # steps 0, 1 and the one after
    assert h2d[2].split("dst=")[1].split()[0] == "0x1000"                 # ... into the buffer set the next step will use (parity 0)
    assert (a.parity, a.prefetched) == (0, 1)
    last_graph = max(i for i, l in enumerate(t) if l.startswith("graph_launch"))
    assert t.index(h2d[2]) > last_graph                                   # issued behind the last step's kernels


def _happens_before(trace):
    """CUDA ordering rules applied to a recorded trace: ops issued to one stream run in issue order; ``event_wait`` makes the
    following ops of its stream depend on the LAST ``event_record`` of that event issued before the wait; ``event_sync`` makes
    every later host call depend on that record.  Returns (ops, reach) with reach[i] = set of op indices that happen before i."""
    ops, last_in_stream, last_record, host_dep = [], {}, {}, set()
    deps = []
    for line in trace:
        kind = line.split()[0]
        f = dict(kv.split("=") for kv in line.split()[1:] if "=" in kv)
        if kind == "event_sync":
            if f["ev"] in last_record:
                host_dep = host_dep | {last_record[f["ev"]]}
            continue
        if kind not in ("memcpy", "graph_launch", "event_record", "event_wait", "ps_apply"):
            continue
        i = len(ops)
        ops.append((kind, f, line))
        d = set(host_dep)
        s = f["stream"]
        if s in last_in_stream:
            d.add(last_in_stream[s])
        if kind == "event_wait" and f["ev"] in last_record:
            d.add(last_record[f["ev"]])
        deps.append(d)
        last_in_stream[s] = i
        if kind == "event_record":
            last_record[f["ev"]] = i
    reach = []
    for i, d in enumerate(deps):
        r = set(d)
        for j in d:
            r |= reach[j]
        reach.append(r)
    return ops, reach


@pytest.mark.parametrize("steps,depth,prefetched,prefetch_next", [(7, 2, 0, 0), (6, 4, 0, 1), (5, 1, 1, 1), (1, 3, 0, 0)])
def test_native_loop_has_no_buffer_hazard_under_cuda_ordering_rules(host, steps, depth, prefetched, prefetch_next):
    """The double-buffered staging is only safe if (a) the copy of step i's batch happens before step i's kernels, (b) the kernels
    that read a buffer set happen before the NEXT copy into that set, (c) a loss row is copied after its step's kernels and before
    the next step's kernels overwrite the device-side partials -- derived from the trace with CUDA's stream / event rules, not
    from the order of the host calls."""
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=steps, depth=depth, prefetched=prefetched, ps=True)
    a.prefetch_next = prefetch_next
    pre = []
    if prefetched:                                   # what step()'s prefetch would have issued before the call
        pre = ["event_wait ev=0xd0 stream=0x58", "memcpy h2d dst=0x1000 src=0x103000 n=400 stream=0x58",
               "memcpy h2d dst=0x2000 src=0x200300 n=40 stream=0x58", "event_record ev=0xa0 stream=0x58"]
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    ops, reach = _happens_before(pre + host.step_emu_trace().decode().splitlines())
    computes = [i for i, (k, f, _) in enumerate(ops) if k == "graph_launch"]
    copies_x = [i for i, (k, f, l) in enumerate(ops) if k == "memcpy" and "h2d" in l and f["n"] == "400"]
    loss = [i for i, (k, f, l) in enumerate(ops) if k == "memcpy" and "d2h" in l]
    assert len(computes) == steps and len(loss) == steps and len(copies_x) == steps + prefetch_next
    for i, c in enumerate(computes):
        par = i % 2
        assert ops[c][1]["exec"] == "0x%x" % (0x900 + par)
        assert copies_x[i] in reach[c], "step %d runs before its batch has landed" % i                       # (a)
        assert ops[copies_x[i]][1]["dst"] == "0x%x" % (0x1000 + par)
        if i + 2 < len(copies_x):
            assert c in reach[copies_x[i + 2]], "batch %d overwrites the buffers step %d still reads" % (i + 2, i)   # (b)
        assert c in reach[loss[i]]                                                                           # (c)
        if i + 1 < steps:
            assert loss[i] in reach[computes[i + 1]]
        assert ops[loss[i]][1]["dst"] == "0x%x" % (0x300000 + 64 * i)


def test_the_hazard_checker_notices_a_missing_dependency(host):
    """Without the copy stream's wait on done[p] the rule (b) above must fail: the checker is not vacuous."""
    host.step_emu_reset(0, 0)
    a, keep = _loop_args(host, steps=5, depth=2, ps=False)
    assert host.dtf_run_loop(ctypes.byref(a)) == 0
    trace = [l for l in host.step_emu_trace().decode().splitlines() if not (l.startswith("event_wait") and "stream=0x58" in l)]
    ops, reach = _happens_before(trace)
    computes = [i for i, (k, f, _) in enumerate(ops) if k == "graph_launch"]
    copies_x = [i for i, (k, f, l) in enumerate(ops) if k == "memcpy" and "h2d" in l and f["n"] == "400"]
    assert any(computes[i] not in reach[copies_x[i + 2]] for i in range(3))
