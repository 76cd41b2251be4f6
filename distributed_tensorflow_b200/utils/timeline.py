"""Step tracing and chrome-trace export (aux "Tracing / profiling").

``RunOptions(trace_level=FULL_TRACE)`` makes every participating task record
one event per executed node (host wall-clock start/end; for CUDA outputs also
a device-side duration from CUDA events); the events come back in
``RunMetadata.step_stats`` and ``Timeline(step_stats).generate_chrome_trace_format()``
renders ``chrome://tracing`` JSON with **one pid per /job/task device** --
reference ``example_in_graph.py:42-43,59,65-68``.

The fabric engine's kernels additionally stamp ``%globaltimer`` into a
per-rank ring buffer (``csrc/ps_apply.cu``), which :func:`events_from_ring`
turns into the same event records, so device-resident steps show up in the
same timeline.
"""
from __future__ import annotations

import json
import threading
import time
from typing import Any, Dict, Iterable, List, Optional

import torch

__all__ = ["StepTracer", "Timeline", "events_from_ring"]

_EPOCH_NS = time.time_ns() - time.perf_counter_ns()      # perf_counter -> wall clock, so tasks line up


class StepTracer:
    """Collects per-node events on one task for one ``Session.run``."""

    def __init__(self, task_name: str):
        self.task_name = task_name
        self._events: List[Dict[str, Any]] = []
        self._cuda = []
        self._lock = threading.Lock()

    def record(self, node, ctx, t0_ns: int, t1_ns: int, out) -> None:
        dev = node.device or ""
        if isinstance(out, torch.Tensor):
            tdev = str(out.device)
            nbytes = out.numel() * out.element_size()
            shape = list(out.shape)
        else:
            tdev, nbytes, shape = "cpu", 0, None
        ev = {"task": self.task_name, "device": "%s (%s)" % (dev or self.task_name, tdev), "name": node.name,
              "op": node.op_type, "inputs": [i.name for i in node.inputs], "start_us": (t0_ns + _EPOCH_NS) / 1e3,
              "dur_us": max((t1_ns - t0_ns) / 1e3, 0.001), "bytes": nbytes, "shape": shape,
              "thread": threading.get_ident() % 100000}
        with self._lock:
            self._events.append(ev)

    def add_event(self, ev: Dict[str, Any]) -> None:
        with self._lock:
            self._events.append(ev)

    def events(self) -> List[Dict[str, Any]]:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        with self._lock:
            return list(self._events)


def events_from_ring(task_name: str, ring: Iterable, names: Dict[int, str], gpu_index: int = 0) -> List[Dict[str, Any]]:
    """Convert (kind, t_start_ns, t_end_ns, step) device records into timeline events."""
    out = []
    for kind, t0, t1, step in ring:
        if t1 < t0 or not t0:
            continue
        out.append({"task": task_name, "device": "%s/device:GPU:%d (kernels)" % (task_name, gpu_index),
                    "name": "%s[step %d]" % (names.get(int(kind), "k%d" % kind), step), "op": names.get(int(kind), "kernel"),
                    "inputs": [], "start_us": t0 / 1e3, "dur_us": max((t1 - t0) / 1e3, 0.001), "bytes": 0, "shape": None,
                    "thread": 0})
    return out


class Timeline:
    def __init__(self, step_stats, graph=None):
        self._events = list(step_stats.step_stats if hasattr(step_stats, "step_stats") else step_stats)

    def generate_chrome_trace_format(self, show_dataflow: bool = True, show_memory: bool = False) -> str:
        pids: Dict[str, int] = {}
        trace: List[Dict[str, Any]] = []
        if not self._events:
            return json.dumps({"traceEvents": []})
        t_min = min(e["start_us"] for e in self._events)
        end_of: Dict[str, Any] = {}
        for e in self._events:
            dev = e["device"]
            if dev not in pids:
                pids[dev] = len(pids)
                trace.append({"name": "process_name", "ph": "M", "pid": pids[dev], "args": {"name": dev}})
            pid = pids[dev]
            ts = e["start_us"] - t_min
            trace.append({"name": e["name"], "cat": "Op", "ph": "X", "pid": pid, "tid": e.get("thread", 0),
                          "ts": ts, "dur": e["dur_us"],
                          "args": {"op": e["op"], "name": e["name"], "inputs": e["inputs"], "shape": e["shape"],
                                   "bytes": e["bytes"]}})
            end_of[e["name"]] = (pid, e.get("thread", 0), ts + e["dur_us"])
        if show_dataflow:
            fid = 0
            for e in self._events:
                pid = pids[e["device"]]
                ts = e["start_us"] - t_min
                for src in e["inputs"]:
                    s = end_of.get(src)
                    if s is None or s[0] == pid:
                        continue            # only cross-device edges (the interesting Send/Recv ones)
                    trace.append({"name": src, "cat": "DataFlow", "ph": "s", "id": fid, "pid": s[0], "tid": s[1],
                                  "ts": s[2]})
                    trace.append({"name": src, "cat": "DataFlow", "ph": "t", "id": fid, "pid": pid,
                                  "tid": e.get("thread", 0), "ts": max(ts, s[2])})
                    fid += 1
        return json.dumps({"traceEvents": trace}, indent=None)
